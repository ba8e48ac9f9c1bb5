"""A/B of the two arrangements of the clip path on the GPU box (SALSA_B200_PIPELINE = split | fused) and of the
eig_rows_kernel occupancy target: per-kernel CUDA-event times per step and bit-level agreement of the features."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import salsa_b200
from salsa_b200 import _native


def run(audio, env, steps=3):
    for k in ('SALSA_B200_PIPELINE',):
        os.environ.pop(k, None)
    os.environ.update(env)
    ex = salsa_b200.SalsaExtractor('foa')
    feat = ex.extract(audio)
    ex.extract(audio, out=feat)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        ex.extract(audio, out=feat)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    # per-kernel times in a second pass (profiling serialises the kernels of the path)
    _native.profile_enable(True)
    _native.profile_read()
    for _ in range(steps):
        ex.extract(audio, out=feat)
    torch.cuda.synchronize()
    kernels = {k: round(v[0] / steps, 3) for k, v in _native.profile_read().items()}
    _native.profile_enable(False)
    return feat, {'env': env, 'ms_per_step': round(ms, 3), 'clips_per_s': round(audio.shape[0] / ms * 1e3, 1), 'kernels_ms': kernels}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    torch.cuda.set_device(0)
    audio = bench.make_clips(torch, n, 'foa', torch.device('cuda:0'), 0)
    base = None
    variants = [{'SALSA_B200_PIPELINE': 'fused'}, {}]
    variants += [{'SALSA_B200_STFT_VARIANT': v} for v in os.environ.get('AB_STFT_VARIANTS', '').split(',') if v]
    for env in variants:
        feat, rec = run(audio, dict(env))
        if base is None:
            base = feat.clone()
        else:
            same_bits = bool(torch.equal(feat.view(torch.int32), base.view(torch.int32)))
            rec['spec_bits_equal_to_fused'] = bool(torch.equal(feat[:, :4].view(torch.int32), base[:, :4].view(torch.int32)))
            d = (feat[:, :4] - base[:, :4]).abs()
            rec['spec_values_differing'] = int((d != 0).sum().item())
            rec['spec_max_abs_diff_db'] = float(d.max().item())
            rec['mask_mismatches_vs_fused'] = int(((feat[:, 4:] != 0) != (base[:, 4:] != 0)).sum().item())
            rec['max_abs_diff_spatial'] = float((feat[:, 4:] - base[:, 4:]).abs().max().item())
            rec['all_bits_equal_to_fused'] = same_bits
        print(json.dumps(rec), flush=True)
        del feat


if __name__ == '__main__':
    main()
