#!/usr/bin/env python
"""Per-operator CUDA-event timing of the CRNN forward (run on the GPU box).

    python scripts/profile_crnn.py [--batch 8] [--frames 4800] [--freq 200]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import salsa_b200
from salsa_b200 import crnn_ops as ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--frames', type=int, default=4800)
    ap.add_argument('--freq', type=int, default=200)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256))
    m.load_state_dict(salsa_b200.crnn.random_state_dict(0))
    x = torch.randn(args.batch, 7, args.frames, args.freq, device='cuda')
    records = []
    orig = {}

    def wrap(name):
        fn = getattr(ops, name)
        orig[name] = fn

        def timed(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record()
            shape = tuple(a[0].shape)
            wshape = tuple(a[1].shape) if len(a) > 1 and hasattr(a[1], 'shape') else ()
            records.append((name, shape, wshape, s, e))
            return out
        setattr(ops, name, timed)

    for name in ('pack_input', 'conv_first', 'conv2d', 'avgpool2', 'freq_mean', 'gemm', 'gru_layer', 'head_finish'):
        wrap(name)
    for rep in range(args.reps):
        records.clear()
        m.forward_ops(x)
        torch.cuda.synchronize()
    total = 0.0
    for name, shape, wshape, s, e in records:
        ms = s.elapsed_time(e)
        total += ms
        flops = 0.0
        if name == 'conv_first':
            B, H, W, _ = shape
            flops = 2.0 * B * H * W * 7 * 64 * 9
        elif name == 'conv2d':
            B, H, W, Cin = shape
            taps, Cout, _ = wshape
            flops = 2.0 * B * H * W * Cin * Cout * taps
        elif name == 'gemm':
            flops = 2.0 * shape[0] * shape[1] * wshape[0]
        print('{:12s} in {:28s} w {:20s} {:9.3f} ms {:8.1f} TFLOP/s'.format(name, str(shape), str(wshape), ms, flops / ms / 1e9 if ms > 0 else 0))
    print('total {:.3f} ms for {} clips -> {:.1f} clips/s'.format(total, args.batch, args.batch / total * 1e3))


if __name__ == '__main__':
    main()
