"""bench.py contract checks that run without a GPU: the reference arm prints exactly one JSON line with the keys the
driver reads, and the roofline constants match SURVEY.md section 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--cpu-seconds', '0.5'], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, check=True)
    lines = [ln for ln in out.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'audio-min/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('SALSA FOA batch: 600 synthetic')
    for key in ('metric', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data'):
        assert key in d


def test_algorithmic_byte_counts():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.AUDIO_BYTES_PER_CLIP == 23040000
    assert bench.feature_bytes_per_clip(200) == 26885600            # SALSA: 7 x 4801 x 200 float32
    assert bench.feature_bytes_per_clip(191) == 25675748            # SALSA-Lite
    assert abs(bench.CRNN_CONV_FLOP_PER_CLIP - 335.52e9) < 0.1e9
