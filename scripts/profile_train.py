#!/usr/bin/env python
"""Kernel-time table of one training step (configs[4]) from torch.profiler: which kernels the 24 ms go to, own and library.
python scripts/profile_train.py [batch]   (run under gpurun; writes gpurun_out/train_kernels.txt)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                      # noqa: E402
import salsa_b200                 # noqa: E402
from salsa_b200 import augment, train   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda:0')
audio = bench.make_clips(torch, B, 'foa', dev, seed=9000, n_samples=8 * bench.FS, chunk=B)
ex = salsa_b200.SalsaExtractor('foa')
aug = augment.BatchAugment(augment.TfmapRandomSwapChannelFoa(n_classes=12), augment.RandomShiftUpDownNp(freq_shift_range=10))
GRAPH = len(sys.argv) > 2 and sys.argv[2] == 'graph'
tr = train.SeldTrainer(salsa_b200.crnn.random_state_dict(0), device=dev, use_graph=GRAPH)
tr.fuse_pool = os.environ.get('NO_FUSE_POOL') is None
g = torch.Generator(device=dev)
g.manual_seed(77)
tgt = {'event_frame_gt': (torch.rand((B, 80, 12), generator=g, device=dev) > 0.8).float(),
       'doa_frame_gt': torch.rand((B, 80, 36), generator=g, device=dev) * 2 - 1}
np.random.seed(1234)


def step():
    feat = ex.extract(audio)[:, :, :640]
    x, _, y_doa = aug(feat, tgt['event_frame_gt'], tgt['doa_frame_gt'])
    return tr.step(x, {'event_frame_gt': tgt['event_frame_gt'], 'doa_frame_gt': y_doa})


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print('ms per step (no profiler):', e0.elapsed_time(e1) / 5, 'graph' if GRAPH else 'eager', tr.graph_error, 'fuse_pool', tr.fuse_pool)
if GRAPH:
    sys.exit(0)
from torch.profiler import ProfilerActivity, profile   # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90)
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/train_kernels.txt', 'w') as f:
    f.write(tab)
# device-side only, kernels
rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
with open('gpurun_out/train_kernels_short.txt', 'w') as f:
    f.write('total kernel ms per step {:.2f}\n'.format(tot))
    for k, ms, c in rows[:50]:
        f.write('{:8.3f} ms {:5d}x  {}\n'.format(ms, c, k[:150]))
print(open('gpurun_out/train_kernels_short.txt').read())
