#!/usr/bin/env python
"""Two eager training steps (batch 8 x (7, 640, 200)) for ncu captures of the training-step kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import salsa_b200
from salsa_b200 import train

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 7, 640, 200, generator=g).cuda()
tgt = {'event_frame_gt': (torch.rand(B, 80, 12, generator=g) > 0.7).float().cuda(), 'doa_frame_gt': (torch.rand(B, 80, 36, generator=g) * 2 - 1).cuda()}
tr = train.SeldTrainer(salsa_b200.crnn.random_state_dict(0))
for _ in range(2):
    print(tr.step(x, tgt).tolist())
torch.cuda.synchronize()
