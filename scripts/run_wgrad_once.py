#!/usr/bin/env python
"""One launch of crnn_conv_wgrad at the conv_block1.conv2 training shape (8 x 640 x 200 x 64 -> 64), for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from salsa_b200 import crnn_ops as ops

x = torch.randn(8, 640, 200, 64, device='cuda').bfloat16()
gy = torch.randn(8, 640, 200, 64, device='cuda').bfloat16()
for _ in range(2):
    ops.conv_wgrad(x, gy)
torch.cuda.synchronize()
