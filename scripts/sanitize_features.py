"""Small feature-path run for compute-sanitizer (memcheck / racecheck) on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import salsa_b200
from oracle import synth

torch.cuda.set_device(0)
for fmt, fmax in (('foa', 9000), ('mic', 4000)):
    audio = np.stack([synth.make_clip(i, fmt, seconds=0.7) for i in range(3)])
    a = torch.from_numpy(audio).cuda()
    for tracking in (True, False):
        ex = salsa_b200.SalsaExtractor(fmt, fmax_doa=fmax, is_tracking=tracking)
        out = ex.extract(a)
        torch.cuda.synchronize()
        print(fmt, tracking, tuple(out.shape), float(out[:, 4:].abs().sum()))
lite = salsa_b200.SalsaLiteExtractor().extract(torch.from_numpy(np.stack([synth.make_clip(7, 'mic', seconds=0.7)])).cuda())
torch.cuda.synchronize()
print('lite', tuple(lite.shape))
os.environ['SALSA_B200_PIPELINE'] = 'fused'
out = salsa_b200.SalsaExtractor('foa').extract(torch.from_numpy(np.stack([synth.make_clip(1, 'foa', seconds=0.7)])).cuda())
torch.cuda.synchronize()
print('fused', tuple(out.shape))
