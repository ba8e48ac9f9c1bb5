"""Small run of the kernels added in round 2 for compute-sanitizer (memcheck / racecheck) on the GPU box: one training step
(every native convolution direction, BatchNorm + dropout + pooling passes, GRU recurrence and back-propagation through time),
the n_fft = 256 clip path, the CRNN forward with the staged DSMEM broadcasts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import salsa_b200
from oracle import synth
from salsa_b200 import train

torch.cuda.set_device(0)
g = torch.Generator().manual_seed(0)
B, T = 3, 64                                     # 3 clips: five empty clip slots in the GRU cluster's group of 8
x = torch.randn(B, 7, T, 200, generator=g).cuda()
tgt = {'event_frame_gt': (torch.rand(B, T // 8, 12, generator=g) > 0.6).float().cuda(),
       'doa_frame_gt': torch.randn(B, T // 8, 36, generator=g).clamp(-1, 1).cuda()}
tr = train.SeldTrainer(salsa_b200.crnn.random_state_dict(0))
for _ in range(2):
    loss = tr.step(x, tgt)
torch.cuda.synchronize()
print('train step', loss.tolist())
audio = torch.from_numpy(np.stack([synth.make_clip(i, 'foa', seconds=0.6) for i in range(2)])).cuda()
out = salsa_b200.SalsaExtractor('foa', n_fft=256, hop_len=150, win_len=256).extract(audio)
torch.cuda.synchronize()
print('n_fft 256', tuple(out.shape), float(out[:, 4:].abs().sum()))
for precision in ('bf16', 'bf16x2'):
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256),
                             precision=precision)
    m.load_state_dict(salsa_b200.crnn.random_state_dict(0))
    y = m.forward_ops(torch.randn(3, 7, 64, 200, generator=g).cuda())
    torch.cuda.synchronize()
    print('forward', precision, {k: tuple(v.shape) for k, v in y.items()})
