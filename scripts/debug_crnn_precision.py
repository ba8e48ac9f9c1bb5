import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import salsa_b200
from salsa_b200 import crnn_ops as ops
from oracle import crnn as ocrnn

sd = ocrnn.make_state_dict(0)
sdd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
x = ocrnn.model_input(2, (2, 7, 128, 200))
for prec in ('bf16', 'bf16x3'):
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256), precision=prec)
    m.load_state_dict(sd)
    P = m.planes
    enc = ops.merge_planes(m.encode(x.cuda()).cpu(), P).permute(0, 3, 1, 2).double()
    with torch.no_grad():
        ref64 = ocrnn.encoder_forward(sdd, x.double())
        ref32 = ocrnn.encoder_forward(sd, x)
    e = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(prec, 'encoder vs fp64 oracle', e(enc, ref64), ' fp32 oracle vs fp64', e(ref32.double(), ref64))
    # decoder on the oracle's encoder output (isolates the decoder)
    enc_in = ops.split_planes(ref32.permute(0, 2, 3, 1).contiguous(), P).cuda()
    y = m.decode(enc_in)
    y64 = ocrnn.decoder_forward(sdd, ref32.double())
    y32 = ocrnn.decoder_forward(sd, ref32)
    for k in y:
        print(prec, 'decoder only', k, e(y[k].cpu().double(), y64[k]), ' fp32 oracle vs fp64', e(y32[k].double(), y64[k]))
    yy = m(x.cuda())
    yf = ocrnn.decoder_forward(sdd, ref64)
    for k in yy:
        print(prec, 'full', k, e(yy[k].cpu().double(), yf[k]))
