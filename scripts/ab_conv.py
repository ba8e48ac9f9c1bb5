"""A/B test of convolution kernel options on the GPU box: correctness vs PyTorch and speed."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from salsa_b200 import crnn_ops as ops, _native

def opt(name, v):
    _native.check(_native.lib().crnn_set_option(name.encode(), v))

def check(B, H, W, Cin, Cout, res=False):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).bfloat16()
    bias = torch.randn(Cout, generator=g)
    r = torch.randn(B, H, W, Cout, generator=g).bfloat16()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1) + bias
    if res:
        ref = ref + r.float()
    ref = ref.clamp(min=0)
    wp = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    out = ops.conv2d(x.cuda(), wp.cuda(), bias.cuda(), residual=r.cuda() if res else None, relu=True).cpu().float()
    return float((out - ref).abs().max() / ref.abs().max())

def speed(B, H, W, Cin, Cout, res=False, reps=5):
    x = torch.randn(B, H, W, Cin, device='cuda').bfloat16()
    wp = torch.randn(9, Cout, Cin, device='cuda').bfloat16()
    bias = torch.randn(Cout, device='cuda')
    r = torch.randn(B, H, W, Cout, device='cuda').bfloat16() if res else None
    out = torch.empty(B, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    ops.conv2d(x, wp, bias, residual=r, relu=True, out=out)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        ops.conv2d(x, wp, bias, residual=r, relu=True, out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    return ms, 2.0 * B * H * W * Cin * Cout * 9 / ms / 1e9

for tma in (0, 1):
    opt('tma_store', tma)
    errs = [check(1, 37, 21, 64, 64), check(2, 16, 8, 64, 64, True), check(1, 33, 12, 128, 256, True), check(1, 20, 9, 256, 512)]
    sp = [speed(8, 4800, 200, 64, 64), speed(8, 2400, 100, 64, 64, True), speed(8, 1200, 50, 128, 128, True), speed(8, 600, 25, 256, 256, True), speed(8, 300, 12, 512, 512, True)]
    print('tma_store', tma, 'errors', ['%.1e' % e for e in errs], 'ms/TFLOPs', ['%.3f/%.0f' % s for s in sp])
