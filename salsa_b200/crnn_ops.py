"""Thin torch-tensor wrappers over the CRNN entry points of libsalsa_b200.so (include/salsa_crnn.h).

Tensors are only device memory here: every function passes raw pointers and sizes through the C ABI.
Activations are bf16 NHWC, see the header for layouts.
"""
import ctypes

import numpy as np
import torch

from . import _native


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _call(name, t, *args):
    """One native call on the device (and that device's current stream) of CUDA tensor `t`; the stream is the entry
    point's last argument."""
    with _native.device_of(t) as st:
        _native.check(getattr(_native.lib(), name)(*args, st))


def _check_act(x):
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()):
        raise ValueError('activation must be a contiguous CUDA bf16 tensor')


def split_planes(t, planes):
    """float tensor (..., C) -> bf16 (..., planes*C): hi | mid | lo planes along the last axis (bf16x3 mode)."""
    t = t.float()
    parts = []
    for _ in range(planes):
        p = t.to(torch.bfloat16)
        parts.append(p)
        t = t - p.float()
    return torch.cat(parts, dim=-1).contiguous()


def merge_planes(t, planes):
    """inverse of split_planes (as float32)."""
    if planes == 1:
        return t.float()
    c = t.shape[-1] // planes
    return sum(t[..., i * c:(i + 1) * c].float() for i in range(planes))


def conv2d(x, w, bias=None, residual=None, relu=False, out=None, out_f32=False, planes=1, pool=False):
    """x (B,H,W,planes*Cin) bf16 NHWC, w (k*k,Cout,planes*Cin) bf16, bias (Cout,) fp32
    -> (B,H,W,planes*Cout) bf16, or (B,H,W,Cout) fp32 with out_f32.  pool=True also applies the 2x2 average
    pooling that follows (fused into the epilogue when planes == 1): -> (B,H//2,W//2,planes*Cout)."""
    _check_act(x)
    B, H, W, CinP = x.shape
    taps, Cout, Cin2 = w.shape
    if Cin2 != CinP or CinP % planes or taps not in (1, 9) or w.dtype != torch.bfloat16 or not w.is_contiguous():
        raise ValueError('weights must be contiguous bf16 (k*k, Cout, planes*Cin)')
    if residual is not None:
        _check_act(residual)
        if tuple(residual.shape) != (B, H, W, planes * Cout):
            raise ValueError('residual shape mismatch')
    fuse_pool = bool(pool) and planes == 1 and not out_f32 and H >= 2 and W >= 2
    if out is None:
        if out_f32:
            out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
        elif fuse_pool:
            out = torch.empty((B, H // 2, W // 2, Cout), dtype=torch.bfloat16, device=x.device)
        else:
            out = torch.empty((B, H, W, planes * Cout), dtype=torch.bfloat16, device=x.device)
    o16, o32 = (None, out) if out.dtype == torch.float32 else (out, None)
    _call('crnn_conv2d', x, _p(x), _p(w), _p(bias), _p(residual), _p(o16), _p(o32), B, H, W, CinP // planes, Cout,
                                            3 if taps == 9 else 1, int(bool(relu)), planes, int(fuse_pool))
    if pool and not fuse_pool:
        out = avgpool2(out, planes=planes)
    return out


def conv_wgrad(x, gy, ksize=3):
    """x (B,H,W,Cin) bf16 NHWC, gy (B,H,W,Cout) bf16 NHWC -> dW (ksize^2, Cout, Cin) fp32: the weight gradient of the 3x3 / pad 1
    (or 1x1) convolution y = conv(x, w), tap-major like the packed weights."""
    _check_act(x)
    _check_act(gy)
    B, H, W, Cin = x.shape
    if tuple(gy.shape[:3]) != (B, H, W):
        raise ValueError('x and gy must share batch and spatial dimensions')
    Cout = gy.shape[3]
    dw = torch.empty((ksize * ksize, Cout, Cin), dtype=torch.float32, device=x.device)
    _call('crnn_conv_wgrad', x, _p(x), _p(gy), _p(dw), B, H, W, Cin, Cout, ksize)
    return dw


def _drop_args(drop):
    """drop = None or (seed: CUDA int64 tensor of one element, salt: int, p: float) -> the three C arguments."""
    if drop is None or drop[2] <= 0.0:
        return ctypes.c_void_p(0), 0, ctypes.c_float(0.0)
    seed, salt, p = drop
    if not (seed.is_cuda and seed.dtype == torch.int64 and seed.numel() == 1):
        raise ValueError('the dropout seed must be a CUDA int64 tensor of one element')
    return _p(seed), int(salt) & 0xffffffff, ctypes.c_float(float(p))


def bn_train_forward(y, gamma, beta, residual=None, relu=True, running_mean=None, running_var=None, momentum=0.1, eps=1e-5, drop=None):
    """Train-mode BatchNorm2d (+ residual) (+ ReLU) (+ dropout) on NHWC bf16: y (..., C) -> (z bf16 like y, stat fp32 (C, 2) =
    mean | 1/std).  running_mean / running_var (fp32 (C,), both or neither) are updated in place like nn.BatchNorm2d does.
    drop = (seed tensor on the device, salt, p): element-wise dropout behind the ReLU, recomputed (not stored) by the backward."""
    _check_act(y)
    C = y.shape[-1]
    n_pix = y.numel() // C
    if residual is not None:
        _check_act(residual)
        if residual.shape != y.shape:
            raise ValueError('residual shape mismatch')
    for t in (gamma, beta):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == C):
            raise ValueError('gamma / beta must be contiguous CUDA float32 (C,)')
    z = torch.empty_like(y)
    stat = torch.empty((C, 2), dtype=torch.float32, device=y.device)
    sums = torch.empty((C, 2), dtype=torch.float64, device=y.device)
    _call('crnn_bn_train_forward', y, _p(y), _p(gamma), _p(beta), _p(residual), _p(z), _p(stat), _p(sums), _p(running_mean), _p(running_var),
          n_pix, C, ctypes.c_float(eps), ctypes.c_float(momentum), int(bool(relu)), *_drop_args(drop))
    return z, stat


def bn_train_backward(dz, z, y, stat, gamma, relu=True, want_residual_grad=False, beta=None, drop=None):
    """-> (dy bf16, d_residual bf16 or None, dgamma fp32 (C,), dbeta fp32 (C,)).  With `relu` and `beta` given and z None, the
    forward had no residual and the ReLU mask is recomputed from y (one tensor less to read)."""
    _check_act(dz)
    _check_act(y)
    C = y.shape[-1]
    n_pix = y.numel() // C
    dy = torch.empty_like(y)
    dres = torch.empty_like(y) if want_residual_grad else None
    sums = torch.empty((C, 2), dtype=torch.float64, device=y.device)
    dgamma = torch.empty((C,), dtype=torch.float32, device=y.device)
    dbeta = torch.empty((C,), dtype=torch.float32, device=y.device)
    mode = 0 if not relu else (2 if (z is None and beta is not None) else 1)
    _call('crnn_bn_train_backward', dz, _p(dz), _p(z), _p(y), _p(stat), _p(gamma), _p(beta), _p(dy), _p(dres), _p(sums), _p(dgamma), _p(dbeta),
          n_pix, C, mode, *_drop_args(drop))
    return dy, dres, dgamma, dbeta


def bn_train_forward_pool(y, gamma, beta, residual=None, running_mean=None, running_var=None, momentum=0.1, eps=1e-5):
    """BatchNorm (+ residual) + ReLU + 2x2 average pooling: y (B,H,W,C) bf16 -> (pooled (B,H//2,W//2,C) bf16, stat (C,2))."""
    _check_act(y)
    B, H, W, C = y.shape
    if residual is not None:
        _check_act(residual)
        if residual.shape != y.shape:
            raise ValueError('residual shape mismatch')
    pooled = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=y.device)
    stat = torch.empty((C, 2), dtype=torch.float32, device=y.device)
    sums = torch.empty((C, 2), dtype=torch.float64, device=y.device)
    _call('crnn_bn_train_forward_pool', y, _p(y), _p(gamma), _p(beta), _p(residual), _p(pooled), _p(stat), _p(sums), _p(running_mean),
          _p(running_var), B, H, W, C, ctypes.c_float(eps), ctypes.c_float(momentum))
    return pooled, stat


def bn_train_backward_pool(dpool, y, residual, stat, gamma, beta, want_residual_grad=False):
    """-> (dy, d_residual or None, dgamma, dbeta) from the gradient of the pooled output."""
    _check_act(dpool)
    _check_act(y)
    B, H, W, C = y.shape
    dy = torch.empty_like(y)
    dres = torch.empty_like(y) if want_residual_grad else None
    sums = torch.empty((C, 2), dtype=torch.float64, device=y.device)
    dgamma = torch.empty((C,), dtype=torch.float32, device=y.device)
    dbeta = torch.empty((C,), dtype=torch.float32, device=y.device)
    _call('crnn_bn_train_backward_pool', y, _p(dpool), _p(y), _p(residual), _p(stat), _p(gamma), _p(beta), _p(dy), _p(dres), _p(sums),
          _p(dgamma), _p(dbeta), B, H, W, C)
    return dy, dres, dgamma, dbeta


def gru_layer_train(xproj, w_hh, b_hh):
    """xproj (B, T, 1536) fp32, w_hh (2, 768, 256) fp32, b_hh (2, 768) fp32 -> (y (B, T, 512) fp32, save (B, T, 2, 4, 256) fp32)."""
    for t in (xproj, w_hh, b_hh):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError('gru_layer_train takes contiguous CUDA float32 tensors')
    B, T, n = xproj.shape
    if n != 1536 or tuple(w_hh.shape) != (2, 768, 256) or tuple(b_hh.shape) != (2, 768):
        raise ValueError('gru_layer_train: hidden size 256, both directions')
    y = torch.empty((B, T, 512), dtype=torch.float32, device=xproj.device)
    save = torch.empty((B, T, 2, 4, 256), dtype=torch.float32, device=xproj.device)
    _call('crnn_gru_layer_train', xproj, _p(xproj), _p(w_hh), _p(b_hh), _p(y), _p(save), B, T)
    return y, save


def gru_layer_backward(dy, y, save, w_hh):
    """dy, y (B, T, 512) fp32, save from gru_layer_train -> (dgi, dgh), both (B, T, 1536) fp32."""
    for t in (dy, y, save, w_hh):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError('gru_layer_backward takes contiguous CUDA float32 tensors')
    B, T, _ = y.shape
    dgi = torch.empty((B, T, 1536), dtype=torch.float32, device=y.device)
    dgh = torch.empty((B, T, 1536), dtype=torch.float32, device=y.device)
    _call('crnn_gru_layer_backward', dy, _p(dy), _p(y), _p(save), _p(w_hh), _p(dgi), _p(dgh), B, T)
    return dgi, dgh


def conv_first(x, w, bias=None, relu=True, planes=1):
    """First convolution: x (B,H,W,planes*16) bf16, w (9,64,planes*16) bf16 -> (B,H,W,planes*64) bf16."""
    _check_act(x)
    B, H, W, C = x.shape
    if C != 16 * planes or tuple(w.shape) != (9, 64, 16 * planes) or w.dtype != torch.bfloat16 or not w.is_contiguous():
        raise ValueError('conv_first expects 16 (padded) input channels and 64 output channels')
    out = torch.empty((B, H, W, planes * 64), dtype=torch.bfloat16, device=x.device)
    _call('crnn_conv_first', x, _p(x), _p(w), _p(bias), _p(out), B, H, W, int(bool(relu)), planes)
    return out


def pad_rows(n):
    """GEMM inputs are allocated with their row count rounded up to a multiple of 8."""
    return (n + 7) // 8 * 8


def gemm(a, w, bias=None, relu=False, M=None, out_f32=False, out=None, planes=1):
    """a (Mpad,planes*K) bf16 with Mpad % 8 == 0, w (N,planes*K) bf16 -> (Mpad,planes*N) bf16 or (Mpad,N) fp32;
    rows >= M are left untouched."""
    _check_act(a)
    Mpad, K = a.shape
    M = Mpad if M is None else M
    if Mpad % 8 != 0 or M > Mpad:
        raise ValueError('a must have a multiple of 8 rows (pad_rows)')
    N, K2 = w.shape
    if K2 != K or w.dtype != torch.bfloat16 or not w.is_contiguous():
        raise ValueError('w must be contiguous bf16 (N, K)')
    if out is None:
        out = (torch.zeros((Mpad, N), dtype=torch.float32, device=a.device) if out_f32 else
               torch.zeros((Mpad, planes * N), dtype=torch.bfloat16, device=a.device))
    o16, o32 = (None, out) if out.dtype == torch.float32 else (out, None)
    _call('crnn_gemm', a, _p(a), _p(w), _p(bias), _p(o16), _p(o32), M, N, K // planes, int(bool(relu)), planes)
    return out


def pack_input(x, t_use=None, c_pad=64, planes=1, scaler=None):
    """(B,C,T,F) fp32 NCHW -> (B,t_use,F,planes*c_pad) bf16 NHWC.  scaler = (mean, std) CUDA float32 tensors
    (n_scaled, 1, F) or (n_scaled, F): the first n_scaled channels become (x - mean) / std."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        raise ValueError('x must be a CUDA float32 tensor (B, C, T, F)')
    x = x.contiguous()
    B, C, T, F = x.shape
    t_use = T if t_use is None else t_use
    y = torch.empty((B, t_use, F, planes * c_pad), dtype=torch.bfloat16, device=x.device)
    mean = std = None
    n_scaled = 0
    if scaler is not None:
        mean = scaler[0].to(x.device, torch.float32).reshape(-1, F).contiguous()
        std = scaler[1].to(x.device, torch.float32).reshape(-1, F).contiguous()
        n_scaled = mean.shape[0]
        if std.shape != mean.shape or n_scaled > C:
            raise ValueError('scaler mean / std must be (n_scaled <= C, 1, F)')
    _call('crnn_pack_input', x, _p(x), _p(y), B, C, T, F, t_use, c_pad, planes, _p(mean), _p(std), n_scaled)
    return y


def avgpool2(x, planes=1):
    _check_act(x)
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, C), dtype=torch.bfloat16, device=x.device)
    _call('crnn_avgpool2', x, _p(x), _p(y), B, H, W, C // planes, planes)
    return y


def avgpool2_backward(dy, H, W):
    """dy (B,H//2,W//2,C) bf16 -> dx (B,H,W,C) bf16."""
    _check_act(dy)
    B, Ho, Wo, C = dy.shape
    if (Ho, Wo) != (H // 2, W // 2):
        raise ValueError('dy does not belong to an input of {} x {}'.format(H, W))
    dx = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dy.device)
    _call('crnn_avgpool2_backward', dy, _p(dy), _p(dx), B, H, W, C)
    return dx


def freq_mean(x, planes=1):
    """(B,H,W,C) -> (pad_rows(B*H), C), mean over W; padding rows are zero."""
    _check_act(x)
    B, H, W, C = x.shape
    y = torch.zeros((pad_rows(B * H), C), dtype=torch.bfloat16, device=x.device)
    _call('crnn_freq_mean', x, _p(x), _p(y), B * H, W, C // planes, planes)
    return y


def gru_layer(xproj, w_hh, b_hh, B, T, planes=1):
    """xproj (>=B*T, 1536) fp32, w_hh (2,768,256) fp32, b_hh (2,768) fp32 -> y (pad_rows(B*T), planes*512) bf16."""
    if xproj.dtype != torch.float32 or xproj.shape[1] != 1536 or not xproj.is_contiguous():
        raise ValueError('xproj must be contiguous fp32 (rows, 1536)')
    y = torch.zeros((pad_rows(B * T), planes * 512), dtype=torch.bfloat16, device=xproj.device)
    _call('crnn_gru_layer', xproj, _p(xproj), _p(w_hh), _p(b_hh), _p(y), B, T, planes)
    return y


def head_finish(z, rows, n_classes):
    logits = torch.empty((rows, n_classes), dtype=torch.float32, device=z.device)
    doa = torch.empty((rows, 3 * n_classes), dtype=torch.float32, device=z.device)
    _call('crnn_head_finish', z, _p(z), _p(logits), _p(doa), rows, n_classes)
    return logits, doa


def interpolate_index(n_in, ratio):
    """Index map of interpolate_tensor (models/model_utils.py:66-70): floor(arange(n_out) / ratio) with the
    division carried out in float32, as torch does for an int64 tensor divided by a Python float."""
    ratio = float(ratio)
    n_out = int(round(n_in * ratio))
    idx = np.floor(np.arange(n_out).astype(np.float32) / np.float32(ratio)).astype(np.int64)
    return idx


def gather_time(x, idx):
    """x (B,n_in,width) fp32 CUDA, idx int array -> (B,len(idx),width)."""
    x = x.contiguous()
    B, n_in, width = x.shape
    if len(idx) and (idx.min() < 0 or idx.max() >= n_in):
        raise IndexError('interpolation index out of range')
    d_idx = torch.from_numpy(np.asarray(idx, dtype=np.int32)).to(x.device)
    out = torch.empty((B, len(idx), width), dtype=torch.float32, device=x.device)
    _call('crnn_gather_time', x, _p(x), _p(d_idx), _p(out), B, n_in, len(idx), width)
    return out


def decode_events(logits, doa, threshold=0.3):
    """logits (rows, n) fp32 CUDA, doa (rows, 3n) fp32 CUDA -> active (rows, n) bool, azi, ele (rows, n) int16
    (write_classwise_output_to_file, models/interfaces.py:224-246)."""
    logits, doa = logits.contiguous(), doa.contiguous()
    rows, n = logits.shape
    if tuple(doa.shape) != (rows, 3 * n):
        raise ValueError('doa must be (rows, 3 * n_classes)')
    active = torch.empty((rows, n), dtype=torch.uint8, device=logits.device)
    azi = torch.empty((rows, n), dtype=torch.int16, device=logits.device)
    ele = torch.empty((rows, n), dtype=torch.int16, device=logits.device)
    _call('crnn_decode_events', logits, _p(logits), _p(doa), rows, n, ctypes.c_float(threshold), _p(active), _p(azi),
                                                   _p(ele))
    return active.bool(), azi, ele


def seld_loss(event_logit, doa_output, event_gt, doa_gt, loss_weight=(0.3, 0.7), with_grad: bool = False):
    """BaseModel.compute_loss for output_format='reg_xyz' (models/interfaces.py:273-355) on CUDA tensors:
    event_logit / event_gt (B, T, n), doa_output / doa_gt (B, T', 3 n) -> float32 tensor (3,) = (loss, sed_loss, doa_loss)
    and, with_grad, the gradients of `loss` with respect to event_logit and doa_output."""
    if not (event_logit.is_cuda and doa_output.is_cuda and event_gt.is_cuda and doa_gt.is_cuda):
        raise ValueError('seld_loss: CUDA tensors expected (there is no CPU fallback)')
    B, T, n = event_logit.shape
    if tuple(event_gt.shape) != (B, T, n) or doa_output.shape[2] != 3 * n or doa_gt.shape[2] != 3 * n:
        raise ValueError('seld_loss: inconsistent shapes')
    N = min(doa_output.shape[1], doa_gt.shape[1])          # compute_masked_reg_loss aligns the time axes (:337-341)
    if N != T:
        raise ValueError('seld_loss: the event and DOA outputs must share their time axis (label rate)')
    f = lambda t: t[:, :N].contiguous().float()
    logit, doa, egt, dgt = f(event_logit), f(doa_output), f(event_gt), f(doa_gt)
    sums = torch.empty(5, dtype=torch.float64, device=logit.device)
    loss = torch.empty(3, dtype=torch.float32, device=logit.device)
    g_logit = torch.empty_like(logit) if with_grad else None
    g_doa = torch.empty_like(doa) if with_grad else None
    _call('crnn_seld_loss', logit, _p(logit), _p(doa), _p(egt), _p(dgt), B * N, n, ctypes.c_float(loss_weight[0]),
                                               ctypes.c_float(loss_weight[1]), _p(sums), _p(loss), _p(g_logit), _p(g_doa))
    return (loss, g_logit, g_doa) if with_grad else loss
