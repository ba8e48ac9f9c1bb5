"""CRNN operators (C ABI of include/salsa_crnn.h) against plain PyTorch fp32 references of the same op.

Inputs and weights are rounded to bf16 first, so the only differences are the accumulation order
(fp32 in both) and the bf16 rounding of the output: tolerance 1e-2 relative to the output scale for
bf16 outputs, 2e-3 for fp32 outputs.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from salsa_b200 import crnn_ops
    assert torch.cuda.is_available()
    return crnn_ops


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-6))


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (300, 1536, 512), (1000, 256, 1024), (8, 64, 128), (2400, 128, 64)])
def test_gemm(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g)
    ref = a.float() @ w.float().T + bias
    a_pad = torch.zeros(ops.pad_rows(M), K, dtype=torch.bfloat16)
    a_pad[:M] = a
    out = ops.gemm(a_pad.cuda(), w.cuda(), bias.cuda(), M=M, out_f32=True)[:M].cpu()
    assert rel_err(out, ref) < 2e-3
    out16 = ops.gemm(a_pad.cuda(), w.cuda(), bias.cuda(), relu=True, M=M)[:M].cpu().float()
    assert rel_err(out16, ref.clamp(min=0)) < 1e-2


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', [
    (1, 16, 8, 64, 64, 3), (2, 40, 25, 64, 64, 3), (1, 33, 12, 128, 256, 3), (1, 20, 11, 256, 512, 3),
    (2, 37, 50, 64, 128, 1), (1, 16, 200, 64, 64, 3), (1, 18, 6, 512, 512, 3)])
def test_conv2d(ops, B, H, W, Cin, Cout, k):
    g = torch.Generator().manual_seed(H * W + Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).bfloat16()
    bias = torch.randn(Cout, generator=g)
    res = torch.randn(B, H, W, Cout, generator=g).bfloat16()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=k // 2) + bias[None, :, None, None]
    ref = ref.permute(0, 2, 3, 1)
    wp = w.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin).contiguous()
    out = ops.conv2d(x.cuda(), wp.cuda(), bias.cuda(), out_f32=True).cpu()
    assert rel_err(out, ref) < 2e-3
    out2 = ops.conv2d(x.cuda(), wp.cuda(), bias.cuda(), residual=res.cuda(), relu=True).cpu().float()
    assert rel_err(out2, (ref + res.float()).clamp(min=0)) < 1e-2


def test_pack_pool_mean(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 7, 33, 21, generator=g)
    y = ops.pack_input(x.cuda(), t_use=32).cpu().float()
    assert y.shape == (2, 32, 21, 64)
    assert torch.equal(y[..., :7], x[:, :, :32].permute(0, 2, 3, 1).bfloat16().float())
    assert torch.all(y[..., 7:] == 0)
    a = torch.randn(2, 9, 25, 64, generator=g).bfloat16()
    p = ops.avgpool2(a.cuda()).cpu().float()
    ref = F.avg_pool2d(a.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert p.shape == (2, 4, 12, 64) and rel_err(p, ref) < 1e-2
    m = ops.freq_mean(a.cuda()).cpu().float()
    assert m.shape == (ops.pad_rows(18), 64)
    assert rel_err(m[:18], a.float().mean(dim=2).reshape(18, 64)) < 1e-2
    assert torch.all(m[18:] == 0)


@pytest.mark.parametrize('B,T', [(1, 40), (3, 75), (9, 20)])
def test_gru_layer(ops, B, T):
    torch.manual_seed(B * 100 + T)
    gru = torch.nn.GRU(input_size=512, hidden_size=256, num_layers=1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, 512)
    with torch.no_grad():
        ref, _ = gru(x)
        xproj = torch.cat([x @ gru.weight_ih_l0.T + gru.bias_ih_l0, x @ gru.weight_ih_l0_reverse.T + gru.bias_ih_l0_reverse], dim=-1)
    w_hh = torch.stack([gru.weight_hh_l0, gru.weight_hh_l0_reverse]).detach().contiguous()
    b_hh = torch.stack([gru.bias_hh_l0, gru.bias_hh_l0_reverse]).detach().contiguous()
    y = ops.gru_layer(xproj.reshape(B * T, 1536).contiguous().cuda(), w_hh.cuda(), b_hh.cuda(), B, T)
    y = y[:B * T].cpu().float().reshape(B, T, 512)
    assert (y - ref).abs().max() < 1e-2           # bf16 output rounding of values in (-1, 1)


def test_interpolate_index_and_gather(ops, golden):
    assert np.array_equal(ops.interpolate_index(40, 2.0), np.repeat(np.arange(40), 2))
    assert np.array_equal(ops.interpolate_index(8, 0.5), np.arange(0, 8, 2))
    x = torch.arange(24, dtype=torch.float32).reshape(2, 4, 3)
    for ratio in (0.5, 2.0, 1.5):
        idx = ops.interpolate_index(4, ratio)
        n_out = int(round(4 * ratio))
        ref = x[:, torch.floor(torch.arange(n_out) / ratio).long()]
        out = ops.gather_time(x.cuda(), idx).cpu()
        assert torch.equal(out, ref)


def test_head_finish(ops):
    z = torch.randn(10, 64)
    logits, doa = ops.head_finish(z.cuda(), 10, 12)
    assert torch.equal(logits.cpu(), z[:, :12])
    assert torch.allclose(doa.cpu(), torch.tanh(z[:, 12:48]), atol=1e-6)


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', [(1, 20, 9, 64, 64, 3), (1, 17, 12, 128, 256, 3), (2, 16, 8, 64, 128, 1)])
def test_conv2d_bf16x3_is_float32_grade(ops, B, H, W, Cin, Cout, k):
    """planes = 3: operands are NOT rounded to bf16 first; the result must match an fp64 reference of the
    float32 inputs to float32 accuracy."""
    g = torch.Generator().manual_seed(H * W + Cin + Cout + 1)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    res = torch.randn(B, H, W, Cout, generator=g)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=k // 2) + bias.double()[None, :, None, None]
    ref = ref.permute(0, 2, 3, 1)
    wp = ops.split_planes(w.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin), 3)
    xs, rs = ops.split_planes(x, 3), ops.split_planes(res, 3)
    assert rel_err(ops.merge_planes(xs, 3), x) < 1e-7
    out = ops.conv2d(xs.cuda(), wp.cuda(), bias.cuda(), out_f32=True, planes=3).cpu()
    assert rel_err(out, ref) < 1e-5          # tensor-core fp32 accumulation: a few 1e-6
    out2 = ops.conv2d(xs.cuda(), wp.cuda(), bias.cuda(), residual=rs.cuda(), relu=True, planes=3).cpu()
    assert rel_err(ops.merge_planes(out2, 3), (ref + res.double()).clamp(min=0)) < 1e-5


def test_gemm_pool_mean_gru_bf16x3(ops):
    g = torch.Generator().manual_seed(77)
    a = torch.randn(40, 128, generator=g)
    w = torch.randn(192, 128, generator=g) / 128 ** 0.5
    out = ops.gemm(ops.split_planes(a, 3).cuda(), ops.split_planes(w, 3).cuda(), None, out_f32=True, planes=3).cpu()
    assert rel_err(out, a.double() @ w.double().T) < 1e-5
    x = torch.randn(1, 8, 10, 64, generator=g)
    p = ops.merge_planes(ops.avgpool2(ops.split_planes(x, 3).cuda(), planes=3).cpu(), 3)
    assert rel_err(p, F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)) < 1e-6
    m = ops.merge_planes(ops.freq_mean(ops.split_planes(x, 3).cuda(), planes=3).cpu(), 3)
    assert rel_err(m[:8], x.mean(dim=2).reshape(8, 64)) < 1e-6


@pytest.mark.parametrize('B,H,W,planes', [(1, 16, 8, 1), (2, 37, 21, 1), (1, 40, 200, 1), (1, 20, 13, 3)])
def test_conv_first(ops, B, H, W, planes):
    g = torch.Generator().manual_seed(B + H + W)
    x = torch.randn(B, 7, H, W, generator=g)
    w = torch.randn(64, 7, 3, 3, generator=g) / 63 ** 0.5
    bias = torch.randn(64, generator=g)
    if planes == 1:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    ref = torch.relu(F.conv2d(x.double(), w.double(), padding=1) + bias.double()[None, :, None, None]).permute(0, 2, 3, 1)
    xp = ops.pack_input(x.cuda(), c_pad=16, planes=planes)
    assert tuple(xp.shape) == (B, H, W, 16 * planes)
    wp = torch.zeros(9, 64, 16)
    wp[:, :, :7] = w.permute(2, 3, 0, 1).reshape(9, 64, 7)
    out = ops.conv_first(xp, ops.split_planes(wp, planes).cuda(), bias.cuda(), relu=True, planes=planes).cpu()
    assert rel_err(ops.merge_planes(out, planes), ref) < (1e-2 if planes == 1 else 1e-5)


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(1, 32, 16, 64, 64), (2, 37, 25, 64, 64), (1, 33, 12, 128, 256), (1, 50, 100, 64, 128)])
def test_conv2d_fused_pool_equals_separate_pool(ops, B, H, W, Cin, Cout):
    """pool=True (epilogue-fused F.avg_pool2d) must be bit-identical to conv2d followed by avgpool2, odd sizes included."""
    g = torch.Generator().manual_seed(H + W + Cin)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16().cuda()
    w = (torch.randn(9, Cout, Cin, generator=g) / (Cin * 9) ** 0.5).bfloat16().cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).bfloat16().cuda()
    sep = ops.avgpool2(ops.conv2d(x, w, bias, residual=res, relu=True))
    fused = ops.conv2d(x, w, bias, residual=res, relu=True, pool=True)
    assert tuple(fused.shape) == (B, H // 2, W // 2, Cout)
    assert torch.equal(fused.view(torch.int16), sep.view(torch.int16))


def test_seld_loss_matches_reference_golden_and_autograd(golden):
    """crnn_seld_loss against (a) the value the unmodified reference BaseModel.compute_loss returned for the same inputs
    (tests/golden/model_cases.npz: loss_values), (b) the oracle restatement and its autograd gradients."""
    from oracle import crnn as ocrnn
    from salsa_b200 import crnn_ops as ops
    logit, doa, egt, dgt = ocrnn.seld_loss_inputs()
    ref = golden('model_cases')['loss_values']
    loss, g_logit, g_doa = ops.seld_loss(logit.cuda(), doa.cuda(), egt.cuda(), dgt.cuda(), with_grad=True)
    np.testing.assert_allclose(loss.cpu().numpy().astype(np.float64), ref, rtol=2e-6, atol=0)
    lt, dt = logit.clone().requires_grad_(True), doa.clone().requires_grad_(True)
    ocrnn.seld_loss(lt, dt, egt, dgt)[0].backward()
    np.testing.assert_allclose(g_logit.cpu().numpy(), lt.grad.numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(g_doa.cpu().numpy(), dt.grad.numpy(), rtol=1e-5, atol=1e-9)
    # other weights, a larger batch, no active cell at all (0 / 0 = NaN like the reference)
    big = ocrnn.seld_loss_inputs(seed=9, batch=32, n_frames=600)
    got = ops.seld_loss(*[t.cuda() for t in big], loss_weight=(0.5, 0.5)).cpu().numpy()
    want = np.array([float(v) for v in ocrnn.seld_loss(*big, loss_weight=(0.5, 0.5))])
    np.testing.assert_allclose(got, want, rtol=5e-6)
    none = ops.seld_loss(logit.cuda(), doa.cuda(), torch.zeros_like(egt).cuda(), torch.zeros_like(dgt).cuda()).cpu().numpy()
    assert np.isfinite(none[1]) and np.isnan(none[2]) and np.isnan(none[0])


def test_adam_and_schedule_match_torch():
    """salsa_b200.optim.Adam + LearningRateScheduler against torch.optim.Adam driven by the schedule arithmetic of
    utilities/learning_utils.py:39-52 (np.interp over the step milestones), over steps that cross a milestone."""
    from salsa_b200 import optim
    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(100_003, generator=g)
    sched = optim.LearningRateScheduler(steps_per_epoch=4, max_epochs=5)
    assert sched.step_milestones == [0, 9, 18, 20]
    ref_p = torch.nn.Parameter(p0.clone().cuda())
    ref = torch.optim.Adam([ref_p], lr=1e-3)
    mine_p = p0.clone().cuda()
    mine = optim.Adam(mine_p, lr=1e-3)
    for step in range(12):
        epoch, batch_idx = divmod(step, 4)
        grad = torch.randn(p0.shape, generator=g).cuda() * (1.0 + step)
        lr, mom = sched.at(epoch, batch_idx)
        assert lr == float(np.interp(step, [0, 9, 18, 20], (1e-4, 1e-2, 1e-3, 1e-4)))
        for group in ref.param_groups:
            group['lr'], group['betas'] = lr, (mom, 0.999)
        ref_p.grad = grad.clone()
        ref.step()
        sched.apply(mine, epoch, batch_idx)
        mine.step(grad)
    err = (mine_p - ref_p.detach()).abs().max().item()
    assert err <= 2e-7 * p0.abs().max().item(), err
    print('adam: max |param diff|', err)
    st = ref.state[ref_p]
    # the moments agree to float32 rounding of their own scale (torch's fused lerp may round the last bit differently)
    for mine_t, ref_t in ((mine.exp_avg, st['exp_avg']), (mine.exp_avg_sq, st['exp_avg_sq'])):
        d = (mine_t - ref_t).abs().max().item()
        assert d <= 1e-6 * ref_t.abs().max().item(), (d, ref_t.abs().max().item())
