"""The on-the-fly training step (salsa_b200.train; BASELINE.json configs[4]): the native convolution
forward / input gradient inside autograd, the whole step against a pure torch float32 step, and the hand-over of trained
weights to the native inference model."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-12)


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 40, 25, 64, 64), (1, 33, 12, 128, 256)])
def test_native_conv_forward_and_input_gradient(B, H, W, Cin, Cout):
    """forward and dgrad run on the tcgen05 kernel (dgrad = the same kernel on flipped / transposed weights); reference =
    float32 F.conv2d + autograd on the bf16-rounded operands.  The weight gradient comes from cuDNN."""
    from salsa_b200.train import NativeConv3x3
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float().cuda().requires_grad_(True)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).bfloat16().float().cuda().requires_grad_(True)
    gy = torch.randn(B, Cout, H, W, generator=g).bfloat16().float().cuda()
    y_ref = F.conv2d(x, w, padding=1)
    gx_ref, gw_ref = torch.autograd.grad(y_ref, (x, w), gy)
    y = NativeConv3x3.apply(x, w)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (B, Cout, H, W)
    gx, gw = torch.autograd.grad(y, (x, w), gy)
    assert rel(y, y_ref) < 1e-2 and rel(gx, gx_ref) < 1e-2 and rel(gw, gw_ref) < 2e-2


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(1, 16, 8, 64, 64), (2, 40, 25, 64, 64), (1, 33, 12, 128, 256), (3, 80, 50, 64, 128)])
def test_native_conv_weight_gradient(B, H, W, Cin, Cout):
    """crnn_conv_wgrad (tcgen05 on MN-major operands, split over the pixel tiles, fp32 atomics) against the float32 weight
    gradient of the same bf16 operands; ragged image borders exercise the zero fill of the TMA boxes."""
    import salsa_b200
    g = torch.Generator().manual_seed(B + H + Cin)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16().cuda()
    gy = torch.randn(B, H, W, Cout, generator=g).bfloat16().cuda()
    dw = salsa_b200.crnn_ops.conv_wgrad(x, gy)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3), gy.float().permute(0, 3, 1, 2), padding=1)
    ref = ref.permute(2, 3, 0, 1).reshape(9, Cout, Cin)
    assert rel(dw, ref) < 1e-4


def _batch(seed=3, B=2, T=128):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 7, T, 200, generator=g).cuda()
    tgt = {'event_frame_gt': (torch.rand(B, T // 8, 12, generator=g) > 0.6).float().cuda(),
           'doa_frame_gt': torch.randn(B, T // 8, 36, generator=g).clamp(-1, 1).cuda()}
    return x, tgt


def test_training_step_matches_pure_torch_steps():
    """Same weights, same batch, dropout off.  Against the same step with cuDNN convolutions under bf16 autocast (the same
    arithmetic class): gradient cosine > 0.9999 -- the native forward / dgrad are drop-ins.  Against the float32 step
    (autocast off): the loss to 1e-3; the gradient direction is only as close as bf16 leaves it with a sign-gradient (MAE)
    loss and batch statistics over two samples (measured 0.96, the same for the cuDNN bf16 step): printed, loose bar."""
    import salsa_b200
    from salsa_b200 import train
    sd = salsa_b200.crnn.random_state_dict(1)
    x, tgt = _batch()

    def make(**kw):
        t = train.SeldTrainer(sd, dropout=False, **kw)
        t.optimizer = None                                    # gradients only
        return t

    nat, t16, t32 = make(), make(native_conv=False), make(native_conv=False, autocast=False)
    l_nat, l_16, l_32 = nat.step(x, tgt), t16.step(x, tgt), t32.step(x, tgt)
    assert torch.allclose(l_nat, l_16, rtol=1e-3, atol=1e-4), (l_nat, l_16)
    assert torch.allclose(l_nat, l_32, rtol=2e-3, atol=2e-3), (l_nat, l_32)
    cos16 = F.cosine_similarity(nat.flat_grad, t16.flat_grad, dim=0).item()
    cos32 = F.cosine_similarity(nat.flat_grad, t32.flat_grad, dim=0).item()
    ref32 = F.cosine_similarity(t16.flat_grad, t32.flat_grad, dim=0).item()
    print('train step: loss {}; gradient cosine native vs cuDNN-bf16 {:.6f}, native vs float32 {:.4f} (cuDNN-bf16 vs float32 {:.4f})'.format(
        l_nat.tolist(), cos16, cos32, ref32))
    assert cos16 > 0.9999
    assert cos32 > 0.9 and abs(cos32 - ref32) < 5e-3
    out = t32.forward(x)
    assert tuple(out['event_frame_logit'].shape) == (2, 8, 12) and tuple(out['doa_frame_output'].shape) == (2, 8, 36)


def test_steps_reduce_the_loss_and_weights_hand_over_to_inference():
    import salsa_b200
    from salsa_b200 import train
    torch.manual_seed(0)
    sd = salsa_b200.crnn.random_state_dict(2)
    x, tgt = _batch(seed=5)
    # dropout off: the run is then deterministic up to the summation order of the weight gradient's atomics
    tr = train.SeldTrainer(sd, lr=1e-3, dropout=False, scheduler=salsa_b200.optim.LearningRateScheduler(steps_per_epoch=10, max_epochs=2))
    losses = [tr.step(x, tgt)[0].item() for _ in range(15)]
    print('losses', ['{:.4f}'.format(v) for v in losses])
    assert min(losses[-3:]) < 0.85 * losses[0] and np.isfinite(losses).all()
    # trained weights -> native inference model; trainer in eval mode (running statistics) is the reference forward
    tr.training = False
    with torch.no_grad():
        want = tr.forward(x)
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256))
    m.load_state_dict(tr.state_dict())
    got = m(x)
    for k in want:
        print('hand-over to the inference model:', k, 'rel-to-scale difference {:.3e}'.format(rel(got[k], want[k])))
        assert rel(got[k], want[k]) < 1e-1, k
