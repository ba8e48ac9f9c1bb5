"""SELD CRNN forward through the C ABI against the oracle (plain PyTorch fp32 restatement pinned to the
reference modules) and the golden outputs of the reference itself.

Arithmetic is bf16 operands with fp32 accumulation, so the 1e-4 target of the float32 reference cannot
hold for the logits; the tolerances below are the measured bf16 envelope (relative to the output
scale): encoder activations 3e-2, logits / DOA 3e-2.  The index map of interpolate_tensor is exact.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def scale_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-6)


@pytest.fixture(scope='module')
def model():
    import salsa_b200
    from oracle import crnn as ocrnn
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                             salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                    freq_pool='avg', decoder_size=256),
                             label_rate=10, feature_rate=80.0)
    m.load_state_dict(ocrnn.make_state_dict(0))
    return m.eval()


def test_forward_matches_golden(model, golden):
    from oracle import crnn as ocrnn
    g = golden('model_cases')
    x = ocrnn.model_input(2, (2, 7, 128, 200))
    enc = model.encode(x.cuda()).float().cpu().permute(0, 3, 1, 2).numpy()
    assert enc.shape == (2, 512, 8, 12)
    assert scale_err(enc, g['model_encoder_out'].astype(np.float32)) < 3e-2
    y = model(x.cuda())
    assert tuple(y['event_frame_logit'].shape) == (2, 8, 12) and tuple(y['doa_frame_output'].shape) == (2, 8, 36)
    assert y['event_frame_logit'].dtype == torch.float32
    e1 = scale_err(y['event_frame_logit'].cpu().numpy(), g['model_event_frame_logit'])
    e2 = scale_err(y['doa_frame_output'].cpu().numpy(), g['model_doa_frame_output'])
    print('bf16 CRNN vs reference: logits {:.3e}, doa {:.3e} (relative to output scale)'.format(e1, e2))
    assert e1 < 3e-2 and e2 < 3e-2


def test_forward_bf16x3_meets_float32_tolerance(golden):
    """The parity mode: three bf16 planes per operand, six plane products per MAC on the same tcgen05 kernels.
    Bar = the float32 target of BASELINE.json: 1e-4 relative (to the output scale) on logits and DOA."""
    import salsa_b200
    from oracle import crnn as ocrnn
    g = golden('model_cases')
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256),
                             precision='bf16x3')
    m.load_state_dict(ocrnn.make_state_dict(0))
    x = ocrnn.model_input(2, (2, 7, 128, 200))
    enc = salsa_b200.crnn_ops.merge_planes(m.encode(x.cuda()).cpu(), 3).permute(0, 3, 1, 2).numpy()
    assert scale_err(enc, g['model_encoder_out'].astype(np.float32)) < 1e-3            # golden copy is float16
    y = m(x.cuda())
    e1 = scale_err(y['event_frame_logit'].cpu().numpy(), g['model_event_frame_logit'])
    e2 = scale_err(y['doa_frame_output'].cpu().numpy(), g['model_doa_frame_output'])
    print('bf16x3 CRNN vs reference: logits {:.3e}, doa {:.3e} (relative to output scale)'.format(e1, e2))
    assert e1 < 1e-4 and e2 < 1e-4
    y2 = m(ocrnn.model_input(3, (1, 7, 96, 191)).cuda())
    assert scale_err(y2['event_frame_logit'].cpu().numpy(), g['lite_event_frame_logit']) < 1e-4
    assert scale_err(y2['doa_frame_output'].cpu().numpy(), g['lite_doa_frame_output']) < 1e-4


def test_forward_lite_shape_matches_golden(model, golden):
    from oracle import crnn as ocrnn
    g = golden('model_cases')
    y = model(ocrnn.model_input(3, (1, 7, 96, 191)).cuda())
    assert tuple(y['event_frame_logit'].shape) == (1, 6, 12)
    assert scale_err(y['event_frame_logit'].cpu().numpy(), g['lite_event_frame_logit']) < 3e-2
    assert scale_err(y['doa_frame_output'].cpu().numpy(), g['lite_doa_frame_output']) < 3e-2


def test_forward_matches_oracle_other_weights_and_batch_invariance():
    import salsa_b200
    from oracle import crnn as ocrnn
    sd = ocrnn.make_state_dict(3)
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256))
    m.load_state_dict(sd)
    x = ocrnn.model_input(11, (3, 7, 80, 200))
    ref = ocrnn.forward(sd, x)
    y = m(x.cuda())
    for k in ref:
        assert scale_err(y[k].cpu().numpy(), ref[k].numpy()) < 3e-2
    # clips are independent: one-by-one gives the same rows (bit-identical)
    for i in range(3):
        yi = m(x[i:i + 1].cuda())
        for k in ref:
            assert torch.equal(yi[k][0], y[k][i])


def test_predict_interpolates_like_reference(model):
    from oracle import crnn as ocrnn
    x = ocrnn.model_input(5, (1, 7, 64, 200))
    y = model(x.cuda())
    p = model.predict(x.cuda())
    for k in y:
        ref = ocrnn.interpolate_tensor(y[k].cpu(), 16 * 10 / 80.0)
        assert torch.equal(p[k].cpu(), ref)


def test_interface_errors():
    import salsa_b200
    with pytest.raises(AssertionError):
        salsa_b200.SeldDecoder(512, decoder_type='foo', freq_pool='avg', decoder_size=256)
    with pytest.raises(NotImplementedError):
        salsa_b200.SeldDecoder(512, decoder_type='transformer', freq_pool='avg', decoder_size=256)
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 7, 32, 200).cuda())
    with pytest.raises(RuntimeError):
        m.train()                                   # no weights loaded yet


def test_fused_scaler_normalisation(model):
    """(x - mean) / std of channels 0..3 fused into the input packing == normalising on the host first (database.py:196-202)."""
    from oracle import crnn as ocrnn
    g = torch.Generator().manual_seed(9)
    x = ocrnn.model_input(6, (1, 7, 64, 200))
    x[:, :4] = x[:, :4] * 12.0 - 55.0                      # dB-like spectrogram channels
    mean = torch.empty(4, 1, 200).uniform_(-60, -50, generator=g)
    std = torch.empty(4, 1, 200).uniform_(8, 15, generator=g)
    xn = x.clone()
    xn[:, :4] = (xn[:, :4] - mean) / std
    ref = model(xn.cuda())
    model.set_scaler(mean.numpy(), std.numpy())
    try:
        out = model(x.cuda())
    finally:
        model.set_scaler(None, None)
    for k in ref:
        assert torch.equal(out[k], ref[k])


def test_event_decoding_matches_oracle(model):
    """sigmoid threshold, arctan2 -> whole degrees, csv rows (interfaces.py:210-258): integer outputs must be identical."""
    from oracle import crnn as ocrnn
    x = ocrnn.model_input(8, (2, 7, 160, 200))
    pred = model.predict(x.cuda())
    rows = model.events(x.cuda(), sed_threshold=0.3)
    assert len(rows) == 2
    for b in range(2):
        ref = ocrnn.decode_events(pred['event_frame_logit'][b:b + 1].cpu(), pred['doa_frame_output'][b:b + 1].cpu(), 0.3)
        assert rows[b] == ref
        assert len(ref) > 10
    # the decoding kernel alone on a dense grid of directions, including the +-180 and +-90 degree edges
    import salsa_b200
    g = torch.Generator().manual_seed(3)
    doa = torch.randn(4000, 36, generator=g)
    doa[:50, 12:24] = 0.0                      # y = 0: azimuth 0 or 180 -> -180
    doa[:25, :12] = -doa[:25, :12].abs()
    logits = torch.randn(4000, 12, generator=g) * 3
    active, azi, ele = salsa_b200.crnn_ops.decode_events(logits.cuda(), doa.cuda(), 0.3)
    xs, ys, zs = doa[:, :12].numpy(), doa[:, 12:24].numpy(), doa[:, 24:].numpy()
    ref_azi = np.around(np.arctan2(ys, xs) * 180.0 / np.pi).astype(int)
    ref_azi[ref_azi == 180] = -180
    ref_ele = np.around(np.arctan2(zs, np.sqrt(xs ** 2 + ys ** 2)) * 180.0 / np.pi).astype(int)
    assert np.array_equal(azi.cpu().numpy().astype(int), ref_azi)
    assert np.array_equal(ele.cpu().numpy().astype(int), ref_ele)
    assert np.array_equal(active.cpu().numpy(), (torch.sigmoid(logits).numpy() >= 0.3))


def test_pipeline_audio_to_outputs():
    """SeldPipeline = SalsaExtractor -> (scaler fused into the input packing) -> SeldModel on device memory: identical to
    running the three stages by hand, and the 4801st frame is dropped like Database does (database.py:205-207)."""
    import salsa_b200
    from oracle import crnn as ocrnn, synth
    audio = torch.from_numpy(np.stack([synth.make_clip(40 + i, 'foa', seconds=3.2) for i in range(2)])).cuda()
    ex = salsa_b200.SalsaExtractor('foa')
    model = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                                 salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                        freq_pool='avg', decoder_size=256), label_rate=10, feature_rate=80.0)
    model.load_state_dict(ocrnn.make_state_dict(0))
    feat = ex.extract(audio)
    mean, std = salsa_b200.compute_scaler([feat])
    pipe = salsa_b200.SeldPipeline(ex, model, scaler=(mean, std))
    T = feat.shape[2]
    assert T == 257 and pipe._n_frames(T) == 256
    out = pipe(audio)
    want = model.forward(feat, n_frames=256)
    for k in want:
        assert tuple(out[k].shape) == (2, 16, 12 if k == 'event_frame_logit' else 36)
        assert torch.equal(out[k], want[k])
    pred = pipe.predict(audio)
    assert tuple(pred['event_frame_logit'].shape) == (2, 32, 12)          # 16 * 10 / 80 -> ratio 2
    rows = pipe.events(audio)
    assert len(rows) == 2 and all(len(r) == 5 for clip in rows for r in clip)
    # the wav files' own 16-bit samples as input: converted on the device, same outputs as their float32 form
    pcm = torch.round(audio * 32768.0).clamp(-32768, 32767).to(torch.int16)
    out16, outf = pipe(pcm), pipe(pcm.float() / 32768.0)
    for k in outf:
        assert torch.equal(out16[k], outf[k])


@pytest.mark.parametrize('precision', ['bf16', 'bf16x2', 'bf16x3'])
def test_one_call_forward_equals_operator_by_operator(precision):
    """`forward` = crnn_forward (the library runs the whole layer schedule on one workspace, as a CUDA graph from the second
    call on) must be bit-identical to `forward_ops` (every layer its own call on torch-allocated tensors), which also pins
    the library's own BatchNorm folding / weight packing (crnn_load_weights) to the Python one."""
    import salsa_b200
    from oracle import crnn as ocrnn
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256),
                             precision=precision)
    m.load_state_dict(ocrnn.make_state_dict(5))
    for shape in ((2, 7, 96, 200), (1, 7, 81, 191)):            # 81 frames: odd sizes through the four floor-mode poolings
        x = ocrnn.model_input(13, shape).cuda()
        n_frames = 80 if shape[2] == 81 else None                # the data layer's trim (database.py:205-207)
        want = m.forward_ops(x, n_frames=n_frames)
        outs = [m.forward(x, n_frames=n_frames) for _ in range(3)]       # direct, captured, replayed
        for got in outs:
            for k in want:
                assert torch.equal(got[k], want[k]), (precision, shape, k)
    mean = torch.full((4, 1, 200), -50.0)
    std = torch.full((4, 1, 200), 12.0)
    x = ocrnn.model_input(14, (1, 7, 64, 200)).cuda()
    m.set_scaler(mean.numpy(), std.numpy())
    try:
        a, b = m.forward(x), m.forward_ops(x)
    finally:
        m.set_scaler(None, None)
    for k in a:
        assert torch.equal(a[k], b[k])


def test_c_abi_model_entry_points_with_caller_owned_workspace():
    """crnn_load_weights / crnn_workspace_bytes / crnn_forward through ctypes alone, as a non-Python host would call them
    (include/salsa_crnn.h; torch only lends device memory): nothing is allocated inside the forward, a too small workspace
    and a missing tensor are reported, results equal the Python-side model."""
    import ctypes
    import salsa_b200
    from salsa_b200 import _native
    from oracle import crnn as ocrnn
    lib = _native.lib()
    sd = ocrnn.make_state_dict(0)
    arrays = {k: np.ascontiguousarray(v.detach().float().numpy()) for k, v in sd.items() if not k.endswith('num_batches_tracked')}

    def table(names):
        entries = [_native.CrnnTensor(k.encode(), arrays[k].ctypes.data_as(ctypes.POINTER(ctypes.c_float)), arrays[k].size) for k in names]
        return (_native.CrnnTensor * len(entries))(*entries), len(entries)

    t, n = table([k for k in arrays if k != 'decoder.gru.weight_hh_l1'])
    handle = ctypes.c_void_p()
    assert lib.crnn_load_weights(ctypes.cast(t, ctypes.c_void_p), n, 1, 12, ctypes.byref(handle)) == _native.SALSA_EINVAL
    assert b'decoder.gru.weight_hh_l1' in lib.salsa_last_error()
    t, n = table(list(arrays))
    _native.check(lib.crnn_load_weights(ctypes.cast(t, ctypes.c_void_p), n, 1, 12, ctypes.byref(handle)))
    B, T, F = 2, 128, 200
    x = ocrnn.model_input(2, (B, 7, T, F)).cuda()
    need = lib.crnn_workspace_bytes(handle, B, T, F)
    assert need > 0
    work = torch.empty(need, dtype=torch.uint8, device='cuda')
    logits = torch.empty((B, T // 16, 12), device='cuda')
    doa = torch.empty((B, T // 16, 36), device='cuda')
    p = lambda a: ctypes.c_void_p(a.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (handle, p(x), B, T, T, F, None, None, 0, p(logits), p(doa))
    assert lib.crnn_forward(*args, p(work), need - 1, st) == _native.SALSA_ENOMEM
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    free_before = torch.cuda.mem_get_info()[0]
    for _ in range(3):
        _native.check(lib.crnn_forward(*args, p(work), need, st))
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() == before
    assert free_before - torch.cuda.mem_get_info()[0] < 64 << 20          # graph / module bookkeeping only, no activations
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(7), salsa_b200.SeldDecoder(512, decoder_type='bigru', freq_pool='avg', decoder_size=256))
    m.load_state_dict(sd)
    want = m.forward_ops(x)
    assert torch.equal(logits, want['event_frame_logit']) and torch.equal(doa, want['doa_frame_output'])
    _native.check(lib.crnn_free_model(handle))


def test_model_train_mode_runs_the_reference_training_step():
    """SeldModel.train() / training_step(batch) / eval(): the reference LightningModule's surface (models/seld_models.py:51-76)
    on salsa_b200.train.SeldTrainer; the weights trained in between reach the inference kernels at eval() and state_dict()."""
    import salsa_b200
    model = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                                 salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                        freq_pool='avg', decoder_size=256), label_rate=10, feature_rate=80.0)
    model.load_state_dict(salsa_b200.crnn.random_state_dict(5))
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 7, 128, 200, generator=g)
    batch = (x, (torch.rand(2, 16, 12, generator=g) > 0.6).float(), torch.randn(2, 16, 36, generator=g).clamp(-1, 1), ['a', 'b'])
    before = model(x.cuda())['event_frame_logit'].clone()
    with pytest.raises(RuntimeError):
        model.training_step(batch, 0)               # not in train mode
    model.train(dropout=False, lr=3e-4)
    losses = [model.training_step(batch, i)['loss'].item() for i in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert model.trainer.optimizer.step_count == 8
    w_trained = model.state_dict()['encoder.conv_block1.conv2.weight']
    model.eval()
    val = model.validation_step(batch, 0)
    assert np.isfinite(val['loss'].item()) and len(val['events']) == 2
    tgt, pred = model.common_step(batch)
    assert pred['event_frame_logit'].shape == tgt['event_frame_gt'].shape == (2, 16, 12)
    after = model(x.cuda())['event_frame_logit']
    assert not torch.equal(before, after)                                       # the inference path sees the new weights
    assert torch.equal(model.state_dict()['encoder.conv_block1.conv2.weight'].cpu(), w_trained.cpu())
