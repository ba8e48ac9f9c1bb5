"""Parity at the shapes the benchmark actually runs (BASELINE.json configs 1-3 and the CRNN metric), not only at the small
shapes of the op tests:

  * CRNN forward vs the float32 oracle at (2, 7, 640, 200) (a training chunk) and (1, 7, 4800, 200) (a full 60 s clip:
    GRU sequence 300, layer-4 maps 300 x 12), in both precision modes, with the measured error printed
    (SURVEY.md section 8c iv; models/seld_models.py:39-49);
  * the recurrent kernel alone at T = 300, B = 32 (the benchmark's batch);
  * SALSA-Lite and SALSA-IPD on a 60 s clip vs the oracle (dataset/salsa_lite_feature_extraction.py:94-123);
  * the float32-FFT variant (`stft_precision=32`) on a 60 s clip with its honest mask-mismatch count.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DELTA = 2 * np.pi * 24000 / (512 * 343.0)


def scale_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-6)


def _model(precision, seed=0):
    import salsa_b200
    from oracle import crnn as ocrnn
    m = salsa_b200.SeldModel(salsa_b200.PannResNet22(n_input_channels=7),
                             salsa_b200.SeldDecoder(512, n_classes=12, output_format='reg_xyz', decoder_type='bigru',
                                                    freq_pool='avg', decoder_size=256),
                             label_rate=10, feature_rate=80.0, precision=precision)
    sd = ocrnn.make_state_dict(seed)
    m.load_state_dict(sd)
    return m, sd


@pytest.fixture(scope='module')
def oracle_outputs():
    """float32 oracle forward at the two benchmark shapes (CPU: ~2 s and ~12 s), shared by both precision modes."""
    from oracle import crnn as ocrnn
    sd = ocrnn.make_state_dict(0)
    out = {}
    for shape, seed in (((2, 7, 640, 200), 21), ((1, 7, 4800, 200), 22)):
        x = ocrnn.model_input(seed, shape)
        out[shape] = (x, ocrnn.forward(sd, x))
    return out


# bars: bf16x3 and bf16x2 are the float32-parity modes (north_star: 1e-4 relative; measured 3-4e-5 and 6-8e-5); bf16 rounds
# every activation to 8 mantissa bits, its envelope is a few 1e-2 of the output scale and is measured, not hidden
@pytest.mark.parametrize('shape', [(2, 7, 640, 200), (1, 7, 4800, 200)])
@pytest.mark.parametrize('precision,bar', [('bf16x3', 1e-4), ('bf16x2', 1e-4), ('bf16', 6e-2)])
def test_crnn_forward_at_benchmark_shapes(oracle_outputs, shape, precision, bar):
    m, _ = _model(precision)
    x, ref = oracle_outputs[shape]
    y = m(x.cuda())
    T16 = shape[2] // 16
    assert tuple(y['event_frame_logit'].shape) == (shape[0], T16, 12)
    assert tuple(y['doa_frame_output'].shape) == (shape[0], T16, 36)
    errs = {k: scale_err(y[k].cpu().numpy(), ref[k].numpy()) for k in ref}
    print('CRNN[{}] {} vs float32 oracle: logits {:.3e}, doa {:.3e} (relative to the output scale)'.format(
        precision, shape, errs['event_frame_logit'], errs['doa_frame_output']))
    for k, e in errs.items():
        assert np.isfinite(e) and e < bar, (k, e)


@pytest.mark.parametrize('planes', [1, 3])
def test_gru_layer_at_benchmark_length(planes):
    """B = 32 clips, T = 300 steps (a 60 s clip at the encoder's 16x time reduction): error growth over the recurrence."""
    import salsa_b200
    ops = salsa_b200.crnn_ops
    B, T = 32, 300
    torch.manual_seed(77)
    gru = torch.nn.GRU(input_size=512, hidden_size=256, num_layers=1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, 512)
    with torch.no_grad():
        ref, _ = gru(x)
        xproj = torch.cat([x @ gru.weight_ih_l0.T + gru.bias_ih_l0, x @ gru.weight_ih_l0_reverse.T + gru.bias_ih_l0_reverse], dim=-1)
    w_hh = torch.stack([gru.weight_hh_l0, gru.weight_hh_l0_reverse]).detach().contiguous()
    b_hh = torch.stack([gru.bias_hh_l0, gru.bias_hh_l0_reverse]).detach().contiguous()
    y = ops.gru_layer(xproj.reshape(B * T, 1536).contiguous().cuda(), w_hh.cuda(), b_hh.cuda(), B, T, planes=planes)
    y = ops.merge_planes(y[:B * T].cpu(), planes).reshape(B, T, 512)
    err = (y - ref).abs()
    # error by position in the sequence: it must not grow along the recurrence
    head, tail = err[:, :30, :256].max().item(), err[:, -30:, :256].max().item()
    print('gru_layer planes={} B=32 T=300: max |err| {:.3e} (first 30 steps {:.3e}, last 30 steps {:.3e})'.format(
        planes, err.max().item(), head, tail))
    assert err.max().item() < (1e-2 if planes == 1 else 2e-5)


@pytest.mark.parametrize('feature_type', ['salsa_lite', 'salsa_ipd'])
def test_lite_full_clip_matches_oracle(feature_type):
    import salsa_b200
    from oracle import salsa as osalsa, synth
    audio = synth.make_clip(9, 'mic', seconds=60.0)
    ref = osalsa.salsa_lite_clip(audio, feature_type)
    out = salsa_b200.SalsaLiteExtractor(feature_type).extract(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]
    assert out.shape == ref.shape == (7, 4801, 191)
    spec_err = np.abs(out[:4] - ref[:4]) / np.maximum(1.0, np.abs(ref[:4]))
    assert spec_err.max() <= 1e-4
    assert np.all(out[4:, :, 42:] == 0) and np.all(ref[4:, :, 42:] == 0)
    # a phase within rounding of +-pi may come out with the other sign: compare away from the cut
    k = np.arange(1, 43)
    phase = np.abs(ref[4:, :, :42]) * (np.pi if feature_type == 'salsa_ipd' else DELTA * k)
    keep = np.abs(phase - np.pi) > 1e-3
    d = np.abs(out[4:, :, :42] - ref[4:, :, :42])
    print('{} 60 s clip: spectrogram max rel err {:.2e}, phase max |err| {:.2e} ({} of {} values within 1e-3 rad of the cut)'.format(
        feature_type, spec_err.max(), d[keep].max(), int((~keep).sum()), keep.size))
    assert d[keep].max() <= 1e-4
    assert keep.mean() > 0.999


def test_fp32_stft_clip_path_honest_mismatch_count():
    """`stft_precision=32` is the fast, NON-parity variant: the tracker follows the float32 spectrum, so a few selections
    flip.  This test pins how far it is from the reference on a full clip: mask mismatches counted (not required to be
    zero), spectrogram and spatial values compared on the common selection."""
    import salsa_b200
    from oracle import salsa as osalsa
    import bench
    audio = bench.make_clips(torch, 1, 'foa', torch.device('cuda'), seed=321)
    ref = osalsa.salsa_clip(audio[0].cpu().numpy(), 'foa')
    out = salsa_b200.SalsaExtractor('foa', stft_precision=32).extract(audio).cpu().numpy()[0]
    sup_a, sup_b = out[4:] != 0, ref[4:] != 0
    mism = int(np.count_nonzero(sup_a != sup_b))
    both = sup_a & sup_b
    spec = np.abs(out[:4] - ref[:4])
    loud = ref[:4] > -60.0                       # away from the near-silent bins where float32 cancellation shows
    sp = np.abs(out[4:] - ref[4:])[both]
    print('stft_precision=32, 60 s FOA clip: {} mask mismatches of {} bins ({:.1e}); spectrogram max |err| {:.2e} dB '
          '({:.2e} dB above -60 dB); spatial max |err| on the common selection {:.2e}'.format(
              mism, sup_b.size, mism / sup_b.size, spec.max(), spec[loud].max(), sp.max()))
    assert mism / sup_b.size < 1e-4
    assert spec[loud].max() < 1e-2 and spec.max() < 1.0
    assert np.quantile(sp, 0.9999) < 1e-3
