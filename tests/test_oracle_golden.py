"""The oracle restatement against the reference: frozen golden vectors (always) and the
live reference checkout (when mounted).  CPU only."""
import numpy as np
import pytest

from oracle import salsa, stft, synth

EIG_CASES = [(fmt, tag, trk, cond) for fmt in ('foa', 'mic')
             for tag, trk, cond in (('t5', True, 5.0), ('t0', True, 0.0), ('n5', False, 5.0), ('t2', True, 2.0))]


@pytest.mark.parametrize('fmt,tag,trk,cond', EIG_CASES)
@pytest.mark.parametrize('batched', [False, True])
def test_eigenvector_matches_golden(golden, fmt, tag, trk, cond, batched):
    g = golden('eigvec_cases')
    fn = salsa.extract_normalized_eigenvector_batched if batched else salsa.extract_normalized_eigenvector
    out = fn(g['X'].copy(), condition_number=cond, n_hopframes=3, is_tracking=trk, audio_format=fmt,
             fs=24000, n_fft=512, lower_bin=1)
    ref = g['{}_{}'.format(fmt, tag)]
    assert out.shape == ref.shape and out.dtype == np.float64
    assert np.array_equal(out != 0, ref != 0)            # valid-bin mask: exact
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)


def test_eigenvector_bad_format():
    X = np.ones((2, 8, 4), dtype=complex)
    with pytest.raises(ValueError):
        salsa.extract_normalized_eigenvector(X, audio_format='xyz', fs=24000, n_fft=512, lower_bin=1)


def test_band_matrix_matches_golden(golden):
    g = golden('clip_cases')
    assert np.array_equal(salsa.band_matrix(512), g['W512'])
    assert np.array_equal(salsa.band_matrix(256), g['W256'])
    first, count, weight = salsa.band_table(512)
    assert first[0] == 1 and first[191] == 192 and first[192] == 193 and first[199] == 249
    assert count[191] == 1 and count[192] == 8 and count[199] == 7 and weight[199] == np.float32(0.125)
    with pytest.raises(AssertionError):
        salsa.band_matrix(1024)


def test_doa_bins():
    assert salsa.doa_bins(24000, 512, 50, 9000) == (1, 192)
    assert salsa.doa_bins(24000, 512, 50, 4000) == (1, 85)
    assert salsa.doa_bins(24000, 512, 50, 2000) == (1, 42)
    assert salsa.doa_bins(24000, 256, 50, 9000) == (1, 96)
    assert salsa.doa_bins(24000, 512, 50, 20000) == (1, 256)


def test_logspec_matches_golden(golden):
    g = golden('clip_cases')
    out = salsa.MagStftExtractor(512, 300, 512).extract(g['audio_foa'])
    assert out.dtype == np.float32
    np.testing.assert_array_equal(out, g['logspec_foa'])
    out = salsa.MagStftExtractor(512, 300, 512, is_compress_high_freq=False).extract(g['audio_foa'])
    np.testing.assert_array_equal(out, g['logspec_foa_nocompress'])


@pytest.mark.parametrize('key,fmt,fmax,trk', [('salsa_foa', 'foa', 9000, True), ('salsa_mic', 'mic', 4000, True),
                                              ('salsa_foa_notracking', 'foa', 9000, False)])
def test_salsa_clip_matches_golden(golden, key, fmt, fmax, trk):
    g = golden('clip_cases')
    out = salsa.salsa_clip(g['audio_' + fmt], fmt, fmax_doa=fmax, is_tracking=trk)
    assert out.shape == (7, 81, 200) and out.dtype == np.float32
    assert np.array_equal(out != 0, g[key] != 0)
    np.testing.assert_allclose(out, g[key], rtol=0, atol=1e-6)


def test_nfft256_oracle_matches_reference_golden(golden):
    """n_fft = 256 / hop 150 (salsa_feature_extraction.py:151-152, :163-170, :300-306): the restatement against outputs of
    the unmodified reference (tests/golden/nfft256_cases.npz, oracle/make_golden.py nfft256)."""
    g, clips = golden('nfft256_cases'), golden('clip_cases')
    foa, mic = clips['audio_foa'][:, :12000], clips['audio_mic'][:, :12000]
    np.testing.assert_array_equal(salsa.MagStftExtractor(256, 150, 256).extract(foa), g['logspec_foa'])
    np.testing.assert_array_equal(salsa.MagStftExtractor(256, 150, 256, is_compress_high_freq=False).extract(foa), g['logspec_foa_nocompress'])
    np.testing.assert_array_equal(salsa.MagStftExtractor(256, 150, 200).extract(foa), g['logspec_foa_win200'])
    for fmt, audio, fmax in (('foa', foa, 9000), ('mic', mic, 4000)):
        out = salsa.salsa_clip(audio, fmt, n_fft=256, hop_length=150, win_length=256, fmax_doa=fmax)
        ref = g['salsa_' + fmt]
        assert out.shape == ref.shape == (7, 81, 100)
        assert np.array_equal(out != 0, ref != 0)
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6)
    for ft in ('salsa_lite', 'salsa_ipd'):
        out = salsa.salsa_lite_clip(mic, ft, n_fft=256, hop_length=150)
        assert out.shape == g[ft].shape == (7, 81, 95)
        np.testing.assert_allclose(out, g[ft], rtol=0, atol=1e-6)


@pytest.mark.parametrize('ft', ['salsa_lite', 'salsa_ipd'])
def test_salsa_lite_matches_golden(golden, ft):
    g = golden('clip_cases')
    out = salsa.salsa_lite_clip(g['audio_mic'], ft)
    assert out.shape == (7, 81, 191) and out.dtype == np.float32
    np.testing.assert_array_equal(out, g[ft])
    assert np.all(out[4:, :, 42:] == 0) and np.any(out[4:, :, 41] != 0)   # crop quirk, lite :118-120


def test_synth_is_deterministic(golden):
    g = golden('clip_cases')
    np.testing.assert_allclose(synth.make_clip(3, 'foa', seconds=1.0), g['audio_foa'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(synth.make_clip(4, 'mic', seconds=1.0), g['audio_mic'], rtol=0, atol=1e-6)


def test_stft_against_torch_and_scipy():
    """The librosa restatement has no librosa to be pinned against; cross-check two other
    independent implementations of the same transform."""
    import scipy.signal
    import torch
    y = synth.make_clip(5, 'foa', seconds=0.5)[0]
    S = stft.stft(y, n_fft=512, hop_length=300, center=True, window='hann', pad_mode='reflect')
    assert S.shape == (257, 1 + len(y) // 300) and S.dtype == np.complex64
    T = torch.stft(torch.from_numpy(y).double(), n_fft=512, hop_length=300,
                   window=torch.hann_window(512, periodic=True, dtype=torch.float64), center=True,
                   pad_mode='reflect', return_complex=True).numpy()
    assert np.abs(S - T).max() <= 2e-7 * np.abs(T).max()
    ypad = np.pad(y.astype(np.float64), 256, mode='reflect')
    _, _, Z = scipy.signal.stft(ypad, window=scipy.signal.get_window('hann', 512, fftbins=True), nperseg=512,
                                noverlap=212, nfft=512, boundary=None, padded=False, scaling='spectrum')
    Z = Z * scipy.signal.get_window('hann', 512, fftbins=True).sum()
    assert np.abs(S - Z[:, :S.shape[1]]).max() <= 2e-7 * np.abs(Z).max()


def test_power_to_db():
    x = np.array([0.0, 1e-12, 1.0, 100.0], dtype=np.float32)
    out = stft.power_to_db(x, ref=1.0, amin=1e-10, top_db=None)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, [-100.0, -100.0, 0.0, 20.0], atol=1e-5)


def test_scaler_matches_sklearn():
    from sklearn import preprocessing
    rng = np.random.default_rng(0)
    feats = [rng.standard_normal((7, 11, 6)).astype(np.float32) * 3 + 1 for _ in range(3)]
    mean, std = salsa.compute_scaler(feats)
    for ch in range(4):
        sc = preprocessing.StandardScaler()
        for f in feats:
            sc.partial_fit(f[ch])
        np.testing.assert_allclose(mean[ch, 0], sc.mean_, rtol=1e-6)
        np.testing.assert_allclose(std[ch, 0], np.sqrt(sc.var_), rtol=1e-6)


# ---------------------------------------------------------------- live reference (build container only)
@pytest.mark.reference
def test_live_reference_eigenvector():
    from oracle import ref_import
    from oracle.make_golden import structured_spectrum
    ref = ref_import.features_module()
    X = structured_spectrum(seed=11, n_bins=12, n_frames=40)
    for fmt in ('foa', 'mic'):
        a = ref.extract_normalized_eigenvector(X.copy(), condition_number=5.0, is_tracking=True, audio_format=fmt,
                                               fs=24000, n_fft=512, lower_bin=1)
        b = salsa.extract_normalized_eigenvector_batched(X.copy(), condition_number=5.0, is_tracking=True,
                                                         audio_format=fmt, fs=24000, n_fft=512, lower_bin=1)
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)


@pytest.mark.reference
def test_live_reference_driver_bodies():
    from oracle import ref_import
    from oracle.make_golden import DATA_CFG
    clip = synth.make_clip(6, 'mic', seconds=0.6)
    a = ref_import.run_driver_body('salsa', clip, DATA_CFG['mic'])
    np.testing.assert_allclose(a, salsa.salsa_clip(clip, 'mic', fmax_doa=4000), rtol=0, atol=1e-6)
    for ft in ('salsa_lite', 'salsa_ipd'):
        a = ref_import.run_driver_body('salsa_lite', clip, DATA_CFG['lite'], feature_type=ft)
        np.testing.assert_array_equal(a, salsa.salsa_lite_clip(clip, ft))


# --------------------------------------------------------------------------------------------------
# CRNN oracle (oracle/crnn.py) against outputs of the unmodified reference modules
# --------------------------------------------------------------------------------------------------
def test_crnn_oracle_matches_golden(golden):
    import torch
    from oracle import crnn as ocrnn
    g = golden('model_cases')
    sd = ocrnn.make_state_dict(0)
    x = ocrnn.model_input(2, (2, 7, 128, 200))
    with torch.no_grad():
        enc = ocrnn.encoder_forward(sd, x)
    y = ocrnn.forward(sd, x)
    assert tuple(enc.shape) == (2, 512, 8, 12)
    np.testing.assert_allclose(enc.numpy(), g['model_encoder_out'].astype(np.float32), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(y['event_frame_logit'].numpy(), g['model_event_frame_logit'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(y['doa_frame_output'].numpy(), g['model_doa_frame_output'], rtol=0, atol=2e-5)
    y2 = ocrnn.forward(sd, ocrnn.model_input(3, (1, 7, 96, 191)))
    assert tuple(y2['event_frame_logit'].shape) == (1, 6, 12)
    np.testing.assert_allclose(y2['event_frame_logit'].numpy(), g['lite_event_frame_logit'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(y2['doa_frame_output'].numpy(), g['lite_doa_frame_output'], rtol=0, atol=2e-5)


def test_interpolate_matches_golden(golden):
    import torch
    from oracle import crnn as ocrnn
    g = golden('model_cases')
    x = torch.from_numpy(g['interp_in'])
    assert np.array_equal(ocrnn.interpolate_tensor(x, 0.5).numpy(), g['interp_half'])
    assert np.array_equal(ocrnn.interpolate_tensor(x, 2.0).numpy(), g['interp_double'])
    assert np.array_equal(ocrnn.interpolate_tensor(torch.arange(40).reshape(1, 40, 1), 2.0).numpy().ravel(), g['interp_40_to_80'])


@pytest.mark.reference
def test_crnn_oracle_matches_live_reference():
    import torch
    from oracle import crnn as ocrnn, ref_import
    model = ref_import.build_reference_seld_model()
    model.load_state_dict(ocrnn.make_state_dict(1), strict=True)
    model.eval()
    x = ocrnn.model_input(7, (1, 7, 64, 200))
    with torch.no_grad():
        ref = model(x)
    out = ocrnn.forward(ocrnn.make_state_dict(1), x)
    for k in ref:
        np.testing.assert_allclose(out[k].numpy(), ref[k].numpy(), rtol=0, atol=2e-5)


# ------------------------------------------------------------------------------------------------
# augmentations: oracle/augment.py against the unmodified reference classes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('fmt', ['foa', 'mic'])
def test_augment_oracle_matches_reference_golden(golden, fmt):
    from oracle import augment as oaug
    g = golden('augment_cases')
    x, y_doa = g['x'], g['y_doa']
    seen = set()
    for seed in range(24):
        np.random.seed(seed)                         # replay the reference's draws in its order
        m = oaug.draw_swap_foa() if fmt == 'foa' else oaug.draw_swap_mic()
        xa, ya = (x, y_doa) if m is None else (oaug.swap_foa if fmt == 'foa' else oaug.swap_mic)(x, y_doa, m)
        sh = oaug.draw_shift(x.shape[2])
        if sh is not None:
            xa = oaug.shift_updown(xa, *sh)
        seen.add((None if m is None else tuple(int(v) for v in m), sh))
        assert np.array_equal(xa, g['{}_{}_x'.format(fmt, seed)]), (fmt, seed)          # bit-exact: permutations, signs, one subtraction
        assert np.array_equal(ya, g['{}_{}_y_doa'.format(fmt, seed)]), (fmt, seed)
    assert len(seen) > 12                            # the seeds exercise many different draws


def test_composite_cutout_draws_and_fill_match_reference_golden(golden):
    """MIC training transforms with CompositeCutout behind the shift (dataset/datamodule.py:76-82): the host-side draws of
    salsa_b200.augment (same NumPy calls in the same order as utilities/transforms.py:58-283) + the oracle's fill reproduce
    the arrays the unmodified reference classes produced under np.random.seed.  No GPU involved."""
    from oracle import augment as oaug
    from salsa_b200 import augment
    g = golden('extras_cases')
    x, y_doa = g['cut_x'], g['cut_y_doa']
    batch = augment.BatchAugment(augment.TfmapRandomSwapChannelMic(n_classes=12), augment.RandomShiftUpDownNp(freq_shift_range=10),
                                 augment.CompositeCutout(image_aspect_ratio=32 / 48, n_zero_channels=3))
    kinds = set()
    for seed in range(30):
        np.random.seed(seed)
        ops, cuts = batch.draw(1, x.shape[2], x.shape[1])
        m = [(ops[0, 1] >> i) & 1 for i in range(3)]
        xa, ya = oaug.swap_mic(x, y_doa, m)
        if ops[0, 2]:
            xa = oaug.shift_updown(xa, int(ops[0, 2]), 'up' if ops[0, 3] == 0 else 'down')
        xa = oaug.cutout_rects(xa, cuts[0], n_zero_channels=3)
        kinds.add(len(cuts[0]))
        assert np.array_equal(xa, g['cut_{}_x'.format(seed)]), seed
        assert np.array_equal(ya, g['cut_{}_y_doa'.format(seed)]), seed
    assert kinds == {0, 1, 2, 8}                     # skipped, random cutout, SpecAugment stripes, holes


def test_linspec_iv_oracle_matches_reference_golden(golden):
    """oracle.salsa.linspec_iv_clip restates LinSpecIvExtractor.extract (dataset/feature_extraction.py:316-358); the golden array
    is the unmodified class's output on the golden FOA clip."""
    from oracle import salsa as osalsa
    out = osalsa.linspec_iv_clip(golden('clip_cases')['audio_foa'])
    ref = golden('extras_cases')['linspeciv_foa']
    assert out.shape == ref.shape == (7, 81, 200)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6)
    # the spectrogram part is MagStftExtractor's (same W, same window): identical to the SALSA golden
    np.testing.assert_allclose(ref[:4], golden('clip_cases')['logspec_foa'], rtol=0, atol=1e-6)


def test_linspec_gcc_oracle_matches_reference_golden(golden):
    """oracle.salsa.linspec_gcc_clip restates LogSpecGccExtractor (dataset/feature_extraction.py:362-482); the golden array is the
    unmodified class's output on the first 0.5 s of the golden MIC clip."""
    from oracle import salsa as osalsa
    out = osalsa.linspec_gcc_clip(golden('clip_cases')['audio_mic'][:, :12000])
    ref = golden('extras_cases')['linspecgcc_mic']
    assert out.shape == ref.shape == (10, 41, 200)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6)
    assert np.abs(ref[4:]).max() > 0.3               # real correlation peaks, not noise
