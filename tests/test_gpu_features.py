"""Parity of the CUDA feature path (through the C ABI) against the oracle and the golden vectors
frozen from the reference.  Needs a B200: run with `-m gpu`.

Tolerances (BASELINE.json north_star: bit-exact for indexing, 1e-4 relative for floats):
  * index maps / valid-bin masks: exact (mismatch count must be 0);
  * log spectrogram (dB, |values| up to 100): |a-b| <= 1e-4 * max(1, |b|);
  * spatial channels (unit vectors / normalised phases, |values| <= ~4): |a-b| <= 1e-4 * max(1, |b|),
    compared on the bins both sides mark valid.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def close(a, b, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert np.all(np.isfinite(a) == np.isfinite(b)), what
    err = np.where(np.isfinite(err), err, 0.0)
    assert err.max() <= RTOL, '{}: max rel err {:.3e} at {}'.format(what, err.max(), np.unravel_index(err.argmax(), err.shape))
    return err.max()


def check_feature(out, ref, n_spec=4, what='feature'):
    """spectrogram channels close; spatial channels: identical support, close values."""
    assert out.shape == ref.shape and out.dtype == np.float32
    close(out[:n_spec], ref[:n_spec], what + ' spectrogram')
    sup_a, sup_b = out[n_spec:] != 0, ref[n_spec:] != 0
    mismatch = int(np.count_nonzero(sup_a != sup_b))
    assert mismatch == 0, '{}: {} valid-bin mask mismatches out of {}'.format(what, mismatch, sup_a.size)
    close(out[n_spec:], ref[n_spec:], what + ' spatial')


@pytest.fixture(scope='module')
def sb():
    import salsa_b200
    assert torch.cuda.is_available()
    return salsa_b200


# ------------------------------------------------------------------------------------------------
# op level seams
# ------------------------------------------------------------------------------------------------
def test_stft_matches_oracle(sb, golden):
    from oracle import salsa as osalsa
    audio = golden('clip_cases')['audio_foa']
    ref = osalsa.multichannel_stft(audio, 512, 300)[1:256].astype(np.complex64)     # (255, T, 4)
    out = sb.stft(audio, n_fft=512, hop_length=300, lower_bin=1, upper_bin=256)
    assert out.shape == ref.shape and out.dtype == np.complex64
    # float64 transform rounded to float32 on both sides: every complex value within one float32 ulp
    assert np.all(np.abs(out - ref) <= 1.2e-7 * np.abs(ref))
    # ... and bit-identical almost everywhere.  Frame 0 is excluded from the count: its reflected
    # frame is symmetric, so its imaginary parts are pure float64 rounding noise (~1e-17) on both sides.
    exact = np.mean(out[:, 1:] == ref[:, 1:])
    assert exact > 0.9999, 'only {:.6f} of the complex64 values are bit-identical'.format(exact)


def test_stft_float32_variant(sb, golden):
    from oracle import salsa as osalsa
    audio = golden('clip_cases')['audio_mic']
    ref = osalsa.multichannel_stft(audio, 512, 300)[1:256]
    out = sb.stft(audio, n_fft=512, hop_length=300, lower_bin=1, upper_bin=256, stft_precision=32)
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize('compress,key', [(True, 'logspec_foa'), (False, 'logspec_foa_nocompress')])
def test_logspec_matches_golden(sb, golden, compress, key):
    g = golden('clip_cases')
    out = sb.MagStftExtractor(n_fft=512, hop_length=300, win_length=512, is_compress_high_freq=compress).extract(
        g['audio_foa'])
    assert out.dtype == np.float32
    close(out, g[key], key)


def test_logspec_other_window(sb, golden):
    from oracle import salsa as osalsa
    audio = golden('clip_cases')['audio_foa']
    ref = osalsa.MagStftExtractor(512, 300, 400, window='hamming').extract(audio)
    out = sb.MagStftExtractor(512, 300, 400, window='hamming').extract(audio)
    close(out, ref, 'hamming/400 window')


# n_fft = 256 (salsa_feature_extraction.py:151-152, :163-170, :300-306): against outputs of the unmodified reference
def test_nfft256_stft_and_logspec(sb, golden):
    from oracle import salsa as osalsa
    g, clips = golden('nfft256_cases'), golden('clip_cases')
    foa = clips['audio_foa'][:, :12000]
    ref = osalsa.multichannel_stft(foa, 256, 150)[1:128].astype(np.complex64)
    out = sb.stft(foa, n_fft=256, hop_length=150, lower_bin=1, upper_bin=128)
    assert out.shape == ref.shape and out.dtype == np.complex64
    assert np.all(np.abs(out - ref) <= 1.2e-7 * np.abs(ref))
    assert np.mean(out[:, 1:] == ref[:, 1:]) > 0.999
    close(sb.MagStftExtractor(256, 150, 256).extract(foa), g['logspec_foa'], 'logspec 256')
    close(sb.MagStftExtractor(256, 150, 256, is_compress_high_freq=False).extract(foa), g['logspec_foa_nocompress'], 'logspec 256 linear')
    close(sb.MagStftExtractor(256, 150, 200).extract(foa), g['logspec_foa_win200'], 'logspec 256, window 200')
    out32 = sb.stft(foa, n_fft=256, hop_length=150, lower_bin=1, upper_bin=128, stft_precision=32)
    assert np.abs(out32 - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize('fmt,fmax', [('foa', 9000), ('mic', 4000)])
def test_nfft256_clip_matches_reference_golden(sb, golden, fmt, fmax):
    g, clips = golden('nfft256_cases'), golden('clip_cases')
    audio = clips['audio_' + fmt][:, :12000]
    ex = sb.SalsaExtractor(audio_format=fmt, n_fft=256, hop_len=150, win_len=256, fmax_doa=fmax)
    out = ex.extract(torch.from_numpy(audio)[None].cuda())[0].cpu().numpy()
    assert out.shape == (7, 81, 100)
    check_feature(out, g['salsa_' + fmt], what='n_fft 256 ' + fmt)
    # the host-buffer entry point takes the same path
    host = ex.extract_host(np.ascontiguousarray(audio[None]))[0]
    assert np.array_equal(host, out)


@pytest.mark.parametrize('ft', ['salsa_lite', 'salsa_ipd'])
def test_nfft256_salsa_lite_matches_reference_golden(sb, golden, ft):
    g, clips = golden('nfft256_cases'), golden('clip_cases')
    audio = np.ascontiguousarray(clips['audio_mic'][:, :12000])
    ref = g[ft]
    ex = sb.SalsaLiteExtractor(feature_type=ft, n_fft=256, hop_len=150)
    out = ex.extract(torch.from_numpy(audio)[None].cuda())[0].cpu().numpy()
    assert out.shape == ref.shape == (7, 81, 95)
    close(out[:4], ref[:4], ft + ' 256 spectrogram')
    # a phase within rounding of +-pi may legitimately come out with the other sign
    scale = np.pi if ft == 'salsa_lite' else 1.0
    k = (np.arange(95) + 1)[None, None, :]
    delta = 2 * np.pi * 24000 / (256 * 343.0)
    raw = ref[4:] * (delta * k if ft == 'salsa_lite' else np.pi)
    ok = np.abs(np.abs(raw) - np.pi) > 1e-3
    assert ok.mean() > 0.99
    close(out[4:][ok], ref[4:][ok], ft + ' 256 phase')
    assert np.all(out[4:, :, ex.upper_bin:] == 0)
    host = ex.extract_host(audio[None])[0]
    assert np.array_equal(host, out)


def test_nfft256_is_rejected_where_it_is_not_built(sb):
    with pytest.raises((ValueError, NotImplementedError)):
        sb.LinSpecIvExtractor(n_fft=256, hop_length=150).extract(np.zeros((4, 12000), np.float32))


def test_extractor_asserts_like_reference(sb):
    with pytest.raises(AssertionError):
        sb.MagStftExtractor(n_fft=1024, hop_length=300)
    with pytest.raises(AssertionError):
        sb.MagStftExtractor(n_fft=512, hop_length=300, win_length=1024)


EIG_CASES = [(fmt, tag, trk, cond) for fmt in ('foa', 'mic')
             for tag, trk, cond in (('t5', True, 5.0), ('t2', True, 2.0), ('t0', True, 0.0), ('n5', False, 5.0))]


@pytest.mark.parametrize('fmt,tag,trk,cond', EIG_CASES)
def test_eigenvector_matches_golden(sb, golden, fmt, tag, trk, cond):
    from oracle import salsa as osalsa
    g = golden('eigvec_cases')
    X = g['X']
    ref = g['{}_{}'.format(fmt, tag)]
    out = sb.extract_normalized_eigenvector(X.copy(), condition_number=cond, n_hopframes=3, is_tracking=trk,
                                            audio_format=fmt, fs=24000, n_fft=512, lower_bin=1)
    assert out.shape == ref.shape and out.dtype == np.float64
    assert np.array_equal(out != 0, ref != 0), 'valid-bin mask differs'
    # Where the two largest eigenvalues are nearly equal the principal eigenvector is not determined
    # by the data (the reference returns whatever LAPACK's rotation is); compare where the gap is >= 2.
    _, aux = osalsa.extract_normalized_eigenvector_batched(
        X.copy(), condition_number=cond, is_tracking=trk, audio_format=fmt, fs=24000, n_fft=512, lower_bin=1,
        return_aux=True)
    s = aux['s']
    gap_ok = s[..., 0] >= 2.0 * s[..., 1]
    sel = np.broadcast_to(gap_ok[None], ref.shape) & (ref != 0)
    assert sel.sum() > 500
    if fmt == 'mic':
        # phases within 1e-3 rad of +-pi may legitimately wrap to the other sign
        delta = 2 * np.pi * 24000 / (512 * 343.0)
        k = (np.arange(X.shape[0]) + 1)[None, :, None]
        near_cut = np.abs(np.abs(ref * delta * k) - np.pi) < 1e-3
        sel &= ~near_cut
    close(out[sel], ref[sel], 'eigenvector {} {}'.format(fmt, tag))


def test_eigenvector_bad_format(sb):
    X = np.ones((2, 8, 4), dtype=complex)
    with pytest.raises(ValueError):
        sb.extract_normalized_eigenvector(X, audio_format='xyz', fs=24000, n_fft=512, lower_bin=1)


# ------------------------------------------------------------------------------------------------
# clip level
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('key,fmt,fmax,trk', [('salsa_foa', 'foa', 9000, True), ('salsa_mic', 'mic', 4000, True)])
def test_salsa_clip_matches_golden(sb, golden, key, fmt, fmax, trk):
    g = golden('clip_cases')
    ex = sb.SalsaExtractor(audio_format=fmt, fmax_doa=fmax, is_tracking=trk)
    audio = torch.from_numpy(g['audio_' + fmt])[None].cuda()
    out = ex.extract(audio).cpu().numpy()[0]
    check_feature(out, g[key], what=key)


def test_salsa_lite_matches_golden(sb, golden):
    g = golden('clip_cases')
    audio = torch.from_numpy(g['audio_mic'])[None].cuda()
    for ft in ('salsa_lite', 'salsa_ipd'):
        out = sb.SalsaLiteExtractor(feature_type=ft).extract(audio).cpu().numpy()[0]
        ref = g[ft]
        assert out.shape == ref.shape == (7, 81, 191)
        close(out[:4], ref[:4], ft + ' spectrogram')
        assert np.all(out[4:, :, 42:] == 0)
        scale = np.pi if ft == 'salsa_ipd' else 1.0
        if ft == 'salsa_lite':
            delta = 2 * np.pi * 24000 / (512 * 343.0)
            ang_ref = ref[4:, :, :42] * delta * np.arange(1, 43)
        else:
            ang_ref = ref[4:, :, :42] * np.pi
        near_cut = np.abs(np.abs(ang_ref) - np.pi) < 1e-3
        close(out[4:, :, :42][~near_cut], ref[4:, :, :42][~near_cut], ft + ' phase')


@pytest.mark.parametrize('fmt,fmax', [('foa', 9000), ('mic', 4000)])
def test_salsa_clip_matches_oracle_5s(sb, fmt, fmax):
    """A longer clip than the golden one, oracle computed on the spot (batched LAPACK form)."""
    from oracle import salsa as osalsa, synth
    audio = synth.make_clip(11, fmt, seconds=5.0)
    ref = osalsa.salsa_clip(audio, fmt, fmax_doa=fmax)
    out = sb.SalsaExtractor(audio_format=fmt, fmax_doa=fmax).extract(torch.from_numpy(audio)[None].cuda())
    check_feature(out.cpu().numpy()[0], ref, what='5 s {}'.format(fmt))


@pytest.mark.parametrize('fmt,fmin,fmax,cond', [
    ('foa', 200, 9000, 5.0),      # lower_bin 4: the first lanes of every 32-bin register group belong to the previous X tile
    ('foa', 1600, 6000, 5.0),     # lower_bin 34 >= 32: the general tiled-store path of stft_kernel
    ('mic', 50, 3000, 5.0),       # upper_bin 64: exactly two full tiles
    ('foa', 50, 9000, 2.0),       # cond_num 2 / 20: other power-iteration schedules than the compiled-in one
    ('foa', 50, 9000, 20.0),
])
def test_salsa_clip_other_ranges_and_thresholds(sb, fmt, fmin, fmax, cond):
    from oracle import salsa as osalsa, synth
    audio = synth.make_clip(17, fmt, seconds=2.0)
    ref, aux = osalsa.salsa_clip(audio, fmt, fmin_doa=fmin, fmax_doa=fmax, cond_num=cond, return_aux=True)
    ex = sb.SalsaExtractor(audio_format=fmt, fmin_doa=fmin, fmax_doa=fmax, cond_num=cond)
    out = ex.extract(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]
    close(out[:4], ref[:4], 'spectrogram')
    assert np.array_equal(out[4:] != 0, ref[4:] != 0), 'valid-bin mask'
    # with cond < 4 the two leading eigenvalues of a kept bin may be close: LAPACK's vector is then only determined
    # to what the gap allows, compare where the gap is at least 4
    gap = (aux['s'][..., 0] >= 4.0 * aux['s'][..., 1]).T                      # (T, n_bins)
    sel = np.zeros(ref[4:].shape, dtype=bool)
    sel[:, :, :gap.shape[1]] = gap[None]
    if fmt == 'mic':
        delta = 2 * np.pi * 24000 / (512 * 343.0)
        n_bins = gap.shape[1]
        phase = np.abs(ref[4:, :, :n_bins] * delta * (np.arange(n_bins) + ex.lower_bin))
        sel[:, :, :n_bins] &= np.abs(phase - np.pi) > 1e-3
    close(out[4:][sel], ref[4:][sel], 'spatial')


def test_salsa_clip_bins_reach_past_the_last_full_tile(sb):
    """fmax_doa 9300 Hz -> upper_bin 198: 197 spatial bins, i.e. 7 tiles of 32 = 224 > the 200 feature columns.  The last
    tile's row store must stop at the feature width (it used to spill into the next row); more spatial bins than feature
    columns is the reference's broadcast error (salsa_feature_extraction.py:373-374) -> ValueError."""
    from oracle import salsa as osalsa, synth
    audio = synth.make_clip(23, 'foa', seconds=2.0)
    ref = osalsa.salsa_clip(audio, 'foa', fmax_doa=9300)
    ex = sb.SalsaExtractor('foa', fmax_doa=9300)
    assert ex.upper_bin - ex.lower_bin == 197
    guard = torch.full((2, 7, ref.shape[1], 200), 7.0, device='cuda')          # clip 1 is a canary behind clip 0
    ex.extract(torch.from_numpy(audio)[None].cuda(), out=guard[:1])
    assert torch.all(guard[1] == 7.0), 'the eigenvector kernel wrote past the feature tensor'
    check_feature(guard[0].cpu().numpy(), ref, what='197 spatial bins')
    with pytest.raises(ValueError):
        sb.SalsaExtractor('foa', fmax_doa=9600).extract(torch.from_numpy(audio)[None].cuda())      # 203 bins > 200 columns


def test_salsa_clip_win_len_applies_to_the_spectrogram_only(sb):
    """win_len configures MagStftExtractor only (:324-325); the spectrum behind the spatial channels is librosa's default
    full-length Hann whatever win_len is (:359-361)."""
    from oracle import salsa as osalsa, synth
    audio = synth.make_clip(29, 'foa', seconds=2.0)
    ref = osalsa.salsa_clip(audio, 'foa', win_length=400)
    base = osalsa.salsa_clip(audio, 'foa')
    assert np.array_equal(ref[4:], base[4:]) and not np.array_equal(ref[:4], base[:4])
    out = sb.SalsaExtractor('foa', win_len=400).extract(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]
    check_feature(out, ref, what='win_len 400')
    # SALSA-Lite reads win_len from the config and never uses it (salsa_lite_feature_extraction.py:44, :97-98)
    a = torch.from_numpy(audio)[None].cuda()
    assert torch.equal(sb.SalsaLiteExtractor(win_len=400).extract(a), sb.SalsaLiteExtractor().extract(a))


def test_tracker_word_count_not_a_multiple_of_the_block(sb):
    """op-level seam with 300 bins: 10 mask words = one full block of 8 warps + a block with 6 surplus warps, which must
    not write mask words past their row."""
    from oracle import salsa as osalsa
    rng = np.random.default_rng(5)
    amp = rng.standard_normal((300, 40, 1)) + 1j * rng.standard_normal((300, 40, 1))
    amp[:, 10:25] *= 30.0                                              # a burst the tracker selects
    steer = rng.standard_normal((300, 1, 4)) + 1j * rng.standard_normal((300, 1, 4))
    X = amp * steer + 0.05 * (rng.standard_normal((300, 40, 4)) + 1j * rng.standard_normal((300, 40, 4)))
    X = X.astype(np.complex64).astype(np.complex128)
    kw = dict(audio_format='foa', fs=24000, n_fft=1024, lower_bin=1)
    ref = osalsa.extract_normalized_eigenvector(X, **kw)
    out = sb.extract_normalized_eigenvector(X, **kw)
    assert np.array_equal(out != 0, ref != 0) and (ref != 0).mean() > 0.1
    close(out, ref, '300-bin eigenvector op')


def test_out_buffers_are_validated(sb):
    lite = sb.SalsaLiteExtractor()
    a = torch.zeros((1, 4, 24000), device='cuda')
    with pytest.raises(ValueError):
        lite.extract(a, out=torch.empty((1, 7, 81, 190), device='cuda'))
    with pytest.raises(ValueError):
        lite.extract(a, out=torch.empty((1, 7, 81, 191), dtype=torch.float64, device='cuda'))
    with pytest.raises(ValueError):
        lite.extract_host(np.zeros((1, 4, 24000), np.float32), out=np.empty((1, 7, 80, 191), np.float32))
    with pytest.raises(ValueError):
        sb.SalsaExtractor('foa').extract(a, out=torch.empty((1, 7, 81, 200)))             # not on the device


def test_salsa_no_tracking_matches_oracle(sb, golden):
    from oracle import salsa as osalsa
    g = golden('clip_cases')
    audio = g['audio_foa']
    ref, aux = osalsa.salsa_clip(audio, 'foa', is_tracking=False, return_aux=True)
    np.testing.assert_allclose(ref, g['salsa_foa_notracking'], rtol=0, atol=1e-6)
    out = sb.SalsaExtractor('foa', is_tracking=False).extract(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]
    close(out[:4], ref[:4], 'no-tracking spectrogram')
    assert np.array_equal(out[4:] != 0, ref[4:] != 0)
    gap_ok = (aux['s'][..., 0] >= 2.0 * aux['s'][..., 1]).T                  # (T, n_bins)
    sel = np.zeros(ref[4:].shape, dtype=bool)
    sel[:, :, :gap_ok.shape[1]] = gap_ok[None]
    close(out[4:][sel], ref[4:][sel], 'no-tracking spatial (gap >= 2)')


def test_batch_equals_single_and_host_path(sb):
    """Clips are independent: a batch gives bit-identical rows to one-by-one calls, and the
    host-buffer entry point (chunked, 3 streams) gives bit-identical results to the device one."""
    from oracle import synth
    clips = np.stack([synth.make_clip(20 + i, 'foa', seconds=1.0 + 0.0 * i) for i in range(5)])
    ex = sb.SalsaExtractor('foa')
    batch = ex.extract(torch.from_numpy(clips).cuda()).cpu().numpy()
    for i in range(5):
        one = ex.extract(torch.from_numpy(clips[i:i + 1]).cuda()).cpu().numpy()[0]
        assert np.array_equal(one, batch[i], equal_nan=True)
    host = ex.extract_host(clips, clips_per_chunk=2)
    assert np.array_equal(host, batch, equal_nan=True)
    lite = sb.SalsaLiteExtractor()
    lb = lite.extract(torch.from_numpy(clips).cuda()).cpu().numpy()
    assert np.array_equal(lite.extract_host(clips, clips_per_chunk=3), lb)


def test_pcm16_host_input_is_bit_identical(sb):
    """16-bit PCM in (the wav files' samples; librosa.load returns sample / 32768, salsa_feature_extraction.py:353) ==
    float32 in, bit for bit; odd sample counts exercise the tail of the conversion kernel."""
    from oracle import synth
    clips = np.stack([synth.make_clip(60 + i, 'foa', seconds=1.0)[:, :23997] for i in range(3)])
    pcm = np.clip(np.round(clips * 32768.0), -32768, 32767).astype(np.int16)
    as_float = (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    ex = sb.SalsaExtractor('foa')
    ref = ex.extract_host(as_float, clips_per_chunk=2)
    out = ex.extract_host(pcm, clips_per_chunk=2)
    assert np.array_equal(out, ref, equal_nan=True) and (ref[:, 4:] != 0).mean() > 0.05
    assert np.array_equal(ref, ex.extract(torch.from_numpy(as_float).cuda()).cpu().numpy(), equal_nan=True)


def test_empty_batch_and_bad_shapes(sb):
    ex = sb.SalsaExtractor('foa')
    out = ex.extract(torch.empty((0, 4, 24000), device='cuda'))
    assert tuple(out.shape) == (0, 7, 81, 200)
    with pytest.raises(ValueError):
        ex.extract(torch.zeros((1, 3, 24000), device='cuda'))
    with pytest.raises(ValueError):
        ex.extract(torch.zeros((1, 4, 100), device='cuda'))       # shorter than the reflect padding
    with pytest.raises(ValueError):
        sb.SalsaExtractor('xyz')


def test_silence_and_ragged_lengths(sb):
    """Digital silence: spectrogram at the -100 dB floor, no valid bins; odd lengths (frame count
    1 + N // hop, last frame reflected) agree with the oracle."""
    from oracle import salsa as osalsa, synth
    ex = sb.SalsaExtractor('foa')
    out = ex.extract(torch.zeros((1, 4, 12000), device='cuda')).cpu().numpy()[0]
    assert np.all(out[:4] == -100.0) and np.all(out[4:] == 0)
    for n in (7321, 9000, 8999):
        audio = synth.make_clip(31, 'foa', seconds=1.0)[:, :n].copy()
        ref = osalsa.salsa_clip(audio, 'foa')
        got = ex.extract(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]
        assert got.shape == ref.shape == (7, 1 + n // 300, 200)
        check_feature(got, ref, what='ragged n={}'.format(n))


def test_scaler_matches_oracle_and_sklearn(sb):
    """compute_scaler (salsa_feature_extraction.py:204-262): StandardScaler statistics of channels 0..3."""
    from oracle import salsa as osalsa, synth
    from sklearn import preprocessing
    clips = np.stack([synth.make_clip(40 + i, 'foa', seconds=1.0) for i in range(3)])
    feats = sb.SalsaExtractor('foa').extract(torch.from_numpy(clips).cuda())
    sc = sb.FeatureScaler()
    sc.partial_fit(feats[:2])
    sc.partial_fit(feats[2:])
    mean, std = sc.finalize()
    assert mean.shape == std.shape == (4, 1, 200) and mean.dtype == np.float32
    f = feats.cpu().numpy()
    for ch in range(4):
        ref = preprocessing.StandardScaler()
        for i in range(3):
            ref.partial_fit(f[i, ch])
        np.testing.assert_allclose(mean[ch, 0], ref.mean_, rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(std[ch, 0], np.sqrt(ref.var_), rtol=1e-5, atol=1e-5)
    omean, ostd = osalsa.compute_scaler([f[i] for i in range(3)])
    np.testing.assert_allclose(mean, omean, rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(std, ostd, rtol=1e-5, atol=1e-5)
    m2, s2 = sb.compute_scaler([feats])
    assert np.array_equal(m2, mean) and np.array_equal(s2, std)


def test_linspec_iv_matches_golden_and_oracle(sb, golden):
    """LinSpecIvExtractor (dataset/feature_extraction.py:273-358) on the shared STFT front-end: spectrogram channels to
    1e-4 max(1, |ref|), intensity-vector channels (values in [-1, 1]) to 1e-4 absolute."""
    from oracle import salsa as osalsa, synth
    ref = golden('extras_cases')['linspeciv_foa']
    ex = sb.LinSpecIvExtractor(n_fft=512, hop_length=300, win_length=512)
    out = ex.extract(golden('clip_cases')['audio_foa'])
    assert out.shape == ref.shape and out.dtype == np.float32
    close(out[:4], ref[:4], 'linspeciv spectrogram')
    err = np.abs(out[4:] - ref[4:])
    print('linspeciv golden: IV max |err| {:.2e}'.format(err.max()))
    assert err.max() <= 1e-4
    clips = np.stack([synth.make_clip(70 + i, 'foa', seconds=3.0) for i in range(2)])
    batch = ex.extract_batch(torch.from_numpy(clips).cuda()).cpu().numpy()
    for i in range(2):
        want = osalsa.linspec_iv_clip(clips[i])
        close(batch[i, :4], want[:4], 'linspeciv spectrogram (3 s)')
        assert np.abs(batch[i, 4:] - want[4:]).max() <= 1e-4
        assert np.all(np.abs(batch[i, 4:]) <= 1.0 + 1e-6)
    with pytest.raises(NotImplementedError):
        sb.LinSpecIvExtractor(n_fft=512, hop_length=300, is_compress_high_freq=False)


def test_script_level_seam_writes_the_reference_layout(sb, tmp_path):
    """extract_features(data_config, ...) (dataset/salsa_feature_extraction.py:265-385): wav directory in, the reference's
    directory layout out, features equal to the oracle on the same 16-bit audio, scaler = compute_scaler of the dev split."""
    import wave
    from oracle import salsa as osalsa, synth
    from salsa_b200 import driver
    data_dir, feat_dir = tmp_path / 'data', tmp_path / 'feat'
    clips = {}
    for split, n in (('foa_dev', 3), ('foa_eval', 1)):
        (data_dir / split).mkdir(parents=True)
        for i in range(n):
            pcm = np.clip(np.round(synth.make_clip(90 + 10 * len(clips) + i, 'foa', seconds=1.0 + 0.5 * (i == 2)) * 32768), -32768, 32767).astype(np.int16)
            name = 'fold1_room1_mix{:03d}.wav'.format(i)
            with wave.open(str(data_dir / split / name), 'wb') as w:
                w.setnchannels(4)
                w.setsampwidth(2)
                w.setframerate(24000)
                w.writeframes(np.ascontiguousarray(pcm.T).tobytes())
            clips[(split, name)] = pcm.astype(np.float32) / np.float32(32768.0)
    cfg = {'data_dir': str(data_dir), 'feature_dir': str(feat_dir),
           'data': dict(format='foa', fs=24000, n_fft=512, hop_len=300, win_len=512, fmin_doa=50, fmax_doa=9000)}
    written = {}
    root = driver.extract_features(cfg, batch_clips=2, writer=lambda path, arrays: written.__setitem__(path, arrays))
    assert root == str(feat_dir / 'salsa' / 'foa' / '24000fs_512nfft_300nhop_5cond_9000fmaxdoa')
    dev_feats = []
    for (split, name), audio in clips.items():
        f = written[str(feat_dir / 'salsa' / 'foa' / '24000fs_512nfft_300nhop_5cond_9000fmaxdoa' / split / name.replace('wav', 'h5'))]['feature']
        ref = osalsa.salsa_clip(audio, 'foa')
        check_feature(f, ref, what='{}/{}'.format(split, name))
        if split == 'foa_dev':
            dev_feats.append(ref)
    sc = written[str(feat_dir / 'salsa' / 'foa' / '24000fs_512nfft_300nhop_5cond_9000fmaxdoa' / 'foa_feature_scaler.h5')]
    mean, std = osalsa.compute_scaler(dev_feats)
    assert sc['mean'].shape == (4, 1, 200) and sc['mean'].dtype == np.float32
    np.testing.assert_allclose(sc['mean'], mean, rtol=0, atol=2e-4)
    np.testing.assert_allclose(sc['std'], std, rtol=0, atol=2e-4)
    assert (feat_dir / 'salsa' / 'foa' / '24000fs_512nfft_300nhop_5cond_9000fmaxdoa' / 'foa_eval').is_dir()
    # task='scaler' (:391-393): compute_scaler on its own, from the feature files of the dev split read back
    root_dir = feat_dir / 'salsa' / 'foa' / '24000fs_512nfft_300nhop_5cond_9000fmaxdoa'
    for path in list(written):
        if '/foa_dev/' in path:
            open(path, 'wb').close()                     # the files the listing finds; their content comes from `written`
    again = {}
    driver.extract_features(cfg, task='scaler', batch_clips=2, writer=lambda path, arrays: again.__setitem__(path, arrays),
                            feature_reader=lambda path: written[path]['feature'])
    assert list(again) == [str(root_dir / 'foa_feature_scaler.h5')]
    np.testing.assert_allclose(again[str(root_dir / 'foa_feature_scaler.h5')]['mean'], sc['mean'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(again[str(root_dir / 'foa_feature_scaler.h5')]['std'], sc['std'], rtol=0, atol=1e-6)


def test_linspec_gcc_matches_golden_and_oracle(sb, golden):
    """LogSpecGccExtractor (dataset/feature_extraction.py:362-482): the 1024-point zero-padded STFT assembled from three
    512-point transforms, unit cross-spectrum phasors, and the inverse transform restricted to 200 lags as a tcgen05 GEMM
    (two bf16 planes per operand).  Spectrogram channels to 1e-4 max(1, |ref|), GCC channels (|values| <= 1) to 1e-4."""
    from oracle import salsa as osalsa, synth
    ref = golden('extras_cases')['linspecgcc_mic']
    ex = sb.LogSpecGccExtractor(n_fft=512, hop_length=300, win_length=512)
    out = ex.extract(golden('clip_cases')['audio_mic'][:, :12000])
    assert out.shape == ref.shape == (10, 41, 200) and out.dtype == np.float32
    close(out[:4], ref[:4], 'linspecgcc spectrogram')
    err = np.abs(out[4:] - ref[4:])
    print('linspecgcc golden: GCC max |err| {:.2e} (peak {:.2f})'.format(err.max(), np.abs(ref[4:]).max()))
    assert err.max() <= 1e-4
    clips = np.stack([synth.make_clip(75 + i, 'mic', seconds=2.0) for i in range(3)])
    ex2 = sb.LogSpecGccExtractor(n_fft=512, hop_length=300, clips_per_chunk=2)          # two chunks
    batch = ex2.extract_batch(torch.from_numpy(clips).cuda()).cpu().numpy()
    for i in range(3):
        want = osalsa.linspec_gcc_clip(clips[i])
        close(batch[i, :4], want[:4], 'linspecgcc spectrogram (2 s)')
        assert np.abs(batch[i, 4:] - want[4:]).max() <= 1e-4
    with pytest.raises(NotImplementedError):
        sb.LogSpecGccExtractor(n_fft=512, hop_length=300, win_length=400)
