"""Host-side logic that needs no GPU: the augmentation draws (same NumPy calls in the same order as the reference) and the
learning-rate / beta1 schedule."""
import numpy as np
import pytest

from oracle import augment as oaug
from salsa_b200 import augment, optim


@pytest.mark.parametrize('fmt', ['foa', 'mic'])
def test_batch_augment_draws_follow_the_reference_order(fmt):
    """BatchAugment.draw consumes np.random exactly like one DataLoader worker of the reference does per sample:
    joint transform (rand < p, randint flags), then the shift (rand < p, randint length, choice of direction)."""
    joint = augment.TfmapRandomSwapChannelFoa() if fmt == 'foa' else augment.TfmapRandomSwapChannelMic()
    batch = augment.BatchAugment(joint, augment.RandomShiftUpDownNp(freq_shift_range=10))
    np.random.seed(123)
    ops = batch.draw(64, 200)
    np.random.seed(123)
    for b in range(64):
        m = oaug.draw_swap_foa() if fmt == 'foa' else oaug.draw_swap_mic()
        flags = 0 if m is None else sum(int(v) << i for i, v in enumerate(m))
        sh = oaug.draw_shift(200, freq_shift_range=10)
        want = [0 if fmt == 'foa' else 1, flags, 0 if sh is None else sh[0], 0 if sh is None or sh[1] == 'up' else 1]
        assert ops[b].tolist() == want, (b, ops[b], want)
    assert len({tuple(r) for r in ops.tolist()}) > 20


def test_shift_range_default_and_argument_checks():
    sh = augment.RandomShiftUpDownNp(always_apply=True)
    np.random.seed(0)
    s, d = sh.draw(200)
    assert sh.freq_shift_range == 16 and 1 <= s < 16 and d in (0, 1)          # int(n_features * 0.08), transforms.py:300-301
    with pytest.raises(ValueError):
        augment.RandomShiftUpDownNp(direction='left')
    assert augment.RandomShiftUpDownNp(always_apply=True, direction='down', freq_shift_range=5).draw(100)[1] == 1


def test_learning_rate_schedule_matches_np_interp():
    """utilities/learning_utils.py:17-52 with the values of experiments/configs/seld.yml: piecewise linear between
    (0, 0.45, 0.9, 1.0) x n_steps."""
    s = optim.LearningRateScheduler(steps_per_epoch=100, max_epochs=50)
    assert s.n_steps == 5000 and s.step_milestones == [0, 2250, 4500, 5000]
    assert s.at(0, 0) == (1e-4, 0.9)
    lr, mom = s.at(22, 50)
    assert lr == pytest.approx(1e-2) and mom == pytest.approx(0.8)
    lr, mom = s.at(11, 25)                                      # half way up the first ramp
    assert lr == pytest.approx((1e-4 + 1e-2) / 2) and mom == pytest.approx(0.85)
    assert s.at(49, 99)[0] == pytest.approx(np.interp(4999, [0, 2250, 4500, 5000], (1e-4, 1e-2, 1e-3, 1e-4)))
    assert s.at(60, 0) == (1e-4, 0.9)                           # beyond the last milestone np.interp holds the end value


def test_driver_descriptions_and_wav_reader(tmp_path):
    """Directory names of the script-level seam (dataset/salsa_feature_extraction.py:315-321, salsa_lite...:69) and the 16-bit
    wav reader behind it."""
    import wave
    import numpy as np
    from salsa_b200 import driver
    cfg = {'data': dict(format='foa', fs=24000, n_fft=512, hop_len=300, win_len=512, fmin_doa=50, fmax_doa=9000)}
    assert driver.feature_description(cfg) == '24000fs_512nfft_300nhop_5cond_9000fmaxdoa'
    assert driver.feature_description(cfg, cond_num=0, is_tracking=False, is_compress_high_freq=False) == \
        '24000fs_512nfft_300nhop_0cond_9000fmaxdoa_notracking_nocompress'
    cfg['data']['fmax_doa'] = 20000                               # clipped at fs / 2 (:299)
    assert driver.feature_description(cfg, 'salsa_lite') == '24000fs_512nfft_300nhop_12000fmaxdoa'
    pcm = (np.random.default_rng(0).integers(-3000, 3000, size=(4, 1000))).astype(np.int16)
    path = str(tmp_path / 'a.wav')
    with wave.open(path, 'wb') as w:
        w.setnchannels(4)
        w.setsampwidth(2)
        w.setframerate(24000)
        w.writeframes(np.ascontiguousarray(pcm.T).tobytes())
    assert np.array_equal(driver.read_wav_pcm16(path, 24000), pcm)
    import pytest
    with pytest.raises(ValueError):
        driver.read_wav_pcm16(path, 16000)
