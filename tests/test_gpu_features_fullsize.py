"""Feature path at BASELINE.json's full clip size (4-channel 24 kHz 60 s, T = 4801 frames): one direct
comparison against the oracle and size-independent properties on a batch.

Properties (all follow from the reference's definition, dataset/salsa_feature_extraction.py:17-129, :177-201):
  * clips are independent: batch rows == one-by-one rows, bit for bit;
  * FOA channel sign flip: negating input channel c (c = 1..3) negates spatial output channel c-1 and leaves
    everything else -- spectrograms, masks, the other two components -- bit-identical (R -> D R D, u -> D u);
  * gain: multiplying the audio by 2 adds 10 log10(4) dB to every spectrogram value above the amin floor and leaves the
    spatial channels unchanged wherever the selection is unchanged (the tracker's 1e-6 floor is the only absolute scale);
  * the spatial FOA vector has unit norm on valid bins, MIC values are bounded by pi / (delta * bin);
  * bins above upper_bin are zero; the time wrap makes frame 0 see the last frames (checked via the oracle).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_FULL = 24000 * 60


@pytest.fixture(scope='module')
def sb():
    import salsa_b200
    assert torch.cuda.is_available()
    return salsa_b200


@pytest.fixture(scope='module')
def clips():
    sys_path_bench = __import__('bench')
    return sys_path_bench.make_clips(torch, 6, 'foa', torch.device('cuda'), seed=123)


def test_full_clip_matches_oracle(sb, clips):
    from oracle import salsa as osalsa
    audio = clips[0].cpu().numpy()
    assert audio.shape == (4, N_FULL)
    ref = osalsa.salsa_clip(audio, 'foa')                       # stacked-LAPACK form, ~10 s
    out = sb.SalsaExtractor('foa').extract(clips[:1]).cpu().numpy()[0]
    assert out.shape == ref.shape == (7, 4801, 200)
    assert (np.abs(out[:4] - ref[:4]) / np.maximum(1.0, np.abs(ref[:4]))).max() <= 1e-4
    sup_a, sup_b = out[4:] != 0, ref[4:] != 0
    assert int(np.count_nonzero(sup_a != sup_b)) == 0, 'valid-bin mask differs on a full 60 s clip'
    assert sup_b.mean() > 0.1
    assert np.abs(out[4:] - ref[4:]).max() <= 1e-4


def test_batch_rows_are_independent(sb, clips):
    ex = sb.SalsaExtractor('foa')
    batch = ex.extract(clips)
    for i in (0, 3, 5):
        one = ex.extract(clips[i:i + 1])
        assert torch.equal(one[0].view(torch.int32), batch[i].view(torch.int32))


@pytest.mark.parametrize('ch', [1, 2, 3])
def test_foa_channel_sign_flip(sb, clips, ch):
    ex = sb.SalsaExtractor('foa')
    base = ex.extract(clips[1:2])[0]
    flipped_audio = clips[1:2].clone()
    flipped_audio[:, ch] *= -1
    out = ex.extract(flipped_audio)[0]
    assert torch.equal(out[:4].view(torch.int32), base[:4].view(torch.int32))          # |X|^2 unchanged
    for c in range(3):
        expect = -base[4 + c] if c == ch - 1 else base[4 + c]
        # identical magnitudes and support; -0.0 == 0.0
        assert torch.equal(out[4 + c] != 0, base[4 + c] != 0)
        assert torch.allclose(out[4 + c], expect, rtol=0, atol=2e-6)


def test_gain_scaling(sb, clips):
    ex = sb.SalsaExtractor('foa')
    base = ex.extract(clips[2:3])[0]
    loud = ex.extract(clips[2:3] * 2.0)[0]
    above_floor = base[:4] > -99.0
    assert torch.allclose(loud[:4][above_floor], base[:4][above_floor] + 10 * np.log10(4.0), rtol=0, atol=1e-4)
    same = (base[4:] != 0) == (loud[4:] != 0)
    assert same.float().mean() > 0.999                  # only bins sitting on the absolute 1e-6 floor may change
    both = (base[4:] != 0) & (loud[4:] != 0)
    assert torch.allclose(loud[4:][both], base[4:][both], rtol=0, atol=1e-5)


def test_value_ranges_and_padding(sb, clips):
    out = sb.SalsaExtractor('foa').extract(clips[3:4])[0]
    assert torch.isfinite(out).all()
    assert torch.all(out[4:, :, 191:] == 0)
    valid = out[4] != 0
    norm = torch.sqrt((out[4:] ** 2).sum(dim=0))
    assert torch.allclose(norm[valid], torch.ones_like(norm[valid]), atol=1e-5)
    mic_audio = __import__('bench').make_clips(torch, 1, 'mic', torch.device('cuda'), seed=321)
    mic = sb.SalsaExtractor('mic', fmax_doa=4000).extract(mic_audio)[0]
    assert torch.all(mic[4:, :, 84:] == 0)
    delta = 2 * np.pi * 24000 / (512 * 343.0)
    bound = torch.tensor(np.pi / (delta * np.arange(1, 85)), dtype=torch.float32, device='cuda')
    assert torch.all(mic[4:, :, :84].abs() <= bound * (1 + 1e-6))
    lite = sb.SalsaLiteExtractor().extract(mic_audio)[0]
    assert lite.shape == (7, 4801, 191) and torch.all(lite[4:, :, 42:] == 0) and torch.isfinite(lite).all()


@pytest.mark.parametrize('fmt,fmax', [('foa', 9000), ('mic', 4000)])
def test_split_and_fused_arrangements_agree(sb, clips, fmt, fmax, monkeypatch):
    """The clip path exists in two arrangements of the same arithmetic (salsa_abi.cu: split = STFT -> X in HBM ->
    tracker -> eig_tile_kernel, fused = X in a shared-memory ring): same selection bit for bit, values to float32
    rounding (the FFT is compiled into two kernels; a handful of spectrogram values differ in the last bit)."""
    audio = clips[:3]
    monkeypatch.delenv('SALSA_B200_PIPELINE', raising=False)
    split = sb.SalsaExtractor(fmt, fmax_doa=fmax).extract(audio)
    monkeypatch.setenv('SALSA_B200_PIPELINE', 'fused')
    fused = sb.SalsaExtractor(fmt, fmax_doa=fmax).extract(audio)
    monkeypatch.delenv('SALSA_B200_PIPELINE', raising=False)
    assert torch.equal(split[:, 4:] != 0, fused[:, 4:] != 0)
    assert (split[:, :4] - fused[:, :4]).abs().max().item() <= 1e-5
    assert ((split[:, :4] != fused[:, :4]).float().mean().item()) < 1e-6
    assert (split[:, 4:] - fused[:, 4:]).abs().max().item() <= 2e-6


def test_full_clip_mic_matches_oracle(sb):
    from oracle import salsa as osalsa, synth
    audio = synth.make_clip(5, 'mic', seconds=60.0)
    ref = osalsa.salsa_clip(audio, 'mic', fmax_doa=4000)
    out = sb.SalsaExtractor('mic', fmax_doa=4000).extract(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]
    assert out.shape == ref.shape == (7, 4801, 200)
    assert (np.abs(out[:4] - ref[:4]) / np.maximum(1.0, np.abs(ref[:4]))).max() <= 1e-4
    sup_a, sup_b = out[4:] != 0, ref[4:] != 0
    assert int(np.count_nonzero(sup_a != sup_b)) == 0
    # a phase within rounding of +-pi may come out with the other sign: compare away from the cut
    delta = 2 * np.pi * 24000 / (512 * 343.0)
    phase = np.abs(ref[4:, :, :84] * delta * np.arange(1, 85))
    keep = np.abs(phase - np.pi) > 1e-3
    assert np.abs(out[4:, :, :84] - ref[4:, :, :84])[keep].max() <= 1e-4
    assert np.all(out[4:, :, 84:] == 0)
