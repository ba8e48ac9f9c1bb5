"""GPU augmentations (salsa_b200.augment -> crnn_augment) against the reference: the golden outputs of the unmodified
`utilities/transforms.py` classes (tests/golden/augment_cases.npz) and the oracle restatement on random draws.  Everything
is index / sign permutation plus single float32 subtractions, so the bar is bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def aug():
    import salsa_b200
    assert torch.cuda.is_available()
    return salsa_b200.augment


@pytest.mark.parametrize('fmt', ['foa', 'mic'])
def test_seeded_draws_reproduce_the_reference(aug, golden, fmt):
    """np.random.seed(k) + the classes of this package pick the same augmentation as the reference's classes did and
    produce the same arrays (one sample per seed, like the reference's DataLoader worker)."""
    g = golden('augment_cases')
    x = torch.from_numpy(g['x'])[None].cuda()
    y_sed = torch.from_numpy(g['y_sed'])[None].cuda()
    y_doa = torch.from_numpy(g['y_doa'])[None].cuda()
    joint = aug.TfmapRandomSwapChannelFoa(n_classes=12) if fmt == 'foa' else aug.TfmapRandomSwapChannelMic(n_classes=12)
    batch = aug.BatchAugment(joint, aug.RandomShiftUpDownNp(freq_shift_range=10))
    for seed in range(24):
        np.random.seed(seed)
        xa, ya_sed, ya_doa = batch(x, y_sed, y_doa)
        assert ya_sed is y_sed
        assert np.array_equal(xa.cpu().numpy()[0], g['{}_{}_x'.format(fmt, seed)]), (fmt, seed)
        assert np.array_equal(ya_doa.cpu().numpy()[0], g['{}_{}_y_doa'.format(fmt, seed)]), (fmt, seed)


@pytest.mark.parametrize('fmt', ['foa', 'mic'])
def test_batch_with_every_op_matches_oracle(aug, fmt):
    """All swap-flag combinations x {no shift, up, down} in one batch at the training chunk size."""
    from oracle import augment as oaug
    n_flags = 16 if fmt == 'foa' else 8
    ops = np.array([[0 if fmt == 'foa' else 1, m, s, d] for m in range(n_flags) for s, d in ((0, 0), (3, 0), (9, 1), (1, 1))], dtype=np.int32)
    B = len(ops)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, 7, 80, 200)).astype(np.float32)
    y_doa = rng.standard_normal((B, 10, 36)).astype(np.float32)
    xa, _, ya = aug.BatchAugment()(torch.from_numpy(x).cuda(), None, torch.from_numpy(y_doa).cuda(), ops=ops)
    xa, ya = xa.cpu().numpy(), ya.cpu().numpy()
    for b in range(B):
        m = [(ops[b, 1] >> i) & 1 for i in range(4 if fmt == 'foa' else 3)]
        xr, yr = (oaug.swap_foa if fmt == 'foa' else oaug.swap_mic)(x[b], y_doa[b], m)
        if ops[b, 2]:
            xr = oaug.shift_updown(xr, int(ops[b, 2]), 'up' if ops[b, 3] == 0 else 'down')
        assert np.array_equal(xa[b].view(np.int32), xr.view(np.int32)), (fmt, b, ops[b])      # bits, including -0.0
        assert np.array_equal(ya[b].view(np.int32), yr.view(np.int32)), (fmt, b, ops[b])


def test_augment_argument_checks(aug):
    x = torch.zeros((2, 7, 8, 16), device='cuda')
    with pytest.raises(ValueError):
        aug.BatchAugment()(x[:, :6], None)
    with pytest.raises(ValueError):
        aug.BatchAugment()(x, None, ops=np.array([[0, 0, 16, 0], [0, 0, 0, 0]], dtype=np.int32))     # shift_len >= n_features
    with pytest.raises(NotImplementedError):
        aug.RandomShiftUpDownNp(n_last_channels=6)
    out, _, y = aug.BatchAugment()(x, None)                   # nothing configured: identity, labels untouched
    assert torch.equal(out, x) and y is None


def test_composite_cutout_reproduces_the_reference(aug, golden):
    """swap -> shift -> CompositeCutout on the device (crnn_augment + crnn_cutout), seeded like the reference's loader: the
    arrays of the unmodified reference classes, bit for bit (fill value = float32(min + (max - min) u) with the sample's
    min / max found on the device)."""
    g = golden('extras_cases')
    x = torch.from_numpy(g['cut_x'])[None].cuda()
    y_doa = torch.from_numpy(g['cut_y_doa'])[None].cuda()
    batch = aug.BatchAugment(aug.TfmapRandomSwapChannelMic(n_classes=12), aug.RandomShiftUpDownNp(freq_shift_range=10),
                             aug.CompositeCutout(image_aspect_ratio=32 / 48, n_zero_channels=3))
    for seed in range(30):
        np.random.seed(seed)
        xa, _, ya = batch(x, None, y_doa)
        assert np.array_equal(xa.cpu().numpy()[0], g['cut_{}_x'.format(seed)]), seed
        assert np.array_equal(ya.cpu().numpy()[0], g['cut_{}_y_doa'.format(seed)]), seed
    # a batch at the training chunk size against the oracle: every sample its own rectangles
    from oracle import augment as oaug
    rng = np.random.default_rng(8)
    xb = (rng.standard_normal((6, 7, 640, 200)) * 12 - 50).astype(np.float32)
    np.random.seed(123)
    cut = aug.CompositeCutout(always_apply=True, image_aspect_ratio=640 / 200, n_zero_channels=3)
    b2 = aug.BatchAugment(None, None, cut)
    ops, cuts = b2.draw(6, 200, 640)
    out, _, _ = b2(torch.from_numpy(xb).cuda(), None, None, ops=(ops, cuts))
    for b in range(6):
        assert np.array_equal(out[b].cpu().numpy(), oaug.cutout_rects(xb[b], cuts[b], n_zero_channels=3)), b
