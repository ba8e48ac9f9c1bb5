"""CPU restatement of the reference's training-time augmentations that commute with the SALSA feature layout
(SURVEY.md section 8 f3): `utilities/transforms.py` TfmapRandomSwapChannelFoa (:368-437), TfmapRandomSwapChannelMic
(:440-523) and RandomShiftUpDownNp (:286-320), with the random draws made explicit.  Test infrastructure: only tests/
may import it.  Pinned by `tests/golden/augment_cases.npz` (outputs of the unmodified reference classes under
np.random.seed, `oracle/make_golden.py`)."""
import numpy as np


def draw_swap_foa(p: float = 0.5):
    """Random numbers in the order MapDataAugmentBase.__call__ (:346-352) and TfmapRandomSwapChannelFoa.apply (:409)
    consume them: returns m (4 flags) or None when the transform is skipped."""
    if not np.random.rand() < p:
        return None
    return np.random.randint(2, size=(4,))


def draw_swap_mic(p: float = 0.5):
    """As draw_swap_foa for TfmapRandomSwapChannelMic.apply (:485): 3 flags."""
    if not np.random.rand() < p:
        return None
    return np.random.randint(2, size=(3,))


def draw_shift(n_features: int, p: float = 0.5, freq_shift_range: int = 10):
    """DataAugmentNumpyBase.__call__ (:45-52) + RandomShiftUpDownNp.apply (:298-305): (shift_len, direction) or None."""
    if not np.random.rand() < p:
        return None
    if freq_shift_range is None:
        freq_shift_range = int(n_features * 0.08)
    shift_len = int(np.random.randint(1, freq_shift_range, 1)[0])
    direction = str(np.random.choice(['up', 'down'], 1)[0])
    return shift_len, direction


def swap_foa(x: np.ndarray, y_doa: np.ndarray, m, n_classes: int = 12):
    """TfmapRandomSwapChannelFoa.apply (:394-437).  x (7, T, F): W Y Z X | Y Z X; y_doa (Ty, 3 n_classes): x | y | z."""
    assert x.shape[0] == 7
    x_new, y_new = x.copy(), y_doa.copy()
    if m[0] == 1:          # swap x and y
        x_new[1], x_new[3] = x[3], x[1]
        x_new[-3], x_new[-1] = x[-1], x[-3]
    if m[1] == 1:
        x_new[-1] = -x_new[-1]
    if m[2] == 1:
        x_new[-3] = -x_new[-3]
    if m[3] == 1:
        x_new[-2] = -x_new[-2]
    n = n_classes
    if y_doa.shape[1] != 3 * n:
        raise NotImplementedError('this output format not yet implemented')
    if m[0] == 1:
        y_new[:, 0:n] = y_doa[:, n:2 * n]
        y_new[:, n:2 * n] = y_doa[:, :n]
    if m[1] == 1:
        y_new[:, 0:n] = -y_new[:, 0:n]
    if m[2] == 1:
        y_new[:, n:2 * n] = -y_new[:, n:2 * n]
    if m[3] == 1:
        y_new[:, 2 * n:] = -y_new[:, 2 * n:]
    return x_new, y_new


def swap_mic(x: np.ndarray, y_doa: np.ndarray, m, n_classes: int = 12):
    """TfmapRandomSwapChannelMic.apply (:470-523).  x (7, T, F): M1 M2 M3 M4 | p12 p13 p14."""
    assert x.shape[0] == 7
    x_new, y_new = x.copy(), y_doa.copy()
    if m[0] == 1:          # swap M2 and M3
        x_new[1], x_new[2] = x[2], x[1]
        x_new[-3], x_new[-2] = x[-2], x[-3]
    if m[1] == 1:          # swap M1 and M4
        c = x_new.copy()
        x_new[0], x_new[3] = c[3], c[0]
        x_new[-1] = -c[-1]
        x_new[-2] = c[-2] - c[-1]
        x_new[-3] = c[-3] - c[-1]
    if m[2] == 1:          # swap M1 and M2, M3 and M4
        c = x_new.copy()
        x_new[0], x_new[1], x_new[2], x_new[3] = c[1], c[0], c[3], c[2]
        x_new[-3] = -c[-3]
        x_new[-2] = c[-1] - c[-3]
        x_new[-1] = c[-2] - c[-3]
    n = n_classes
    if y_doa.shape[1] != 3 * n:
        raise NotImplementedError('this doa format not yet implemented')
    if m[0] == 1:
        y_new[:, 0:n] = y_doa[:, n:2 * n]
        y_new[:, n:2 * n] = y_doa[:, :n]
    if m[1] == 1:
        temp = -y_new[:, 0:n].copy()
        y_new[:, 0:n] = -y_new[:, n:2 * n]
        y_new[:, n:2 * n] = temp
    if m[2] == 1:
        y_new[:, n:2 * n] = -y_new[:, n:2 * n]
        y_new[:, 2 * n:] = -y_new[:, 2 * n:]
    return x_new, y_new


def shift_updown(x: np.ndarray, shift_len: int, direction: str, n_last_channels: int = 0):
    """RandomShiftUpDownNp.apply (:306-320), mode='reflect'."""
    n_features = x.shape[2]
    new = x.copy()
    part = new if n_last_channels == 0 else new[:-n_last_channels]
    if direction == 'up':
        shifted = np.pad(part, ((0, 0), (0, 0), (shift_len, 0)), mode='reflect')[:, :, 0:n_features]
    else:
        shifted = np.pad(part, ((0, 0), (0, 0), (0, shift_len)), mode='reflect')[:, :, shift_len:]
    if n_last_channels == 0:
        return shifted
    new[:-n_last_channels] = shifted
    return new


def cutout_rects(x: np.ndarray, rects, n_zero_channels: int = None):
    """The fill step shared by RandomCutoutNp (:100-123), SpecAugmentNp (:170-194) and RandomCutoutHoleNp (:233-252),
    is_filled_last_channels=True: rects = [(top, bottom, left, right, u)] in order; the value of a rectangle is
    np.random.uniform(min, max) = min + (max - min) * u with min / max of x BEFORE any cut, stored into the float32 array."""
    lo, hi = np.min(x), np.max(x)
    out = x.copy()
    for top, bottom, left, right, u in rects:
        c = float(lo) + (float(hi) - float(lo)) * u
        if n_zero_channels is None:
            out[:, top:bottom, left:right] = c
        else:
            out[:-n_zero_channels, top:bottom, left:right] = c
            out[-n_zero_channels:, top:bottom, left:right] = 0.0
    return out
