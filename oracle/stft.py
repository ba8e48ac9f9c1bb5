"""NumPy restatement of `librosa.stft` / `librosa.power_to_db` (librosa 0.8.0).

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.

librosa is an un-vendored dependency of the reference (`requirements.yml:101`,
librosa==0.8.0) and is not installed here.  Call sites on the hot path:
`dataset/salsa_feature_extraction.py:186-192` and `:360-361`,
`dataset/salsa_lite_feature_extraction.py:97-98` (stft);
`salsa_feature_extraction.py:195`, `salsa_lite_feature_extraction.py:105`
(power_to_db).  Published algorithm (librosa 0.8.0 `core/spectrum.py`):

* window = `scipy.signal.get_window(window, win_length, fftbins=True)`
  (periodic, float64), centre-padded with zeros to `n_fft`;
* `center=True` -> `np.pad(y, n_fft // 2, mode=pad_mode)`;
* frames `y[t*hop : t*hop + n_fft]`, `n_frames = 1 + (len(y_padded) - n_fft) // hop`;
* `np.fft.rfft(window * frames, axis=0)` -- the float64 window promotes the
  product, so the transform runs in float64;
* the result is stored into a complex64 array when `y` is float32
  (`util.dtype_r2c`), i.e. the float64 spectrum is ROUNDED to float32.

Parity for this file is unpinned against librosa itself (not installable
here); it is cross-checked against `torch.stft` and `scipy.signal.stft` in
`tests/test_oracle_golden.py`.
"""
import numpy as np
import scipy.signal


def fft_window(window: str, win_length: int, n_fft: int) -> np.ndarray:
    """Periodic window, float64, zero-padded symmetrically to n_fft."""
    w = scipy.signal.get_window(window, win_length, fftbins=True).astype(np.float64)
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad), mode='constant')
    return w


def n_stft_frames(n_samples: int, hop_length: int) -> int:
    """Frames produced with center=True: 1 + n_samples // hop."""
    return 1 + n_samples // hop_length


def stft(y, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
         pad_mode='reflect', dtype=None):
    """Returns (1 + n_fft//2, n_frames); complex64 for float32 input, else complex128."""
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError('stft expects a mono signal, got shape {}'.format(y.shape))
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    w = fft_window(window, win_length, n_fft)
    if center:
        if n_fft > y.shape[-1]:
            pass  # librosa only warns
        y = np.pad(y, int(n_fft // 2), mode=pad_mode)
    if y.shape[0] < n_fft:
        raise ValueError('input too short for n_fft={}'.format(n_fft))
    n_frames = 1 + (y.shape[0] - n_fft) // hop_length
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    out = np.empty((1 + n_fft // 2, n_frames), dtype=dtype, order='F')
    # strided view (n_fft, n_frames) over the padded signal, processed in column blocks
    frames = np.lib.stride_tricks.as_strided(
        y, shape=(n_fft, n_frames), strides=(y.strides[0], y.strides[0] * hop_length), writeable=False)
    block = max(1, (2 ** 18) // n_fft)
    wcol = w[:, None]
    for s in range(0, n_frames, block):
        e = min(s + block, n_frames)
        out[:, s:e] = np.fft.rfft(wcol * frames[:, s:e], axis=0)
    return out


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """10*log10(max(amin, S)) - 10*log10(max(amin, |ref|)); dtype of S is kept."""
    S = np.asarray(S)
    if amin <= 0:
        raise ValueError('amin must be strictly positive')
    if np.issubdtype(S.dtype, np.complexfloating):
        magnitude = np.abs(S)
    else:
        magnitude = S
    ref_value = np.abs(ref(magnitude)) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ValueError('top_db must be non-negative')
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec
