"""Host-side mirror of the reference feature interfaces, running on libsalsa_b200.so.

Same names, argument meaning and error behaviour as
`dataset/salsa_feature_extraction.py` (`extract_normalized_eigenvector` :17-129,
`MagStftExtractor` :132-201) and the per-clip bodies of its `extract_features` (:353-377) and of
`dataset/salsa_lite_feature_extraction.py` (:94-123).  NumPy in / NumPy out for the
op-level seams (so a parity test reads like a test of the reference); CUDA tensors in / out
for the batch extractors.  PyTorch is only used for device memory and streams.
"""
import ctypes
import math

import numpy as np
import torch

from . import _native
from ._native import SalsaParams

__all__ = ['doa_bins', 'MagStftExtractor', 'LinSpecIvExtractor', 'LogSpecGccExtractor', 'extract_normalized_eigenvector', 'SalsaExtractor',
           'SalsaLiteExtractor', 'stft', 'FeatureScaler', 'compute_scaler']


def doa_bins(fs, n_fft, fmin_doa, fmax_doa):
    """(lower_bin, upper_bin) as computed at salsa_feature_extraction.py:296-304."""
    fmax_doa = min(fmax_doa, fs // 2)
    lower_bin = int(math.floor(fmin_doa * n_fft / float(fs)))
    upper_bin = int(math.floor(fmax_doa * n_fft / float(fs)))
    return max(1, lower_bin), upper_bin


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('salsa_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _check_out(out, shape, device):
    if not isinstance(out, torch.Tensor) or tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 \
            or not out.is_contiguous() or out.device != device:
        raise ValueError('out must be a contiguous float32 tensor of shape {} on {}'.format(tuple(shape), device))


def _window_table(window, win_length, n_fft):
    """None for the built-in periodic Hann, else a float64 table like librosa builds
    (scipy.signal.get_window(..., fftbins=True), centre padded)."""
    if window == 'hann':
        return None
    import scipy.signal
    w = scipy.signal.get_window(window, win_length, fftbins=True).astype(np.float64)
    lpad = (n_fft - win_length) // 2
    return np.ascontiguousarray(np.pad(w, (lpad, n_fft - win_length - lpad)))


def _params(n_clips, n_samples, fs=24000, n_fft=512, hop_len=300, win_len=None, lower_bin=1, upper_bin=192,
            audio_format='foa', is_tracking=True, is_compress_high_freq=True, n_hopframes=3, cond_num=5.0,
            stft_precision=64, window_table=None):
    if audio_format not in ('foa', 'mic'):
        raise ValueError('audio format {} is not valid'.format(audio_format))
    p = SalsaParams()
    p.n_clips, p.n_chans, p.n_samples = int(n_clips), 4, int(n_samples)
    p.fs, p.n_fft, p.hop_len = int(fs), int(n_fft), int(hop_len)
    p.win_len = int(n_fft if win_len is None else win_len)
    p.lower_bin, p.upper_bin = int(lower_bin), int(upper_bin)
    p.audio_format = _native.FORMAT_FOA if audio_format == 'foa' else _native.FORMAT_MIC
    p.is_tracking = int(bool(is_tracking))
    p.is_compress_high_freq = int(bool(is_compress_high_freq))
    p.n_hopframes = int(n_hopframes)
    p.stft_precision = int(stft_precision)
    p.cond_num = float(cond_num)
    if window_table is not None:
        p._keepalive = window_table
        p.window = window_table.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    return p


# ------------------------------------------------------------------------------------------------
# op level
# ------------------------------------------------------------------------------------------------
def stft(audio, n_fft=512, hop_length=300, win_length=None, window='hann', lower_bin=0, upper_bin=None,
         stft_precision=64):
    """Multichannel centred STFT of one clip on the GPU: (4, N) float32 -> complex64
    (upper_bin - lower_bin, T, 4), the layout the reference builds at :359-366.
    Bin 0 and the Nyquist bin are never used by the reference and are not produced (lower_bin >= 1)."""
    _require_cuda()
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    if upper_bin is None:
        upper_bin = n_fft // 2
    lower_bin = max(1, lower_bin)
    p = _params(1, audio.shape[1], n_fft=n_fft, hop_len=hop_length, win_len=win_length, lower_bin=lower_bin,
                upper_bin=upper_bin, stft_precision=stft_precision,
                window_table=_window_table(window, win_length or n_fft, n_fft))
    lib = _native.lib()
    T = lib.salsa_n_frames(p.n_samples, p.hop_len)
    d_audio = torch.from_numpy(audio).cuda()
    X = torch.empty((T, 4, upper_bin - lower_bin, 2), dtype=torch.float32, device='cuda')
    _native.check(lib.salsa_stft(ctypes.byref(p), _ptr(d_audio), _ptr(X), None, None, _stream()))
    Xc = torch.view_as_complex(X).permute(2, 0, 1).contiguous().cpu().numpy()
    return Xc


class MagStftExtractor:
    """Log-linear spectrogram extractor, same constructor and `extract` contract as the
    reference class (salsa_feature_extraction.py:132-201)."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int = None, window: str = 'hann',
                 is_compress_high_freq: bool = True, stft_precision: int = 64):
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.window = window
        self.win_length = self.n_fft if win_length is None else win_length
        assert self.win_length <= self.n_fft, 'Windown length is greater than nfft!'
        assert n_fft == 512 or n_fft == 256, 'nfft is not 512 or 256'
        self.is_compress_high_freq = is_compress_high_freq
        self.stft_precision = stft_precision
        self.n_bands = (200 if n_fft == 512 else 100) if is_compress_high_freq else n_fft // 2

    def extract(self, audio_input: np.ndarray) -> np.ndarray:
        """(4, n_samples) float32 -> (4, n_timeframes, n_bands) float32."""
        _require_cuda()
        audio = np.ascontiguousarray(audio_input, dtype=np.float32)
        if audio.ndim != 2 or audio.shape[0] != 4:
            raise ValueError('audio_input must be (4, n_samples), got {}'.format(audio.shape))
        p = _params(1, audio.shape[1], n_fft=self.n_fft, hop_len=self.hop_length, win_len=self.win_length, upper_bin=self.n_fft // 2,
                    is_compress_high_freq=self.is_compress_high_freq, stft_precision=self.stft_precision,
                    window_table=_window_table(self.window, self.win_length, self.n_fft))
        lib = _native.lib()
        T = lib.salsa_n_frames(p.n_samples, p.hop_len)
        d_audio = torch.from_numpy(audio).cuda()
        spec = torch.empty((4, T, self.n_bands), dtype=torch.float32, device='cuda')
        _native.check(lib.salsa_stft(ctypes.byref(p), _ptr(d_audio), None, _ptr(spec), None, _stream()))
        return spec.cpu().numpy()


class LinSpecIvExtractor:
    """Log-linear spectrogram + intensity vector (FOA), same constructor and `extract` contract as the reference class
    (dataset/feature_extraction.py:273-358).  `extract_batch` is the device-resident batch form."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int = None, window: str = 'hann',
                 is_compress_high_freq: bool = True, stft_precision: int = 64):
        self.n_fft, self.hop_length, self.window = n_fft, hop_length, window
        self.eps = 1e-8
        self.win_length = self.n_fft if win_length is None else win_length
        assert self.win_length <= self.n_fft, 'Windown length is greater than nfft!'
        assert n_fft == 512 or n_fft == 256, 'nfft is not 512 or 256'
        if n_fft != 512 or not is_compress_high_freq:
            raise NotImplementedError('salsa_b200 implements n_fft = 512 with is_compress_high_freq=True')
        self.stft_precision = stft_precision
        self.n_bands = 200
        self._workspace = None

    def extract_batch(self, audio: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """audio (B, 4, N) float32 CUDA -> (B, 7, T, 200) float32 CUDA, asynchronous on the current stream."""
        _require_cuda()
        if audio.dim() != 3 or audio.shape[1] != 4 or audio.dtype != torch.float32 or not audio.is_cuda:
            raise ValueError('audio must be a CUDA float32 tensor of shape (B, 4, N)')
        audio = audio.contiguous()
        B, _, N = audio.shape
        p = _params(B, N, n_fft=self.n_fft, hop_len=self.hop_length, win_len=self.win_length, stft_precision=self.stft_precision,
                    window_table=_window_table(self.window, self.win_length, self.n_fft))
        lib = _native.lib()
        T = lib.salsa_n_frames(N, self.hop_length)
        if out is None:
            out = torch.empty((B, 7, T, self.n_bands), dtype=torch.float32, device=audio.device)
        else:
            _check_out(out, (B, 7, T, self.n_bands), audio.device)
        need = lib.salsa_linspec_iv_workspace_bytes(ctypes.byref(p))
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != audio.device:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=audio.device)
        with _native.device_of(audio) as st:
            _native.check(lib.salsa_linspec_iv(ctypes.byref(p), _ptr(audio), _ptr(out), _ptr(self._workspace), self._workspace.numel(), st))
        return out

    def extract(self, audio_input: np.ndarray) -> np.ndarray:
        """(4, n_samples) float32 -> (7, n_timeframes, 200) float32, like the reference."""
        audio = np.ascontiguousarray(audio_input, dtype=np.float32)
        if audio.ndim != 2 or audio.shape[0] != 4:
            raise ValueError('audio_input must be (4, n_samples), got {}'.format(audio.shape))
        return self.extract_batch(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]


class LogSpecGccExtractor:
    """Log-linear spectrogram + GCC-PHAT of the six microphone pairs (MIC format), same constructor and `extract` contract
    as the reference class (dataset/feature_extraction.py:362-482): (4, n_samples) -> (10, n_timeframes, 200).
    `extract_batch` is the device-resident batch form (it walks the batch in chunks: the workspace is ~0.4 GB per 60 s clip)."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int = None, window: str = 'hann',
                 is_compress_high_freq: bool = True, stft_precision: int = 64, clips_per_chunk: int = 16):
        self.n_fft, self.hop_length, self.window = n_fft, hop_length, window
        self.win_length = self.n_fft if win_length is None else win_length
        assert self.win_length <= self.n_fft, 'Windown length is greater than nfft!'
        assert n_fft == 512 or n_fft == 256, 'nfft is not 512 or 256'
        if n_fft != 512 or not is_compress_high_freq or self.win_length != n_fft or window != 'hann':
            raise NotImplementedError('salsa_b200 implements n_fft = win_length = 512, window="hann", is_compress_high_freq=True')
        self.stft_precision = stft_precision
        self.n_freqs = 200
        self.clips_per_chunk = int(clips_per_chunk)
        self._workspace = None

    def extract_batch(self, audio: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """audio (B, 4, N) float32 CUDA -> (B, 10, T, 200) float32 CUDA."""
        _require_cuda()
        if audio.dim() != 3 or audio.shape[1] != 4 or audio.dtype != torch.float32 or not audio.is_cuda:
            raise ValueError('audio must be a CUDA float32 tensor of shape (B, 4, N)')
        audio = audio.contiguous()
        B, _, N = audio.shape
        lib = _native.lib()
        T = lib.salsa_n_frames(N, self.hop_length)
        if out is None:
            out = torch.empty((B, 10, T, self.n_freqs), dtype=torch.float32, device=audio.device)
        else:
            _check_out(out, (B, 10, T, self.n_freqs), audio.device)
        for c0 in range(0, B, self.clips_per_chunk):
            n = min(self.clips_per_chunk, B - c0)
            p = _params(n, N, n_fft=self.n_fft, hop_len=self.hop_length, win_len=self.win_length, stft_precision=self.stft_precision)
            need = lib.salsa_logspec_gcc_workspace_bytes(ctypes.byref(p))
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != audio.device:
                self._workspace = torch.empty(need, dtype=torch.uint8, device=audio.device)
            with _native.device_of(audio) as st:
                _native.check(lib.salsa_logspec_gcc(ctypes.byref(p), _ptr(audio[c0:c0 + n]), _ptr(out[c0:c0 + n]), _ptr(self._workspace),
                                                    self._workspace.numel(), st))
        return out

    def extract(self, audio_input: np.ndarray) -> np.ndarray:
        audio = np.ascontiguousarray(audio_input, dtype=np.float32)
        if audio.ndim != 2 or audio.shape[0] != 4:
            raise ValueError('audio_input must be (4, n_samples), got {}'.format(audio.shape))
        return self.extract_batch(torch.from_numpy(audio)[None].cuda()).cpu().numpy()[0]


def extract_normalized_eigenvector(X, condition_number: float = 5.0, n_hopframes: int = 3, is_tracking: bool = True,
                                   audio_format: str = 'foa', fs: int = None, n_fft: int = None,
                                   lower_bin: int = None):
    """Drop-in for the reference function of the same name (salsa_feature_extraction.py:17-129).

    X: (n_bins, n_frames, n_chans=4) complex, already cropped to the DOA band.
    Returns (3, n_bins, n_frames) float64, zeros where the bin is not a valid single-source bin.
    """
    _require_cuda()
    if audio_format not in ('foa', 'mic'):
        raise ValueError('audio format {} is not valid'.format(audio_format))
    X = np.ascontiguousarray(X, dtype=np.complex128)
    if X.ndim != 3 or X.shape[2] != 4:
        raise ValueError('X must be (n_bins, n_frames, 4), got {}'.format(X.shape))
    n_bins, n_frames, _ = X.shape
    if audio_format == 'mic' and (fs is None or n_fft is None or lower_bin is None):
        raise TypeError("audio_format='mic' needs fs, n_fft and lower_bin")   # reference: TypeError on None arithmetic
    lower = 1 if lower_bin is None else int(lower_bin)
    lower = max(lower, 1) if audio_format == 'foa' else lower
    p = _params(1, 0, fs=fs or 24000, n_fft=n_fft or 512, lower_bin=lower, upper_bin=lower + n_bins,
                audio_format=audio_format, is_tracking=is_tracking, n_hopframes=n_hopframes,
                cond_num=condition_number)
    lib = _native.lib()
    st = _stream()
    d_ref = torch.from_numpy(X.view(np.float64).reshape(n_bins, n_frames, 4, 2)).cuda()
    d_X = torch.empty((n_frames, 4, n_bins, 2), dtype=torch.float32, device='cuda')
    d_pow = torch.empty((n_frames, n_bins), dtype=torch.float64, device='cuda')
    _native.check(lib.salsa_spectrum_from_reference(_ptr(d_ref), _ptr(d_X), _ptr(d_pow), n_bins, n_frames, 4, st))
    d_mask = None
    if is_tracking:
        d_mask = torch.empty((n_frames, (n_bins + 31) // 32), dtype=torch.int32, device='cuda')
        _native.check(lib.salsa_tracker(_ptr(d_pow), _ptr(d_mask), 1, n_frames, n_bins, st))
    d_eig = torch.empty((3, n_frames, n_bins), dtype=torch.float32, device='cuda')
    _native.check(lib.salsa_eigenvector(ctypes.byref(p), _ptr(d_X), _ptr(d_mask), _ptr(d_eig), n_frames, st))
    return d_eig.permute(0, 2, 1).contiguous().cpu().numpy().astype(np.float64)


# ------------------------------------------------------------------------------------------------
# clip level (batched, device resident)
# ------------------------------------------------------------------------------------------------
class SalsaExtractor:
    """Batched SALSA extractor: the per-clip body of `extract_features`
    (salsa_feature_extraction.py:353-377) for (B, 4, N) clips resident in HBM.

    Arguments mirror `extract_features` (:265-270) and the `data:` block of the config yml."""

    def __init__(self, audio_format='foa', fs=24000, n_fft=512, hop_len=300, win_len=512, fmin_doa=50,
                 fmax_doa=9000, cond_num=5, n_hopframes=3, is_tracking=True, is_compress_high_freq=True,
                 stft_precision=64):
        if audio_format not in ('foa', 'mic'):
            raise ValueError('Unknown audio format {}'.format(audio_format))
        assert n_fft == 512 or n_fft == 256, 'only 256 or 512 fft is supported'
        self.audio_format, self.fs, self.n_fft, self.hop_len, self.win_len = audio_format, fs, n_fft, hop_len, win_len
        self.lower_bin, self.upper_bin = doa_bins(fs, n_fft, fmin_doa, fmax_doa)
        self.cond_num, self.n_hopframes, self.is_tracking = cond_num, n_hopframes, is_tracking
        self.is_compress_high_freq = is_compress_high_freq
        self.stft_precision = stft_precision
        self.freq_dim = (200 if n_fft == 512 else 100) if is_compress_high_freq else n_fft // 2
        self._workspace = None

    def _make_params(self, n_clips, n_samples):
        return _params(n_clips, n_samples, fs=self.fs, n_fft=self.n_fft, hop_len=self.hop_len, win_len=self.win_len,
                       lower_bin=self.lower_bin, upper_bin=self.upper_bin, audio_format=self.audio_format,
                       is_tracking=self.is_tracking, is_compress_high_freq=self.is_compress_high_freq,
                       n_hopframes=self.n_hopframes, cond_num=self.cond_num, stft_precision=self.stft_precision)

    def n_frames(self, n_samples):
        return 1 + n_samples // self.hop_len

    def extract(self, audio: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """audio (B, 4, N) float32 CUDA tensor -> (B, 7, T, freq_dim) float32 CUDA tensor.
        Asynchronous on the current stream."""
        _require_cuda()
        if audio.dim() != 3 or audio.shape[1] != 4 or audio.dtype != torch.float32 or not audio.is_cuda:
            raise ValueError('audio must be a CUDA float32 tensor of shape (B, 4, N)')
        audio = audio.contiguous()
        B, _, N = audio.shape
        p = self._make_params(B, N)
        lib = _native.lib()
        T = self.n_frames(N)
        if out is None:
            out = torch.empty((B, 7, T, self.freq_dim), dtype=torch.float32, device=audio.device)
        else:
            _check_out(out, (B, 7, T, self.freq_dim), audio.device)
        need = lib.salsa_workspace_bytes(ctypes.byref(p))
        if need == 0 and B > 0:
            _native.check(_native.SALSA_EINVAL)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != audio.device:
            self._workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=audio.device)
        with _native.device_of(audio) as st:
            _native.check(lib.salsa_extract(ctypes.byref(p), _ptr(audio), _ptr(out), _ptr(self._workspace),
                                            self._workspace.numel(), st))
        return out

    def extract_host(self, audio: np.ndarray, out: np.ndarray = None, clips_per_chunk: int = 16) -> np.ndarray:
        """Host buffers in and out ((B, 4, N) float32 or int16 PCM -> (B, 7, T, freq_dim) float32); the library
        streams chunks of clips through the GPU.  Accepts NumPy arrays or pinned CPU torch tensors."""
        _require_cuda()
        a = audio.numpy() if isinstance(audio, torch.Tensor) else audio
        if a.ndim != 3 or a.shape[1] != 4 or a.dtype not in (np.float32, np.int16) or not a.flags['C_CONTIGUOUS']:
            raise ValueError('audio must be a C-contiguous float32 (or int16 PCM) array of shape (B, 4, N)')
        B, _, N = a.shape
        T = self.n_frames(N)
        if out is None:
            out = np.empty((B, 7, T, self.freq_dim), dtype=np.float32)
        o = out.numpy() if isinstance(out, torch.Tensor) else out
        if o.shape != (B, 7, T, self.freq_dim) or o.dtype != np.float32 or not o.flags['C_CONTIGUOUS']:
            raise ValueError('out must be a C-contiguous float32 array of shape {}'.format((B, 7, T, self.freq_dim)))
        p = self._make_params(B, N)
        # int16: the wav files' own samples (what librosa.load divides by 32768, :353); half the bytes over PCIe
        fn = _native.lib().salsa_extract_host_pcm16 if a.dtype == np.int16 else _native.lib().salsa_extract_host
        _native.check(fn(ctypes.byref(p), ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(o.ctypes.data), int(clips_per_chunk)))
        return out


class SalsaLiteExtractor:
    """Batched SALSA-Lite / SALSA-IPD extractor: the per-clip body of
    salsa_lite_feature_extraction.py:94-123.  Arguments mirror its `extract_features` (:18-20)
    and the config yml (fmax_doa 2000, spectrogram cut-off 9 kHz hard-coded at :57)."""

    def __init__(self, feature_type='salsa_lite', fs=24000, n_fft=512, hop_len=300, win_len=512, fmin_doa=50,
                 fmax_doa=2000, fmax_spec=9000, stft_precision=64):
        assert feature_type in ['salsa_lite', 'salsa_ipd'], 'Invalid feature type {}'.format(feature_type)
        self.feature_type, self.fs, self.n_fft, self.hop_len, self.win_len = feature_type, fs, n_fft, hop_len, win_len
        self.lower_bin, self.upper_bin = doa_bins(fs, n_fft, fmin_doa, fmax_doa)
        self.cutoff_bin = int(math.floor(fmax_spec * n_fft / float(fs)))
        assert self.upper_bin <= self.cutoff_bin, \
            'Upper bin for spatial feature is higher than cutoff bin for spectrogram!'
        self.freq_dim = self.cutoff_bin - self.lower_bin
        self.stft_precision = stft_precision
        self.mode = _native.LITE_NIPD if feature_type == 'salsa_lite' else _native.LITE_IPD

    def _make_params(self, n_clips, n_samples):
        # win_len is read from the config but never used by the reference (salsa_lite_feature_extraction.py:44, :97-98:
        # librosa's default full-length Hann): accepted for the same signature, not forwarded
        return _params(n_clips, n_samples, fs=self.fs, n_fft=self.n_fft, hop_len=self.hop_len, win_len=self.n_fft,
                       lower_bin=self.lower_bin, upper_bin=self.upper_bin, audio_format='mic',
                       stft_precision=self.stft_precision)

    def n_frames(self, n_samples):
        return 1 + n_samples // self.hop_len

    def extract(self, audio: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """audio (B, 4, N) float32 CUDA tensor -> (B, 7, T, cutoff_bin - lower_bin) float32."""
        _require_cuda()
        if audio.dim() != 3 or audio.shape[1] != 4 or audio.dtype != torch.float32 or not audio.is_cuda:
            raise ValueError('audio must be a CUDA float32 tensor of shape (B, 4, N)')
        audio = audio.contiguous()
        B, _, N = audio.shape
        T = self.n_frames(N)
        if out is None:
            out = torch.empty((B, 7, T, self.freq_dim), dtype=torch.float32, device=audio.device)
        else:
            _check_out(out, (B, 7, T, self.freq_dim), audio.device)
        p = self._make_params(B, N)
        with _native.device_of(audio) as st:
            _native.check(_native.lib().salsa_lite_extract(ctypes.byref(p), self.cutoff_bin, self.mode, _ptr(audio),
                                                           _ptr(out), st))
        return out

    def extract_host(self, audio: np.ndarray, out: np.ndarray = None, clips_per_chunk: int = 16) -> np.ndarray:
        _require_cuda()
        a = audio.numpy() if isinstance(audio, torch.Tensor) else audio
        if a.ndim != 3 or a.shape[1] != 4 or a.dtype != np.float32 or not a.flags['C_CONTIGUOUS']:
            raise ValueError('audio must be a C-contiguous float32 array of shape (B, 4, N)')
        B, _, N = a.shape
        T = self.n_frames(N)
        if out is None:
            out = np.empty((B, 7, T, self.freq_dim), dtype=np.float32)
        o = out.numpy() if isinstance(out, torch.Tensor) else out
        if o.shape != (B, 7, T, self.freq_dim) or o.dtype != np.float32 or not o.flags['C_CONTIGUOUS']:
            raise ValueError('out must be a C-contiguous float32 array of shape {}'.format((B, 7, T, self.freq_dim)))
        p = self._make_params(B, N)
        _native.check(_native.lib().salsa_lite_extract_host(
            ctypes.byref(p), self.cutoff_bin, self.mode, ctypes.c_void_p(a.ctypes.data),
            ctypes.c_void_p(o.ctypes.data), int(clips_per_chunk)))
        return out


# ------------------------------------------------------------------------------------------------
# scaler (compute_scaler, salsa_feature_extraction.py:204-262)
# ------------------------------------------------------------------------------------------------
class FeatureScaler:
    """Streaming counterpart of `compute_scaler`: per-channel (0..3), per-frequency mean and standard deviation
    of the spectrogram channels over all frames of all clips (StandardScaler statistics, population variance).
    `partial_fit` takes (B, 7, T, F) CUDA feature batches as the extractors produce them; `finalize` returns what
    the reference stores in `<fmt>_feature_scaler.h5`: mean, std as float32 (4, 1, F).  With a process group the
    sums are all-reduced first (one collective of 4*F*2 + 1 float64 values)."""

    n_feature_channels = 4      # "hard coded number" (:224)

    def __init__(self):
        self._sums = None
        self._frames = 0

    def partial_fit(self, features: torch.Tensor):
        _require_cuda()
        if features.dim() != 4 or features.shape[1] < self.n_feature_channels or features.dtype != torch.float32 or not features.is_cuda:
            raise ValueError('features must be a CUDA float32 tensor (B, >=4, T, F)')
        features = features.contiguous()
        B, C, T, F = features.shape
        if self._sums is None:
            self._sums = torch.zeros((self.n_feature_channels, F, 2), dtype=torch.float64, device=features.device)
        elif self._sums.shape[1] != F:
            raise ValueError('feature dimension changed from {} to {}'.format(self._sums.shape[1], F))
        _native.same_device(features, self._sums)
        with _native.device_of(features) as st:
            _native.check(_native.lib().salsa_scaler_accumulate(_ptr(features), B, C, T, F, _ptr(self._sums), st))
        self._frames += B * T
        return self

    @staticmethod
    def reduce_statistics(sums: torch.Tensor, frames: float, group=None):
        """(sums float64 (4, F, 2), frames seen) of this process -> mean, std float32 (4, 1, F) over ALL processes."""
        import torch.distributed as dist
        sums = sums.clone()
        n = torch.tensor([float(frames)], dtype=torch.float64, device=sums.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(sums, group=group)
            dist.all_reduce(n, group=group)
        mean = sums[:, :, 0] / n
        var = torch.clamp(sums[:, :, 1] / n - mean * mean, min=0.0)
        return (mean[:, None, :].to(torch.float32).cpu().numpy(), torch.sqrt(var)[:, None, :].to(torch.float32).cpu().numpy())

    def finalize(self, group=None):
        if self._sums is None:
            raise RuntimeError('no features seen')
        return self.reduce_statistics(self._sums, self._frames, group)


def compute_scaler(feature_batches):
    """(mean, std), each float32 (4, 1, F), over an iterable of (B, 7, T, F) CUDA feature batches."""
    sc = FeatureScaler()
    for fb in feature_batches:
        sc.partial_fit(fb)
    return sc.finalize()
