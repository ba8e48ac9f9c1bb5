"""Deterministic synthetic multichannel clips.

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.

clip i -> rng = default_rng(2021 + i); two sources, each 4th-order Butterworth band-pass
(300-6000 Hz) white noise, gated on/off over random sub-intervals, amplitude 0.1, mixed
through FOA encoding gains (channel order W, Y, Z, X) or through the per-mic fractional
delays of a 4.2 cm tetrahedral array, plus N(0, 1e-3^2) sensor noise; float32 (4, N).
"""
import numpy as np
import scipy.signal

MIC_AZI_ELE_DEG = ((45, 35), (-45, -35), (135, -35), (-135, 35))
MIC_RADIUS_M = 0.042
SOUND_SPEED = 343.0


def _unit(azi_deg, ele_deg):
    az, el = np.deg2rad(azi_deg), np.deg2rad(ele_deg)
    return np.array([np.cos(az) * np.cos(el), np.sin(az) * np.cos(el), np.sin(el)])


def _gated_source(rng, n, fs):
    sos = scipy.signal.butter(4, [300.0, 6000.0], btype='bandpass', fs=fs, output='sos')
    s = scipy.signal.sosfilt(sos, rng.standard_normal(n))
    s *= 0.1 / max(np.std(s), 1e-12)
    gate = np.zeros(n)
    n_seg = int(rng.integers(1, 4))
    for _ in range(n_seg):
        a = int(rng.integers(0, max(1, n - n // 8)))
        b = int(min(n, a + rng.integers(n // 8, max(n // 8 + 1, n // 2))))
        gate[a:b] = 1.0
    # 5 ms raised-cosine edges so the gates do not click
    ramp = max(2, int(0.005 * fs))
    k = np.hanning(2 * ramp + 1)
    gate = np.clip(np.convolve(gate, k / k.sum(), mode='same'), 0.0, 1.0)
    return s * gate


def make_clip(index: int, audio_format: str = 'foa', fs: int = 24000, seconds: float = 60.0,
              n_sources: int = 2) -> np.ndarray:
    """(4, int(fs*seconds)) float32."""
    rng = np.random.default_rng(2021 + index)
    n = int(round(fs * seconds))
    out = np.zeros((4, n))
    for _ in range(n_sources):
        s = _gated_source(rng, n, fs)
        az = rng.uniform(-180.0, 180.0)
        el = rng.uniform(-45.0, 45.0)
        if audio_format == 'foa':
            a, e = np.deg2rad(az), np.deg2rad(el)
            gains = np.array([1.0, np.sin(a) * np.cos(e), np.sin(e), np.cos(a) * np.cos(e)])
            out += gains[:, None] * s[None, :]
        elif audio_format == 'mic':
            S = np.fft.rfft(s)
            f = np.fft.rfftfreq(n, 1.0 / fs)
            u = _unit(az, el)
            for m, (ma, me) in enumerate(MIC_AZI_ELE_DEG):
                tau = -MIC_RADIUS_M * float(np.dot(u, _unit(ma, me))) / SOUND_SPEED
                out[m] += np.fft.irfft(S * np.exp(-2j * np.pi * f * tau), n)
        else:
            raise ValueError('Unknown audio format {}'.format(audio_format))
    out += 1e-3 * rng.standard_normal(out.shape)
    return out.astype(np.float32)
