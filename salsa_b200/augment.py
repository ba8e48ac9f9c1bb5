"""Training-time augmentations of SALSA feature batches on the GPU, with the reference's class names and random draws.

Host-side mirror of `utilities/transforms.py` for the transforms the reference composes for the SALSA features
(`dataset/datamodule.py:45-83`): `TfmapRandomSwapChannelFoa` (:368-437), `TfmapRandomSwapChannelMic` (:440-523),
`RandomShiftUpDownNp` (:286-320) and `CompositeCutout` (:257-283, MIC format).  The reference applies them per sample to NumPy arrays inside the DataLoader workers;
here the DRAWS are made per sample with NumPy exactly as the reference makes them (same calls in the same order, so a
seeded run picks the same augmentations) and the arithmetic runs once for the whole device batch in `crnn_augment`
(libsalsa_b200.so): index and sign permutations plus single float32 subtractions, bit-identical to the reference.
"""
import ctypes

import numpy as np
import torch

from . import _native

__all__ = ['TfmapRandomSwapChannelFoa', 'TfmapRandomSwapChannelMic', 'RandomShiftUpDownNp', 'CompositeCutout', 'BatchAugment']


class _Draw:
    def __init__(self, always_apply: bool = False, p: float = 0.5):
        self.always_apply = always_apply
        self.p = p

    def _fires(self) -> bool:
        # MapDataAugmentBase.__call__ / DataAugmentNumpyBase.__call__ (:346-352, :45-52)
        return True if self.always_apply else bool(np.random.rand() < self.p)


class TfmapRandomSwapChannelFoa(_Draw):
    """Random swap / negation of the x, y, z axes of FOA features and reg_xyz labels."""
    format = _native.FORMAT_FOA

    def __init__(self, always_apply: bool = False, p: float = 0.5, n_classes: int = 12):
        super().__init__(always_apply, p)
        self.n_classes = n_classes

    def draw(self) -> int:
        """Swap flags of one sample (bit i = m[i] of the reference's np.random.randint(2, size=(4,)), :409); 0 = skipped."""
        if not self._fires():
            return 0
        m = np.random.randint(2, size=(4,))
        return int(m[0] | (m[1] << 1) | (m[2] << 2) | (m[3] << 3))


class TfmapRandomSwapChannelMic(_Draw):
    """Random microphone swaps of MIC features (tetrahedral array) and reg_xyz labels."""
    format = _native.FORMAT_MIC

    def __init__(self, always_apply: bool = False, p: float = 0.5, n_classes: int = 12):
        super().__init__(always_apply, p)
        self.n_classes = n_classes

    def draw(self) -> int:
        if not self._fires():
            return 0
        m = np.random.randint(2, size=(3,))           # :485
        return int(m[0] | (m[1] << 1) | (m[2] << 2))


class RandomShiftUpDownNp(_Draw):
    """Random shift of the spectrogram up or down along frequency with reflect padding (all channels)."""

    def __init__(self, always_apply=False, p=0.5, freq_shift_range: int = None, direction: str = None, mode='reflect',
                 n_last_channels: int = 0):
        super().__init__(always_apply, p)
        if mode != 'reflect' or n_last_channels != 0:
            raise NotImplementedError('salsa_b200 implements the SALSA configuration: mode="reflect", n_last_channels=0')
        if direction not in (None, 'up', 'down'):
            raise ValueError('direction must be None, "up" or "down"')
        self.freq_shift_range = freq_shift_range
        self.direction = direction

    def draw(self, n_features: int):
        """(shift_len, direction flag) of one sample, (0, 0) = skipped (:298-305)."""
        if not self._fires():
            return 0, 0
        if self.freq_shift_range is None:
            self.freq_shift_range = int(n_features * 0.08)
        shift_len = int(np.random.randint(1, self.freq_shift_range, 1)[0])
        direction = np.random.choice(['up', 'down'], 1)[0] if self.direction is None else self.direction
        return shift_len, 0 if direction == 'up' else 1


class CompositeCutout(_Draw):
    """Random cutout / SpecAugment stripes / cutout holes (utilities/transforms.py:257-283 and its three parts :58-254), the
    transform the reference composes behind the frequency shift for MIC SALSA features (dataset/datamodule.py:76-82).
    Every variant is a list of rectangles with a fill value between the sample's min and max; the draws are made here in
    the reference's order, the fill happens on the device (`crnn_cutout`)."""
    MAX_RECTS = 8

    def __init__(self, always_apply: bool = False, p: float = 0.5, image_aspect_ratio: float = 1, n_zero_channels: int = None,
                 is_filled_last_channels: bool = True):
        super().__init__(always_apply, p)
        if not is_filled_last_channels:
            raise NotImplementedError('salsa_b200 implements is_filled_last_channels=True (the reference configuration)')
        self.n_zero_channels = n_zero_channels
        # RandomCutoutNp.__init__ (:77-85)
        self.s_l, self.s_h, self.r_1, self.r_2 = 0.02, 0.3, 0.3, 1 / 0.3
        if image_aspect_ratio > 1:
            self.r_1 = self.r_1 * image_aspect_ratio
        elif image_aspect_ratio < 1:
            self.r_2 = self.r_2 * image_aspect_ratio
        # RandomCutoutHoleNp.__init__ (:218-220) with its defaults
        self.n_max_holes, self.max_h_size, self.max_w_size = 8, int(np.max((8, 5))), int(np.max((8, 5)))

    def draw(self, n_frames: int, n_features: int):
        """Rectangles of one sample: list of (top, bottom, left, right, u) with exclusive ends, [] = skipped."""
        if not self._fires():
            return []
        img_h, img_w = n_frames, n_features
        choice = np.random.randint(0, 3, 1)[0]                       # CompositeCutout.apply (:277)
        if choice == 0:                                               # RandomCutoutNp.apply (:100-110)
            s = np.random.uniform(self.s_l, self.s_h) * img_h * img_w
            r = np.random.uniform(self.r_1, self.r_2)
            w = int(np.min((int(np.sqrt(s / r)), img_w - 1)))
            h = int(np.min((int(np.sqrt(s * r)), img_h - 1)))
            left = int(np.random.randint(0, img_w - w))
            top = int(np.random.randint(0, img_h - h))
            u = float(np.random.uniform(0.0, 1.0))                    # the draw of uniform(min_value, max_value)
            return [(top, top + h, left, left + w, u)]
        if choice == 1:                                               # SpecAugmentNp.apply (:160-194), one stripe each way
            time_max_width = int(np.max((1, int(0.15 * n_frames))))
            freq_max_width = int(np.max((1, int(0.2 * n_features))))
            dur = int(np.random.randint(1, time_max_width, 1)[0])
            start = int(np.random.randint(0, n_frames - dur, 1)[0])
            u1 = float(np.random.uniform(0.0, 1.0, 1)[0])
            rects = [(start, start + dur, 0, n_features, u1)]
            dur = int(np.random.randint(1, freq_max_width, 1)[0])
            start = int(np.random.randint(0, n_features - dur, 1)[0])
            u2 = float(np.random.uniform(0.0, 1.0, 1)[0])
            rects.append((0, n_frames, start, start + dur, u2))
            return rects
        rects = []                                                    # RandomCutoutHoleNp.apply (:233-252)
        for _ in range(self.n_max_holes):
            w, h = self.max_w_size, self.max_h_size
            left = int(np.random.randint(0, img_w - w))
            top = int(np.random.randint(0, img_h - h))
            u = float(np.random.uniform(0.0, 1.0))
            rects.append((top, top + h, left, left + w, u))
        return rects


class BatchAugment:
    """joint transform (one of the two channel swaps, or None) followed by the frequency shift (or None), the order of
    `SeldDataset.__getitem__` (dataset/dataloader.py:54-58), for a device batch."""

    def __init__(self, joint_transform=None, transform=None, cutout=None):
        """cutout: a `CompositeCutout` applied after the shift (the order of ComposeTransformNp at datamodule.py:76-82)."""
        self.joint, self.shift, self.cutout = joint_transform, transform, cutout

    def draw(self, batch: int, n_features: int, n_frames: int = None):
        """ops (B, 4) int32 = {format, swap flags, shift_len, direction}; sample by sample like the reference's loader.
        With a cutout configured (needs n_frames) returns (ops, cuts): cuts[b] = the rectangles of sample b."""
        ops = np.zeros((batch, 4), dtype=np.int32)
        cuts = []
        for b in range(batch):
            if self.joint is not None:
                ops[b, 0], ops[b, 1] = self.joint.format, self.joint.draw()
            if self.shift is not None:
                ops[b, 2], ops[b, 3] = self.shift.draw(n_features)
            if self.cutout is not None:
                cuts.append(self.cutout.draw(n_frames, n_features))
        return ops if self.cutout is None else (ops, cuts)

    def _apply_cutout(self, out: torch.Tensor, cuts):
        B, C, T, F = out.shape
        K = CompositeCutout.MAX_RECTS
        rects = np.zeros((B, K, 4), dtype=np.int32)
        u = np.zeros((B, K), dtype=np.float64)
        n = np.zeros((B,), dtype=np.int32)
        for b, cl in enumerate(cuts):
            if len(cl) > K:
                raise ValueError('at most {} rectangles per sample'.format(K))
            n[b] = len(cl)
            for k, (top, bottom, left, right, uk) in enumerate(cl):
                rects[b, k], u[b, k] = (top, bottom, left, right), uk
        d_rects, d_u, d_n = (torch.from_numpy(a).to(out.device) for a in (rects, u, n))
        minmax = torch.empty((B, 2), dtype=torch.float32, device=out.device)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        nz = self.cutout.n_zero_channels or 0
        with _native.device_of(out) as st:
            _native.check(_native.lib().crnn_cutout(vp(out), vp(d_rects), vp(d_n), vp(d_u), vp(minmax), B, C, T, F, nz, st))
        return out

    def __call__(self, x: torch.Tensor, y_sed: torch.Tensor, y_doa: torch.Tensor = None, ops: np.ndarray = None):
        """x (B, 7, T, F) float32 CUDA, y_doa (B, Ty, 3 n_classes) float32 CUDA or None -> (x_new, y_sed, y_doa_new)."""
        if not torch.cuda.is_available():
            raise RuntimeError('salsa_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        if x.dim() != 4 or x.shape[1] != 7 or x.dtype != torch.float32 or not x.is_cuda:
            raise ValueError('x must be a CUDA float32 tensor of shape (B, 7, T, F)')
        x = x.contiguous()
        B, _, T, F = x.shape
        cuts = None
        if ops is None:
            ops = self.draw(B, F, T)
            if self.cutout is not None:
                ops, cuts = ops
        elif isinstance(ops, tuple):
            ops, cuts = ops
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        if ops.shape != (B, 4) or (ops[:, 2] >= F).any() or (ops[:, 2] < 0).any():
            raise ValueError('ops must be (B, 4) with 0 <= shift_len < n_features')
        d_ops = torch.from_numpy(ops).to(x.device)
        out = torch.empty_like(x)
        y_out, Ty, n = None, 0, 0
        if y_doa is not None:
            if y_doa.dim() != 3 or y_doa.shape[0] != B or y_doa.shape[2] % 3 or y_doa.dtype != torch.float32 or not y_doa.is_cuda:
                raise ValueError('y_doa must be a CUDA float32 tensor of shape (B, Ty, 3 * n_classes)')
            y_doa = y_doa.contiguous()
            Ty, n = y_doa.shape[1], y_doa.shape[2] // 3
            y_out = torch.empty_like(y_doa)
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
        _native.same_device(x, y_doa)
        with _native.device_of(x) as st:
            _native.check(_native.lib().crnn_augment(vp(x), vp(out), vp(y_doa), vp(y_out), vp(d_ops), B, T, F, Ty, n, st))
        if cuts is not None and any(len(c) for c in cuts):
            self._apply_cutout(out, cuts)
        return out, y_sed, y_out
