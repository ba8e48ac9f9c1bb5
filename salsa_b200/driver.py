"""Script-level seam of the reference: `extract_features(data_config, ...)` with the reference's arguments, directory layout
and file contract (dataset/salsa_feature_extraction.py:265-385, :204-262; dataset/salsa_lite_feature_extraction.py:18-130),
on the batched CUDA extractors.

    <feature_dir>/<feature_type>/<format>/<description>/<split>/<clip>.h5      dataset 'feature' (7, T, F) float32
    <feature_dir>/<feature_type>/<format>/<description>/<format>_feature_scaler.h5   datasets 'mean', 'std' (4, 1, F) float32

What is different from the reference is only how the work is scheduled: clips are read as the 16-bit PCM the wav files hold
(the reference's `librosa.load(..., dtype=np.float32)` is sample / 32768, applied on the device), pushed through the GPU in
batches, and the scaler statistics are accumulated on the device while the features are still there instead of reading every
h5 file back.  File IO is pluggable: `reader(path) -> (C, N) int16 | float32` and `writer(path, {name: array})`; the defaults
use the standard library's `wave` and `h5py` (h5 IO itself is outside this package's scope: without h5py pass a writer).
`task` is the reference's: 'feature_scaler' (both), 'feature', or 'scaler' (the dev split's feature files are read back through
`feature_reader(path) -> (7, T, F)`, like compute_scaler does).
"""
import os
import shutil
import wave

import numpy as np
import torch

from . import _native
from .features import FeatureScaler, SalsaExtractor, SalsaLiteExtractor, doa_bins

__all__ = ['extract_features', 'extract_features_lite', 'feature_description', 'read_wav_pcm16', 'write_h5', 'read_h5_feature', 'pcm16_to_float']


def _load_config(data_config):
    if isinstance(data_config, dict):
        return data_config
    import yaml
    with open(data_config, 'r') as stream:
        return yaml.safe_load(stream)


def feature_description(cfg, feature_type='salsa', cond_num=5, is_tracking=True, is_compress_high_freq=True):
    """The directory name the reference derives from the configuration (salsa :315-321; lite :69)."""
    d = cfg['data']
    fmax_doa = int(np.min((d['fmax_doa'], d['fs'] // 2)))
    if feature_type != 'salsa':
        return '{}fs_{}nfft_{}nhop_{}fmaxdoa'.format(d['fs'], d['n_fft'], d['hop_len'], fmax_doa)
    desc = '{}fs_{}nfft_{}nhop_{}cond_{}fmaxdoa'.format(d['fs'], d['n_fft'], d['hop_len'], int(cond_num), fmax_doa)
    if not is_tracking:
        desc += '_notracking'
    if not is_compress_high_freq:
        desc += '_nocompress'
    return desc


def read_wav_pcm16(path, fs=None):
    """(C, N) int16 of a 16-bit PCM wav file (the dataset's format).  The reference resamples to `fs` in librosa.load; the
    dataset is already at the configured rate, anything else is refused rather than silently resampled differently."""
    with wave.open(path, 'rb') as w:
        if w.getsampwidth() != 2:
            raise ValueError('{}: {}-byte samples, expected 16-bit PCM'.format(path, w.getsampwidth()))
        if fs is not None and w.getframerate() != fs:
            raise ValueError('{}: sampling rate {} differs from the configured {}'.format(path, w.getframerate(), fs))
        data = np.frombuffer(w.readframes(w.getnframes()), dtype='<i2').reshape(-1, w.getnchannels())
    return np.ascontiguousarray(data.T)


def write_h5(path, arrays):
    try:
        import h5py
    except ImportError as exc:
        raise RuntimeError('h5py is needed for the default writer (pass writer=... to store the features another way)') from exc
    with h5py.File(path, 'w') as hf:
        for name, arr in arrays.items():
            hf.create_dataset(name, data=arr, dtype=np.float32)


def read_h5_feature(path):
    try:
        import h5py
    except ImportError as exc:
        raise RuntimeError("h5py is needed for the default feature reader of task='scaler' (pass feature_reader=...)") from exc
    with h5py.File(path, 'r') as hf:
        return hf['feature'][:]


def pcm16_to_float(pcm: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """int16 CUDA tensor -> float32 CUDA tensor, sample / 32768 (what librosa.load returns for a 16-bit wav)."""
    import ctypes
    if not (pcm.is_cuda and pcm.dtype == torch.int16):
        raise ValueError('pcm must be a CUDA int16 tensor')
    pcm = pcm.contiguous()
    if out is None:
        out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape == pcm.shape and out.device == pcm.device):
        raise ValueError('out must be a contiguous CUDA float32 tensor of the shape of pcm')
    with _native.device_of(pcm) as st:
        _native.check(_native.lib().salsa_pcm16_to_float(ctypes.c_void_p(pcm.data_ptr()), ctypes.c_void_p(out.data_ptr()), pcm.numel(), st))
    return out


def _run(cfg, feature_type, extractor, description, splits, task, batch_clips, reader, writer, device, feature_reader=None):
    feature_reader = feature_reader or read_h5_feature
    d = cfg['data']
    audio_format = d['format']
    root = os.path.join(cfg['feature_dir'], feature_type, audio_format, description)
    scaler = FeatureScaler()
    if task in ('feature_scaler', 'feature'):
        for split in splits:
            audio_dir = os.path.join(cfg['data_dir'], split)
            feature_dir = os.path.join(root, split)
            shutil.rmtree(feature_dir, ignore_errors=True)          # "Empty feature folder" (:337-339)
            os.makedirs(feature_dir, exist_ok=True)
            names = sorted(os.listdir(audio_dir))
            pending = []                                             # (name, audio) of equal length

            def flush():
                if not pending:
                    return
                audio = np.stack([a for _, a in pending])
                t = torch.from_numpy(audio).to(device)
                t = pcm16_to_float(t) if t.dtype == torch.int16 else t.float()
                feat = extractor.extract(t)
                if split.endswith('_dev') and task == 'feature_scaler':
                    scaler.partial_fit(feat)                        # compute_scaler's statistics (:204-262), dev split only
                host = feat.cpu().numpy()
                for (name, _), f in zip(pending, host):
                    writer(os.path.join(feature_dir, name.replace('wav', 'h5')), {'feature': f})
                pending.clear()

            for name in names:
                audio = reader(os.path.join(audio_dir, name))
                if audio.ndim != 2 or audio.shape[0] != 4:
                    raise ValueError('{}: expected 4 channels, got shape {}'.format(name, audio.shape))
                if pending and (audio.shape != pending[0][1].shape or audio.dtype != pending[0][1].dtype or len(pending) >= batch_clips):
                    flush()
                pending.append((name, audio))
            flush()
    if task == 'scaler':
        # compute_scaler as the reference runs it on its own (:204-262): the feature files of the dev split are read back;
        # the statistics are still accumulated on the device, one file batch at a time
        train_dir = os.path.join(root, audio_format + '_dev')
        pending = []

        def flush_scaler():
            if pending:
                scaler.partial_fit(torch.from_numpy(np.stack(pending)).to(device))
                pending.clear()

        for name in sorted(os.listdir(train_dir)):
            f = np.ascontiguousarray(feature_reader(os.path.join(train_dir, name)), dtype=np.float32)
            assert f.shape[0] == 7, 'only support n_channels = 7, got {}'.format(f.shape[0])
            if pending and (f.shape != pending[0].shape or len(pending) >= batch_clips):
                flush_scaler()
            pending.append(f)
        flush_scaler()
    if task in ('feature_scaler', 'scaler'):
        mean, std = scaler.finalize()
        writer(os.path.join(root, audio_format + '_feature_scaler.h5'), {'mean': mean, 'std': std})
    return root


def extract_features(data_config='configs/tnsse2021_salsa_feature_config.yml', cond_num: float = 5, n_hopframes: int = 3,
                     is_tracking: bool = True, is_compress_high_freq: bool = True, task: str = 'feature_scaler', *,
                     batch_clips: int = 16, reader=None, writer=None, device='cuda', feature_reader=None):
    """Drop-in for dataset/salsa_feature_extraction.py: extract_features (:265-385).  Returns the feature directory."""
    cfg = _load_config(data_config)
    d = cfg['data']
    audio_format = d['format']
    if audio_format == 'foa':
        splits = ['foa_dev', 'foa_eval']
    elif audio_format == 'mic':
        splits = ['mic_dev', 'mic_eval']
    else:
        raise ValueError('Unknown audio format {}'.format(audio_format))
    assert d['n_fft'] == 512 or d['n_fft'] == 256, 'only 256 or 512 fft is supported'
    ex = SalsaExtractor(audio_format=audio_format, fs=d['fs'], n_fft=d['n_fft'], hop_len=d['hop_len'], win_len=d['win_len'],
                        fmin_doa=d['fmin_doa'], fmax_doa=d['fmax_doa'], cond_num=cond_num, n_hopframes=n_hopframes,
                        is_tracking=is_tracking, is_compress_high_freq=is_compress_high_freq)
    desc = feature_description(cfg, 'salsa', cond_num, is_tracking, is_compress_high_freq)
    print('Feature description: {}'.format(desc))
    return _run(cfg, 'salsa', ex, desc, splits, task, batch_clips, reader or (lambda p: read_wav_pcm16(p, d['fs'])), writer or write_h5,
                torch.device(device), feature_reader)


def extract_features_lite(data_config='configs/tnsse2021_salsa_lite_feature_config.yml', feature_type: str = 'salsa_lite',
                          task: str = 'feature_scaler', *, batch_clips: int = 16, reader=None, writer=None, device='cuda',
                          feature_reader=None):
    """Drop-in for dataset/salsa_lite_feature_extraction.py: extract_features (:18-130)."""
    assert feature_type in ['salsa_lite', 'salsa_ipd'], 'Invalid feature type {}'.format(feature_type)
    cfg = _load_config(data_config)
    d = cfg['data']
    assert d['format'] == 'mic', 'SALSA-Lite and SALSA-IPD are only for MIC format!'
    ex = SalsaLiteExtractor(feature_type=feature_type, fs=d['fs'], n_fft=d['n_fft'], hop_len=d['hop_len'], win_len=d['win_len'],
                            fmin_doa=d['fmin_doa'], fmax_doa=d['fmax_doa'])
    desc = feature_description(cfg, feature_type)
    print('Feature description: {}'.format(desc))
    return _run(cfg, feature_type, ex, desc, ['mic_dev', 'mic_eval'], task, batch_clips,
                reader or (lambda p: read_wav_pcm16(p, d['fs'])), writer or write_h5, torch.device(device), feature_reader)


# same helper under the name the config parsing of the reference suggests
__all__.append('doa_bins')
