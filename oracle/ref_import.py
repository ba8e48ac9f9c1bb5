"""Import the UNMODIFIED reference from /root/reference behind stub modules.

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.  Only usable where the reference
checkout is mounted (the build container); the GPU box does not have it, so nothing on the
`-m gpu` / smoke / bench path may call this.  Recipe: SURVEY.md appendix B.

Missing third-party modules are replaced by empty stubs (`fire`, `h5py`, `munch`), by this
repo's restatement (`librosa.stft` / `librosa.power_to_db` -> `oracle.stft`), or by the
thinnest possible shim (`pytorch_lightning.LightningModule` = `torch.nn.Module`).
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get('SALSA_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'dataset', 'salsa_feature_extraction.py'))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _ensure_path():
    if not available():
        raise RuntimeError('reference checkout not found at {}'.format(REFERENCE_ROOT))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def features_module():
    """-> the reference module `dataset.salsa_feature_extraction` (verbatim)."""
    from . import stft as _stft
    _ensure_path()
    _stub('fire', Fire=lambda *a, **k: None)
    _stub('h5py')
    try:
        import librosa  # noqa: F401  (use the real one if it ever becomes available)
    except ImportError:
        _stub('librosa', stft=_stft.stft, power_to_db=_stft.power_to_db, __oracle_shim__=True)
    import importlib
    return importlib.import_module('dataset.salsa_feature_extraction')


def other_features_module():
    """-> the reference module `dataset.feature_extraction` (verbatim): the IV / GCC-PHAT / mel families.  Only the
    classes that need nothing from librosa but `stft` / `power_to_db` can be instantiated (no `librosa.filters`)."""
    features_module()                      # installs the stubs
    import importlib
    return importlib.import_module('dataset.feature_extraction')


def transforms_module():
    """-> the reference module `utilities.transforms` (verbatim; NumPy only)."""
    _ensure_path()
    import importlib
    return importlib.import_module('utilities.transforms')


def models_module():
    """-> the reference package `models` (verbatim torch modules)."""
    import torch.nn as nn
    _ensure_path()

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    _stub('pytorch_lightning', LightningModule=LightningModule)
    _stub('h5py')
    pkg = _stub('metrics')
    pkg.__path__ = []
    for sub in ('dcase_utils', 'SELD2020_evaluation_metrics', 'SELD2021_evaluation_metrics'):
        setattr(pkg, sub, _stub('metrics.' + sub))
    import importlib
    return importlib.import_module('models')


def build_reference_seld_model(n_input_channels=7, n_classes=12, decoder_size=256):
    """Reference SeldModel(PannResNet22 + SeldDecoder bigru/avg) on CPU, eval mode not set."""
    models = models_module()
    enc = models.PannResNet22(n_input_channels=n_input_channels, p_dropout=0.0)
    dec = models.SeldDecoder(n_output_channels=enc.n_output_channels, n_classes=n_classes,
                             output_format='reg_xyz', decoder_type='bigru', freq_pool='avg',
                             decoder_size=decoder_size)
    tmp = tempfile.mkdtemp(prefix='salsa_ref_meta_')
    os.makedirs(os.path.join(tmp, 'metadata_dev'), exist_ok=True)
    return models.SeldModel(encoder=enc, decoder=dec, label_rate=10, feature_rate=80.0,
                            loss_weight=[0.3, 0.7], gt_meta_root_dir=tmp, output_format='reg_xyz',
                            eval_version='2021')


# ------------------------------------------------------------------------------------------
# Verbatim execution of the per-clip DRIVER BODIES.  The reference drivers cannot be called
# (they need wav directories, h5py and the removed np.int / np.float), so the relevant source
# lines are read from the reference checkout at run time, dedented and exec'd in a prepared
# namespace -- the reference text itself runs; nothing is copied into this repo.
# ------------------------------------------------------------------------------------------
_DRIVER_LINES = {
    # file, (setup first, last), (body first, last)   -- 1-based inclusive
    'salsa': ('dataset/salsa_feature_extraction.py', (290, 313), (355, 377)),
    'salsa_lite': ('dataset/salsa_lite_feature_extraction.py', (40, 66), (95, 123)),
}


def _source_block(path, first, last):
    import textwrap
    with open(os.path.join(REFERENCE_ROOT, path), 'r') as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent(''.join(lines))


def run_driver_body(kind: str, audio_input, data_cfg: dict, **params):
    """Run the reference's per-clip body on an in-memory clip.

    kind 'salsa': params cond_num, n_hopframes, is_tracking, is_compress_high_freq.
    kind 'salsa_lite': params feature_type ('salsa_lite' | 'salsa_ipd').
    Returns the reference's `audio_feature` cast to float32 as its h5 writer does
    (salsa_feature_extraction.py:380-382).
    """
    import numpy as np
    from . import stft as _stft
    ref = features_module()
    path, setup, body = _DRIVER_LINES[kind]

    class _NP:  # numpy with the two aliases numpy>=1.24 removed (used at :302-303, lite :52-53,58)
        int = int
        float = float

        def __getattr__(self, name):
            return getattr(np, name)

    ns = dict(np=_NP(), librosa=sys.modules['librosa'], cfg={'data': dict(data_cfg)},
              extract_normalized_eigenvector=ref.extract_normalized_eigenvector,
              audio_input=audio_input)
    if kind == 'salsa':
        ns.update(cond_num=params.get('cond_num', 5), n_hopframes=params.get('n_hopframes', 3),
                  is_tracking=params.get('is_tracking', True),
                  is_compress_high_freq=params.get('is_compress_high_freq', True))
    else:
        ns.update(feature_type=params.get('feature_type', 'salsa_lite'))
    exec(compile(_source_block(path, *setup), path + ':setup', 'exec'), ns)
    if kind == 'salsa':
        ns['stft_feature_extractor'] = ref.MagStftExtractor(
            n_fft=ns['n_fft'], hop_length=ns['hop_length'], win_length=ns['win_length'],
            is_compress_high_freq=ns['is_compress_high_freq'])
    exec(compile(_source_block(path, *body), path + ':body', 'exec'), ns)
    return np.asarray(ns['audio_feature']).astype(np.float32)
