"""salsa_b200 -- B200-native (sm_100a) SALSA feature extraction and SELD CRNN forward.

Drop-in for the hot path of thomeou/SALSA: `dataset/salsa_feature_extraction.py`,
`dataset/salsa_lite_feature_extraction.py` and `models.seld_models.SeldModel.forward`.
All compute runs in libsalsa_b200.so (hand-written CUDA behind a C ABI, include/salsa_b200.h);
this package is the thin host side.  There is no CPU fallback.
"""
from . import _native  # noqa: F401
from .features import (FeatureScaler, LinSpecIvExtractor, LogSpecGccExtractor, MagStftExtractor, SalsaExtractor, SalsaLiteExtractor, compute_scaler,  # noqa: F401
                       doa_bins, extract_normalized_eigenvector, stft)

from .crnn import PannResNet22, SeldDecoder, SeldModel  # noqa: F401
from . import crnn_ops  # noqa: F401
from . import augment  # noqa: F401
from . import optim  # noqa: F401
from .pipeline import SeldPipeline  # noqa: F401
from . import driver  # noqa: F401
from . import train  # noqa: F401

__version__ = '0.1.0'
