"""The C-ABI library loads without a GPU and exports every symbol the headers under include/ declare."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, 'include'))):
        text = open(os.path.join(ROOT, 'include', fn)).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        names += re.findall(r'\b((?:salsa|crnn)_[a-z0-9_]+)\s*\(', text)
    return sorted(set(names))


def test_headers_declare_something():
    names = declared_symbols()
    assert 'salsa_extract' in names and 'crnn_conv2d' in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    from salsa_b200 import _native
    if not os.path.isfile(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), 'libsalsa_b200.so does not export {}'.format(name)
    # and the ctypes table binds exactly the declared set
    assert sorted(_native.SIGNATURES) == declared_symbols()


def test_version_and_frame_arithmetic_without_gpu():
    from salsa_b200 import _native
    lib = _native.lib()
    assert lib.salsa_version().startswith(b'salsa_b200')
    assert lib.salsa_n_frames(1440000, 300) == 4801
    assert lib.salsa_n_frames(24000, 300) == 81
    p = _native.SalsaParams()
    p.n_fft, p.is_compress_high_freq = 512, 1
    assert lib.salsa_feat_dim(ctypes.byref(p)) == 200
    p.is_compress_high_freq = 0
    assert lib.salsa_feat_dim(ctypes.byref(p)) == 256


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import numpy as np
    import salsa_b200
    with pytest.raises(RuntimeError):
        salsa_b200.MagStftExtractor(512, 300).extract(np.zeros((4, 2400), np.float32))
    with pytest.raises(RuntimeError):
        salsa_b200.extract_normalized_eigenvector(np.ones((2, 8, 4), complex), fs=24000, n_fft=512, lower_bin=1)


def test_doa_bins_host_logic():
    import salsa_b200
    assert salsa_b200.doa_bins(24000, 512, 50, 9000) == (1, 192)
    assert salsa_b200.doa_bins(24000, 512, 50, 4000) == (1, 85)
    assert salsa_b200.doa_bins(24000, 512, 50, 2000) == (1, 42)
    ex = salsa_b200.SalsaLiteExtractor()
    assert (ex.lower_bin, ex.upper_bin, ex.cutoff_bin, ex.freq_dim) == (1, 42, 192, 191)
    with pytest.raises(ValueError):
        salsa_b200.SalsaExtractor('xyz')


def test_interpolate_index_host_logic():
    import numpy as np
    import torch
    from salsa_b200 import crnn_ops
    for n_in, ratio in ((40, 2.0), (300, 2.0), (24, 0.5), (10, 1.5), (7, 0.3)):
        n_out = int(round(n_in * ratio))
        ref = torch.floor(torch.arange(n_out) / ratio).long().numpy()
        assert np.array_equal(crnn_ops.interpolate_index(n_in, ratio), ref)
