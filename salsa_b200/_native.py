"""ctypes binding of libsalsa_b200.so (C ABI declared in include/salsa_b200.h and include/salsa_crnn.h).

The library is the product; there is no Python / CPU fallback: loading fails loudly when the
shared object has not been built (`python -c "import __graft_entry__ as g; g.build()"`).
"""
import contextlib
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SALSA_B200_LIB: another build of the same library (A/B runs of kernel variants, scripts/build_variant.sh)
LIB_PATH = os.environ.get('SALSA_B200_LIB') or os.path.join(_HERE, 'libsalsa_b200.so')

SALSA_OK, SALSA_EINVAL, SALSA_ECUDA, SALSA_ENOMEM = 0, -1, -2, -3
FORMAT_FOA, FORMAT_MIC = 0, 1
LITE_NIPD, LITE_IPD = 0, 1


class SalsaParams(ctypes.Structure):
    """salsa_params_t"""
    _fields_ = [
        ('n_clips', ctypes.c_int32), ('n_chans', ctypes.c_int32), ('n_samples', ctypes.c_int32),
        ('fs', ctypes.c_int32), ('n_fft', ctypes.c_int32), ('hop_len', ctypes.c_int32),
        ('win_len', ctypes.c_int32), ('lower_bin', ctypes.c_int32), ('upper_bin', ctypes.c_int32),
        ('audio_format', ctypes.c_int32), ('is_tracking', ctypes.c_int32),
        ('is_compress_high_freq', ctypes.c_int32), ('n_hopframes', ctypes.c_int32),
        ('stft_precision', ctypes.c_int32), ('cond_num', ctypes.c_double),
        ('window', ctypes.POINTER(ctypes.c_double)),
    ]


class CrnnTensor(ctypes.Structure):
    """crnn_tensor_t"""
    _fields_ = [('name', ctypes.c_char_p), ('data', ctypes.POINTER(ctypes.c_float)), ('numel', ctypes.c_int64)]


class NativeError(RuntimeError):
    pass


_lib = None

_P = ctypes.POINTER(SalsaParams)
_vp, _i32, _sz, _u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t, ctypes.c_uint64

# name -> (restype, argtypes); must list every symbol include/salsa_b200.h declares
SIGNATURES = {
    'salsa_last_error': (ctypes.c_char_p, []),
    'salsa_version': (ctypes.c_char_p, []),
    'salsa_n_frames': (_i32, [_i32, _i32]),
    'salsa_feat_dim': (_i32, [_P]),
    'salsa_stft': (ctypes.c_int, [_P, _vp, _vp, _vp, _vp, _vp]),
    'salsa_tracker': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    'salsa_spectrum_from_reference': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    'salsa_eigenvector': (ctypes.c_int, [_P, _vp, _vp, _vp, _i32, _vp]),
    'salsa_workspace_bytes': (_sz, [_P]),
    'salsa_extract': (ctypes.c_int, [_P, _vp, _vp, _vp, _sz, _vp]),
    'salsa_lite_extract': (ctypes.c_int, [_P, _i32, _i32, _vp, _vp, _vp]),
    'salsa_linspec_iv_workspace_bytes': (_sz, [_P]),
    'salsa_linspec_iv': (ctypes.c_int, [_P, _vp, _vp, _vp, _sz, _vp]),
    'salsa_logspec_gcc_workspace_bytes': (_sz, [_P]),
    'salsa_logspec_gcc': (ctypes.c_int, [_P, _vp, _vp, _vp, _sz, _vp]),
    'salsa_extract_host': (ctypes.c_int, [_P, _vp, _vp, _i32]),
    'salsa_lite_extract_host': (ctypes.c_int, [_P, _i32, _i32, _vp, _vp, _i32]),
    'salsa_extract_host_pcm16': (ctypes.c_int, [_P, _vp, _vp, _i32]),
    'salsa_pcm16_to_float': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    'salsa_scaler_accumulate': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    'salsa_host_release': (ctypes.c_int, []),
    'salsa_launch_count': (_u64, [ctypes.c_int]),
    'salsa_profile_enable': (ctypes.c_int, [ctypes.c_int]),
    'salsa_profile_read': (ctypes.c_int, [_i32, _vp, _vp, _vp]),
    # include/salsa_crnn.h
    'crnn_load_weights': (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp]),
    'crnn_free_model': (ctypes.c_int, [_vp]),
    'crnn_workspace_bytes': (_sz, [_vp, _i32, _i32, _i32]),
    'crnn_forward': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    'crnn_model_option': (ctypes.c_int, [ctypes.c_char_p, _i32]),
    'crnn_conv2d': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_conv_wgrad': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_bn_train_forward': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int64, _i32, ctypes.c_float, ctypes.c_float, _i32, _vp, ctypes.c_uint32, ctypes.c_float, _vp]),
    'crnn_bn_train_backward': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int64, _i32, _i32, _vp, ctypes.c_uint32, ctypes.c_float, _vp]),
    'crnn_bn_train_forward_pool': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, ctypes.c_float, ctypes.c_float, _vp]),
    'crnn_bn_train_backward_pool': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    'crnn_conv_first': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_set_option': (ctypes.c_int, [ctypes.c_char_p, _i32]),
    'crnn_gemm': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_pack_input': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp]),
    'crnn_avgpool2': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_avgpool2_backward': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    'crnn_freq_mean': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    'crnn_gru_layer': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    'crnn_gru_layer_train': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    'crnn_gru_layer_backward': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    'crnn_head_finish': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    'crnn_decode_events': (ctypes.c_int, [_vp, _vp, _i32, _i32, ctypes.c_float, _vp, _vp, _vp, _vp]),
    'crnn_gather_time': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    'crnn_augment': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_cutout': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    'crnn_adam_step': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _i32, _vp]),
    'crnn_adam_hyper': (ctypes.c_int, [ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _i32, _vp]),
    'crnn_adam_step_hyper': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, _vp, _vp]),
    'crnn_seld_loss': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, _i32, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp]),
}


def lib():
    """The loaded shared library (loads on first use)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise NativeError(
                '{} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"`; '
                'salsa_b200 has no CPU fallback'.format(LIB_PATH))
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != SALSA_OK:
        msg = lib().salsa_last_error().decode('utf-8', 'replace')
        if rc == SALSA_EINVAL:
            raise ValueError(msg)
        raise NativeError('salsa_b200 error {}: {}'.format(rc, msg))


@contextlib.contextmanager
def device_of(t):
    """Makes the device of CUDA tensor `t` current for the duration of a native call and yields that device's current
    stream as the `void *stream` argument: the library launches on the CURRENT device (its tables, kernel attributes and
    staging buffers are per device), so a tensor on another GPU must switch the device first."""
    import torch
    if not t.is_cuda:
        raise ValueError('a CUDA tensor is required (salsa_b200 has no CPU fallback)')
    with torch.cuda.device(t.device):
        yield ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def same_device(*tensors):
    """Raises unless every given CUDA tensor (None entries are skipped) lives on one device."""
    devs = {t.device for t in tensors if t is not None}
    if len(devs) > 1:
        raise ValueError('tensors live on different devices: {}'.format(sorted(str(d) for d in devs)))


def profile_enable(on=True):
    lib().salsa_profile_enable(1 if on else 0)


def profile_read(max_entries=16):
    """{kernel name: (total_ms, launches)} since the previous read (synchronises)."""
    names = ctypes.create_string_buffer(32 * max_entries)
    ms = (ctypes.c_double * max_entries)()
    cnt = (ctypes.c_int64 * max_entries)()
    n = lib().salsa_profile_read(max_entries, ctypes.cast(names, ctypes.c_void_p), ctypes.cast(ms, ctypes.c_void_p),
                                 ctypes.cast(cnt, ctypes.c_void_p))
    out = {}
    for i in range(n):
        name = names.raw[32 * i:32 * (i + 1)].split(b'\0', 1)[0].decode()
        out[name] = (ms[i], cnt[i])
    return out
