"""ctypes binding of libgaiaseg_b200.so -- the C ABI declared in include/gaiaseg_b200.h.

The library holds every CUDA kernel of the hot path.  There is no fallback: if it is missing or the
device is not a B200 (sm_100), the product path raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgaiaseg_b200.so')


class GsError(RuntimeError):
    pass


class ConvGeom(Structure):
    """struct gs_conv_geom (include/gaiaseg_b200.h)."""
    _fields_ = [(n, c_int32) for n in ('N', 'H', 'W', 'Ho', 'Wo', 'Ci', 'Co', 'Ci_max', 'Co_max', 'kh', 'kw',
                                       'stride', 'pad', 'dil', 'x_ld', 'y_ld')]


class BnBwdFuse(Structure):
    """struct gs_bn_bwd_fuse (include/gaiaseg_b200.h)."""
    _fields_ = [('y', c_void_p), ('y_ld', c_int32), ('z', c_void_p), ('z_ld', c_int32), ('aff', c_void_p),
                ('relu', c_int32), ('sums', c_void_p)]


class SyncDesc(Structure):
    """struct gs_sync_desc (include/gaiaseg_b200.h): the SyncBN peer exchange a DynBN kernel runs itself."""
    _fields_ = [('peer_inboxes', c_void_p), ('rank', c_int32), ('world', c_int32), ('seq_dev', c_void_p),
                ('phase', c_int32), ('reserved', c_int32)]


class AugParams(Structure):
    """struct gs_aug_params (include/gaiaseg_b200.h): every pre-drawn decision of the train pipeline for one sample."""
    _fields_ = ([(n, c_int32) for n in ('H0', 'W0', 'new_h', 'new_w', 'crop_h', 'crop_w', 'out_h', 'out_w')] +
                [('box_y', c_int32 * 11), ('box_x', c_int32 * 11), ('flip', c_int32), ('has_brightness', c_int32),
                 ('brightness', c_float), ('contrast_first', c_int32), ('has_contrast', c_int32), ('contrast', c_float),
                 ('has_saturation', c_int32), ('saturation', c_float), ('has_hue', c_int32), ('hue', c_int32),
                 ('mean', c_float * 3), ('inv_std', c_float * 3), ('cat_max_ratio', c_float), ('ignore_index', c_int32)])


_P = c_void_p
_I = c_int32
_L = c_int64
_F = c_float
_D = c_double
_G = POINTER(ConvGeom)

# name -> (restype, argtypes); must list every symbol include/gaiaseg_b200.h declares
PROTOTYPES = {
    'gs_version': (_I, []),
    'gs_last_error': (c_char_p, []),
    'gs_device_check': (_I, []),
    'gs_launch_count': (_L, []),
    'gs_reset_launch_count': (None, []),
    'gs_debug_set_trace': (_I, [_P]),
    'gs_conv2d_fwd': (_I, [_G, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    'gs_conv2d_fwd_syncbn': (_I, [_G, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    'gs_conv2d_fwd_bn': (_I, [_G, _P, _P, _P, _P, _P, _D, _P, _P, _P, _P, _F, _F, _P, _P, _I, _I, _P, _I, _P]),
    'gs_conv2d_dgrad_workspace_bytes': (_L, [_G]),
    'gs_conv2d_dgrad': (_I, [_G, _P, _P, _P, _P, _I, _P, _P, _P]),
    'gs_conv2d_wgrad': (_I, [_G, _P, _P, _P, _P]),
    'gs_im2col_image': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'gs_bn_stats': (_I, [_P, _L, _I, _I, _P, _P]),
    'gs_bn_finalize': (_I, [_P, _D, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P]),
    'gs_bn_eval_affine': (_I, [_I, _P, _P, _P, _P, _F, _P, _P, _P]),
    'gs_bn_apply': (_I, [_P, _I, _P, _P, _P, _I, _I, _P, _I, _L, _I, _P]),
    'gs_bn_apply_train': (_I, [_P, _I, _P, _D, _P, _P, _P, _P, _F, _F, _P, _P, _I, _I, _P, _I, _L, _I, _P, _P]),
    'gs_bn_bwd': (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _D, _L, _I, _P, _I, _P, _I, _P, _P, _P, _P]),
    'gs_bn_bwd_reduce': (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _L, _I, _P, _P, _P]),
    'gs_bn_bwd_apply': (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _D, _L, _I, _P, _I, _P, _I, _P, _P, _P, _P]),
    'gs_affine_bwd': (_I, [_P, _I, _P, _I, _P, _L, _I, _P, _I, _P, _I, _P]),
    'gs_bn_bwd_param': (_I, [_P, _I, _P, _P, _I, _P]),
    'gs_maxpool3x3s2_fwd': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P]),
    'gs_maxpool3x3s2_bwd': (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    'gs_adaptive_avgpool_fwd': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    'gs_adaptive_avgpool_bwd': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    'gs_upsample_bf16_fwd': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    'gs_upsample_bf16_bwd': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    'gs_zero_channels': (_I, [_P, _I, _L, _I, _P]),
    'gs_copy_channels': (_I, [_P, _I, _P, _I, _L, _I, _P]),
    'gs_add_channels': (_I, [_P, _I, _P, _I, _L, _I, _P]),
    'gs_nchw_f32_to_nhwc_bf16': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    'gs_nhwc_bf16_to_nchw_f32': (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    'gs_scale_nc': (_I, [_P, _I, _P, _P, _I, _I, _L, _I, _P]),
    'gs_cast_f32_bf16': (_I, [_P, _I, _P, _I, _L, _I, _P]),
    'gs_colsum_f32': (_I, [_P, _I, _L, _I, _P, _P]),
    'gs_upsample_ce_record_bytes': (_L, [_I, _I, _I]),
    'gs_upsample_ce_fwd': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    'gs_upsample_ce_bwd': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _F, _P, _P, _I, _P]),
    'gs_upsample_argmax': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'gs_upsample_bilinear_f32': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    'gs_comm_inbox_bytes': (_L, [_I]),
    'gs_ipc_alloc': (_I, [_L, _P, _P]),
    'gs_ipc_open': (_I, [_P, _P]),
    'gs_ipc_close': (_I, [_P]),
    'gs_ipc_free': (_I, [_P]),
    'gs_syncbn_allreduce': (_I, [_P, _I, _P, _I, _I, _P, _P, _P, _P]),
    'gs_comm_flags_bytes': (_L, []),
    'gs_grad_allreduce': (_I, [_P, _L, _L, _P, _I, _I, _P, _P]),
    'gs_sgd_flat': (_I, [_P, _P, _P, _L, _P, _I, _P, _P]),
    'gs_aug_workspace_bytes': (_L, []),
    'gs_aug_choose_crop': (_I, [_P, _P, _P, _P, _P]),
    'gs_aug_fused': (_I, [_P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises GsError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                      f'(or `make -C gaia_seg_b200/csrc`). There is no CPU / PyTorch fallback for the hot path.')
    import torch  # noqa: F401  (makes libcudart.so.12 resident before dlopen)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gs_last_error().decode('utf-8', 'replace')


def check(code, what=''):
    if code != 0:
        raise GsError(f'{what} failed ({code}): {last_error()}')


_device_ok = False


def require_device():
    """Fail loudly unless the current CUDA device is a B200 (sm_100)."""
    global _device_ok
    if _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise GsError('gaia_seg_b200: no CUDA device -- the hot path has no CPU fallback')
    check(load().gs_device_check(), 'gs_device_check')
    _device_ok = True


# bench.py sets this to a list to time EVERY C-ABI call with CUDA events on the launching stream
# (entries: (entry point name, start event, end event)); None in normal operation.
PROFILE_CALLS = None


def call(name, *args):
    """Call an int-returning entry point and raise on error."""
    if PROFILE_CALLS is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        PROFILE_CALLS.append((name, e0, e1))
    else:
        rc = getattr(_lib or load(), name)(*args)
    if rc != 0:
        raise GsError(f'{name} failed ({rc}): {last_error()}')


def launch_count():
    return int(load().gs_launch_count())


def reset_launch_count():
    load().gs_reset_launch_count()
