"""ctypes binding of libtmx.so (include/tmx.h).  No fallback of any kind: if the
CUDA library is missing or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libtmx.so')

TMX_ABI_VERSION = 6

# flags / enums (include/tmx.h)
CONV_LRELU, CONV_RESIDUAL, CONV_UP2_IN, CONV_UP2_OUT, CONV_HALO_REPLICATE, CONV_TORGB, CONV_XMERGE = 1, 2, 4, 8, 16, 32, 64
CONV_HALO_ZERO = 128
CONV_W_PER_SAMPLE = 256
ALGO_AUTO, ALGO_FFMA, ALGO_TC, ALGO_TC_K32 = 0, 1, 2, 3
BLEND_COPY, BLEND_MATTE, BLEND_LERP = 0, 1, 2
WGRAD_X_SLACK = 1

c_f32p = C.c_void_p
c_u16p = C.c_void_p


class ConvDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('Cin', C.c_int32), ('Cout', C.c_int32),
                ('k', C.c_int32), ('flags', C.c_uint32), ('algo', C.c_int32), ('wscale', C.c_float),
                ('lrelu_alpha', C.c_float), ('rgb_cout', C.c_int32), ('rgb_tanh', C.c_int32),
                ('rgb_wscale', C.c_float)]


class ConvIO(C.Structure):
    _fields_ = [('x_f32', C.c_void_p), ('x_hi', C.c_void_p), ('x_lo', C.c_void_p), ('w', C.c_void_p),
                ('w_hi', C.c_void_p), ('w_lo', C.c_void_p), ('bias', C.c_void_p), ('residual', C.c_void_p),
                ('y_f32', C.c_void_p), ('y_hi', C.c_void_p), ('y_lo', C.c_void_p), ('rgb_w', C.c_void_p),
                ('rgb_b', C.c_void_p), ('y_rgb', C.c_void_p)]


class GradDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32), ('C', C.c_int32), ('src_kind', C.c_int32),
                ('fold', C.c_int32), ('mask_kind', C.c_int32), ('phase_pack', C.c_int32), ('alpha', C.c_float),
                ('dbias_scale', C.c_float)]


class GradIO(C.Structure):
    _fields_ = [('g', C.c_void_p), ('add', C.c_void_p), ('y_mask', C.c_void_p), ('dz_hi', C.c_void_p),
                ('dz_lo', C.c_void_p), ('dz_f32', C.c_void_p), ('dbias', C.c_void_p)]


class BlendDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('C', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('H', C.c_int32),
                ('W', C.c_int32), ('K', C.c_int32), ('mode', C.c_int32), ('math_f32', C.c_int32),
                ('src_bcast', C.c_int32), ('src_reverse', C.c_uint32), ('pin_rows', C.c_uint64),
                ('pin_cols', C.c_uint64), ('c_off', C.c_int32), ('C_total', C.c_int32)]


class BlendIO(C.Structure):
    _fields_ = [('src', C.c_void_p * 4), ('idx_h', C.c_void_p * 4), ('idx_w', C.c_void_p * 4),
                ('ramp_h', C.c_void_p * 4), ('ramp_w', C.c_void_p * 4), ('t', C.c_void_p),
                ('out_nchw', C.c_void_p), ('out_nhwc', C.c_void_p)]


_I, _F, _P = C.c_int, C.c_float, C.c_void_p
_SIGNATURES = {
    'tmx_abi_version': (C.c_int, []),
    'tmx_last_error': (C.c_char_p, []),
    'tmx_create': (C.c_int, [_I, C.POINTER(_P)]),
    'tmx_destroy': (C.c_int, [_P]),
    'tmx_device_info': (C.c_int, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    'tmx_launch_count': (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    'tmx_conv2d_fwd': (C.c_int, [_P, C.POINTER(ConvDesc), C.POINTER(ConvIO), _P]),
    'tmx_conv_weights_prepare': (C.c_int, [_P, _P, _F, _I, _I, _I, _I, _I, _P, _P, _P]),
    'tmx_conv_weights_prepare_xmerge': (C.c_int, [_P, _P, _F, _I, _I, _P, _P, _P]),
    'tmx_mbstd_fwd': (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_dense_workspace_bytes': (C.c_int, [_I, _I, _I, C.POINTER(C.c_size_t)]),
    'tmx_dense_fwd': (C.c_int, [_P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _F, _P]),
    'tmx_split_halo_pack': (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'tmx_split_halo_unpack': (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    'tmx_fromrgb_fwd': (C.c_int, [_P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    'tmx_torgb_fwd': (C.c_int, [_P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_avgpool2_fwd': (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    'tmx_avgpool2_pack': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'tmx_nchw_to_nhwc': (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_nhwc_to_nchw': (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_conv2d_dgrad': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'tmx_conv2d_dgrad_gp': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, C.POINTER(GradDesc),
                                      C.POINTER(GradIO), C.POINTER(C.c_int), _P]),
    'tmx_conv_weights_transpose': (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    'tmx_conv2d_wgrad_workspace_bytes': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_size_t)]),
    'tmx_conv2d_wgrad': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _I, _P]),
    'tmx_conv_wgrad_unphase': (C.c_int, [_P, _P, _P, _I, _I, _P]),
    'tmx_torgb_bwd': (C.c_int, [_P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_fromrgb_bwd': (C.c_int, [_P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _I, _P]),
    'tmx_dense_bwd_input': (C.c_int, [_P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _F, _P]),
    'tmx_mbstd_bwd': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_loss_l1_grad': (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _F, _P]),
    'tmx_latent_gather_bwd': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, C.c_uint64, C.c_uint64, _I, _P]),
    'tmx_latent_noise_fwd': (C.c_int, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_latent_noise_bwd': (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_latent_gather_bwd_window': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, C.c_uint64,
                                               C.c_uint64, _I, _P]),
    'tmx_row_sum': (C.c_int, [_P, _P, _P, _I, _I, _F, _I, _I, _P]),
    'tmx_axpb': (C.c_int, [_P, _P, _P, C.c_int64, _F, _F, _P]),
    'tmx_mbstd_tangent': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_mbstd_curvature': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'tmx_dense_wgrad': (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _F, _P]),
    'tmx_scale_rows': (C.c_int, [_P, _P, _P, _P, _I, C.c_int64, _P]),
    'tmx_gp_coefficients': (C.c_int, [_P, _P, _P, _P, _I, _F, _F, _P]),
    'tmx_add_f32': (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    'tmx_kl_terms': (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, _F, _P]),
    'tmx_convert_output': (C.c_int, [_P, _P, _P, C.c_int64, _I, _I, _F, _F, _I, _I, _P]),
    'tmx_weighted_sum': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'tmx_tanh_f32': (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    'tmx_tanh_bwd': (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    'tmx_pixel_norm': (C.c_int, [_P, _P, _P, C.c_int64, _I, _F, _P]),
    'tmx_pixel_norm_bwd': (C.c_int, [_P, _P, _P, _P, C.c_int64, _I, _F, _P]),
    'tmx_bias_act': (C.c_int, [_P, _P, _P, _P, C.c_int64, _I, _I, _F, _P]),
    'tmx_grad_prepare': (C.c_int, [_P, C.POINTER(GradDesc), C.POINTER(GradIO), _P]),
    'tmx_nonfinite_check': (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    'tmx_adam_step': (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _F, _F, _F, _F, _F, _P, _P, _P]),
    'tmx_ema_update': (C.c_int, [_P, _P, _P, C.c_int64, _F, _P]),
    'tmx_nonfinite_mark': (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    'tmx_adam_update': (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _F, _F, _F, _F, _F, _P, _P, _P, _P]),
    'tmx_adam_advance': (C.c_int, [_P, _P, _F, _F, _P, _P, _P]),
    'tmx_vgg_preprocess': (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    'tmx_vgg_preprocess_bwd': (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    'tmx_gram_fwd': (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    'tmx_gram_fwd_tc_workspace_bytes': (C.c_int, [_P, _I, _I, _I, _I, C.POINTER(C.c_size_t)]),
    'tmx_gram_fwd_tc': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    'tmx_gram_sym_split': (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _P]),
    'tmx_gram_l1': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P, _I, _P]),
    'tmx_gram_bwd': (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    'tmx_window_copy': (C.c_int, [_P, _P, _P, C.c_int64, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    'tmx_latent_blend': (C.c_int, [_P, C.POINTER(BlendDesc), C.POINTER(BlendIO), _P]),
    'tmx_perm_indices_from_uniforms': (C.c_int, [C.POINTER(C.c_double), C.c_int64, _I, _I, _I,
                                                  C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_lib = None


def load():
    """Load libtmx.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'texturemixer_b200: %s is missing - build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(or `python texturemixer_b200/build.py`).  There is no CPU or PyTorch fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.tmx_abi_version() != TMX_ABI_VERSION:
        raise RuntimeError('texturemixer_b200: libtmx.so ABI %d != binding ABI %d - rebuild'
                           % (lib.tmx_abi_version(), TMX_ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().tmx_last_error().decode('utf-8', 'replace')
        raise RuntimeError('libtmx %s failed (%d): %s' % (what, rc, msg))
