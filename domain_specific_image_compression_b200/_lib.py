"""ctypes binding of libsic.so (include/sic.h).  No torch types cross this boundary: raw pointers, ints, stream handle.

The product has NO CPU fallback: if the library is missing (and cannot be built) import of any op raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsic.so")

# include/sic.h enums
QUANT_NONE, QUANT_ROUND, QUANT_NOISE_TENSOR, QUANT_NOISE_PHILOX = 0, 1, 2, 3
LIK_STUDENTT_DENSITY, LIK_GAUSSIAN, LIK_STUDENTT_CDFDIFF = 0, 1, 2
PARAM_BROADCAST, PARAM_SPATIAL, PARAM_CHANNEL = 0, 1, 2
DENSE_SERIAL, DENSE_PIPELINED = 0, 1

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_long
_z = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/sic.h declares (tests/test_abi.py checks this)
PROTOTYPES = {
    "sic_version": (_i, []),
    "sic_last_error": (ctypes.c_char_p, []),
    "sic_kernel_count": (_i, []),
    "sic_kernel_name": (ctypes.c_char_p, [_i]),
    "sic_kernel_registers": (_i, [ctypes.c_char_p]),
    "sic_bottleneck_workspace_bytes": (_z, [_i, _i, _i]),
    "sic_bottleneck_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "sic_bottleneck_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "sic_gdn_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "sic_gdn_bwd_workspace_bytes": (_z, [_i, _i, _i]),
    "sic_gdn_bwd_partials": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _z, _p]),
    "sic_gdn_bwd_fold": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "sic_gdn_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "sic_gdn_dense_fwd": (_i, [_p, _p, _p, _l, _i, _i, _p, _p]),
    "sic_gdn_dense_fwd_variant": (_i, [_p, _p, _p, _l, _i, _i, _p, _i, _p]),
    "sic_gdn_dense_bwd_part_rows": (_i, [_l, _i]),
    "sic_gdn_dense_bwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _i, _p]),
    "sic_gdn_dense_dgamma_workspace_bytes": (_z, [_l, _i]),
    "sic_gdn_dense_dgamma": (_i, [_p, _p, _l, _i, _p, _p, _z, _p]),
    "sic_conv0_gdn_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "sic_conv0_gdn_bwd_workspace_bytes": (_z, [_i, _i, _i, _i]),
    "sic_conv0_gdn_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _z, _p]),
    "sic_deconv_rgb_col2im": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "sic_deconv_rgb_im2col": (_i, [_p, _i, _i, _i, _p, _p]),
    "sic_hyper_tail_save_floats": (_z, [_i, _i, _i]),
    "sic_hyper_tail_scratch_floats": (_z, [_i, _i, _i]),
    "sic_hyper_tail_fwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, ctypes.c_float, ctypes.c_float, _p, _p, _p, _p]),
    "sic_hyper_tail_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, ctypes.c_float, ctypes.c_float] + [_p] * 11),
    "sic_ssim_tiles": (_l, [_i, _i]),
    "sic_ssim_fwd": (_i, [_p, _p, _i, _i, _i, ctypes.c_float, ctypes.c_float, _p, _p, _p, _p]),
    "sic_ssim_fwd_pool": (_i, [_p, _p, _i, _i, _i, ctypes.c_float, ctypes.c_float, _p, _p, _p, _p, _p, _p]),
    "sic_ssim_bwd_pool": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "sic_ssim_fwd_ex": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, ctypes.c_float, ctypes.c_float, _p, _p, _p, _p, _p, _p]),
    "sic_ssim_bwd_ex": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "sic_msssim_combine": (_i, [_p, _p, _p, _p, _i, _i, _p, _i, _p, _p, _p]),
    "sic_ssim_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "sic_rd_loss_fwd": (_i, [_p, _i, _p, _i, _p, _i, _l, ctypes.c_float, _p, _p, _p, _p, _p]),
    "sic_rd_loss_bwd": (_i, [_p, _p, _l, ctypes.c_float, _i, _i, _i, _p, _p, _p, _p]),
    "sic_bias_act_fwd": (_i, [_p, _p, _l, _i, _i, _p]),
    "sic_bias_grad_workspace_bytes": (_z, [_l, _i]),
    "sic_bias_act_bwd": (_i, [_p, _p, _l, _i, _p, _p, _p, _z, _p]),
    "sic_pack_flat": (_i, [_p, _p, _p, _i, _p, _p]),
    "sic_clip_adam_workspace_bytes": (_z, [_l]),
    "sic_clip_adam_step": (_i, [_p, _p, _p, _p, _l, _p] + [ctypes.c_float] * 7 + [_p, _p, _z, _p]),
    "sic_quantize_indices": (_i, [_p, _i, _l, _i, _i, _p, _p, _p, _p]),
    "sic_build_cdf_tables": (_i, [_i, _p, _p, _i, _i, _i, _p, _p, _i, _p, _p]),
    "sic_rans_encode": (_i, [_p, _p, _p, _i, _l, _l, _l, _i, _p, _l, _p, _p]),
    "sic_rans_encode_workspace_bytes": (_z, [_i, _l]),
    "sic_rans_encode_ws": (_i, [_p, _p, _p, _i, _l, _l, _l, _i, _p, _l, _p, _p, _z, _p]),
    "sic_rans_decode": (_i, [_p, _p, _p, _p, _i, _l, _l, _l, _i, _l, _p, _p, _p]),
    "sic_rans_encode_host": (_l, [_p, _l, _p, _i, _i, _l, _p, _l]),
    "sic_rans_decode_host": (_i, [_p, _l, _l, _p, _i, _i, _l, _p]),
}

_lock = threading.Lock()
_lib = None


class SicError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libsic.so, building it with nvcc first if the in-tree binary is absent.  Raises if neither works."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError here == ABI mismatch: fail loudly
            fn.restype, fn.argtypes = res, args
        if lib.sic_version() != 100:
            raise SicError(f"libsic.so version {lib.sic_version()} does not match the Python host (100); rebuild")
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sic_last_error()
        raise SicError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
