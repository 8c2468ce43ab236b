"""ctypes binding of libpnpflow_sm100a.so (the C ABI declared in include/pnpflow_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C pnpflow_b200/csrc``.  There is no
fallback of any kind: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PNPF_LIB selects another BUILD of the same library (profiling / A-B builds under ab/); there is no other implementation
LIB_PATH = os.environ.get("PNPF_LIB") or os.path.join(_HERE, "libpnpflow_sm100a.so")

_lib = None


class UNetConfigC(C.Structure):
    _fields_ = [("input_channels", C.c_int), ("input_height", C.c_int), ("ch", C.c_int), ("num_levels", C.c_int),
                ("ch_mult", C.c_int * 8), ("num_res_blocks", C.c_int), ("num_attn_resolutions", C.c_int),
                ("attn_resolutions", C.c_int * 8)]


class OperatorC(C.Structure):
    _fields_ = [("kind", C.c_int), ("half_size", C.c_int), ("mask", C.c_void_p), ("sf", C.c_int),
                ("taps", C.c_void_p), ("ksize", C.c_int), ("scratch", C.c_void_p)]


OP_IDENTITY, OP_BOX, OP_MASK, OP_SR, OP_BLUR, OP_SR_BICUBIC = range(6)

# name -> (restype, argtypes); must list every symbol include/pnpflow_b200.h declares (tests check this)
_VP, _I, _LL, _F, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
SYMBOLS = {
    "pnpf_abi_version": (_I, []),
    "pnpf_last_error": (C.c_char_p, []),
    "pnpf_create": (_I, [C.POINTER(UNetConfigC), C.POINTER(_VP)]),
    "pnpf_destroy": (None, [_VP]),
    "pnpf_load_weight": (_I, [_VP, C.c_char_p, _VP, C.POINTER(C.c_int64), _I]),
    "pnpf_num_weights": (_I, [_VP]),
    "pnpf_weight_name": (C.c_char_p, [_VP, _I]),
    "pnpf_weight_shape": (_I, [_VP, _I, C.POINTER(C.c_int64 * 4), C.POINTER(_I)]),
    "pnpf_finalize_weights": (_I, [_VP]),
    "pnpf_workspace_bytes": (_SZ, [_VP, _I]),
    "pnpf_bind_workspace": (_I, [_VP, _VP, _SZ, _I]),
    "pnpf_unet_forward": (_I, [_VP, _VP, _VP, _VP, _I, _VP]),
    "pnpf_debug_num_ops": (_I, [_VP]),
    "pnpf_debug_op_name": (C.c_char_p, [_VP, _I]),
    "pnpf_debug_forward_partial": (_I, [_VP, _VP, _VP, _I, _I, _VP]),
    "pnpf_debug_read_op_output": (_I, [_VP, _I, _I, _VP, _SZ, C.POINTER(C.c_int * 3), _VP]),
    "pnpf_debug_op_info": (_I, [_VP, _I, C.POINTER(_I), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pnpf_debug_op_impl": (C.c_char_p, [_VP, _I]),
    "pnpf_profile_forward": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _I, _VP]),
    "pnpf_unet_flops_per_image": (C.c_double, [_VP]),
    "pnpf_unet_num_launches": (_I, [_VP]),
    "pnpf_apply_H": (_I, [C.POINTER(OperatorC), _VP, _VP, _I, _I, _I, _I, _VP]),
    "pnpf_apply_H_adj": (_I, [C.POINTER(OperatorC), _VP, _VP, _I, _I, _I, _I, _VP]),
    "pnpf_datafit_step": (_I, [C.POINTER(OperatorC), _VP, _VP, _VP, _F, _I, _I, _I, _I, _VP]),
    "pnpf_datafit_step_laplace": (_I, [C.POINTER(OperatorC), _VP, _VP, _VP, _F, _I, _I, _I, _I, _VP]),
    "pnpf_interp": (_I, [_VP, _VP, _F, _VP, _LL, _I, _VP]),
    "pnpf_push_accum": (_I, [_VP, _VP, _F, _I, _VP, _LL, _VP]),
    "pnpf_euler_step": (_I, [_VP, _VP, _F, _F, _I, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "pnpf_step": (_I, [_VP, C.POINTER(OperatorC), _I, _VP, _VP, _VP, _F, _F, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pnpf_conv2d_nhwc": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _I, _I, _I, _VP, _I, _VP, _VP, _VP, _I, _VP]),
    "pnpf_gn_conv2d_nhwc": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _I, _I, _VP, _I, _VP]),
    "pnpf_pack_subpixel_pair_weights": (_I, [_VP, _I, _I, _I, _VP]),
    "pnpf_gemm_nt": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "pnpf_attn_core_nhwc": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "pnpf_fold_subpixel_weights": (_I, [_VP, _I, _I, _I, _I, _VP]),
    "pnpf_upconv2x_nhwc": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _I, _VP, _VP]),
}


def load():
    """Load the shared library (once) and declare all prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(pnpflow_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pnpf_abi_version() != 1:
        raise RuntimeError("libpnpflow_sm100a.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().pnpf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"pnpflow_b200: {msg} (rc={rc})")


def stream_ptr(stream=None) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
