"""ctypes binding of `libpemp_b200.so` (C ABI in `include/pemp_b200.h`).

The library is loaded on first use.  If it has not been built there is NO fallback: `lib()` raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# PEMP_B200_LIB selects a variant build of the SAME library (kernel-tuning experiments); there is still no fallback
LIB_PATH = os.environ.get("PEMP_B200_LIB") or os.path.join(_HERE, "libpemp_b200.so")
_lib = None

P, I, LL, F, SZ = c_void_p, c_int, c_longlong, c_float, c_size_t

# name -> (restype, argtypes); mirrors include/pemp_b200.h one to one
SIGNATURES = {
    "pemp_abi_version": (I, []),
    "pemp_strerror": (c_char_p, [I]),
    "pemp_check_device": (I, []),
    "pemp_mask_nearest": (I, [P, I, I, I, I, I, P, P]),
    "pemp_mask_nearest_labels": (I, [P, I, I, I, I, I, P, P]),
    "pemp_map_pool_workspace_bytes": (SZ, [I, I, I, I]),
    "pemp_map_pool_lowres": (I, [P, LL, P, P, LL, I, I, I, I, F, P, P, P, SZ, P]),
    "pemp_weighted_gap": (I, [P, P, I, I, I, P, P, SZ, P]),
    "pemp_meta_proto_attn_workspace_bytes": (SZ, [I, I, I, I, I]),
    "pemp_upsample_argmax_hist": (I, [P, I, I, I, I, I, P, P, P, I, P, P]),
    "pemp_meta_proto_attn_train": (I, [P, LL, P, P, P, LL, I, I, I, I, I, F, P, P, P, P, P, SZ, P]),
    "pemp_meta_proto_attn_bwd_workspace_bytes": (SZ, [I, I, I, I, I]),
    "pemp_meta_proto_attn_bwd": (I, [P, LL, P, P, P, LL, P, P, P, P, I, I, I, I, I, P, LL, P, P, SZ, P]),
    "pemp_cosine_match_bwd_workspace_bytes": (SZ, [I, I, I, I, I]),
    "pemp_cosine_match_bwd": (I, [P, LL, P, P, P, I, I, I, I, I, F, P, LL, P, P, P, SZ, P]),
    "pemp_cosine_sim_bwd": (I, [P, LL, P, P, P, I, I, I, I, I, F, P, LL, P, P, P, SZ, P]),
    "pemp_map_pool_lowres_bwd": (I, [P, P, LL, P, P, I, I, I, I, F, P, LL, P]),
    "pemp_canet_concat": (I, [P, LL, P, I, I, I, I, P, P]),
    "pemp_upsample_ce_workspace_bytes": (SZ, [I, I, I, I, I]),
    "pemp_upsample_ce": (I, [P, P, I, P, I, I, I, I, I, P, P, P, SZ, P]),
    "pemp_boundary_weight_workspace_bytes": (SZ, [I, I, I]),
    "pemp_boundary_weight": (I, [P, I, I, I, I, F, P, P, SZ, P]),
    "pemp_comm_workspace_bytes": (SZ, [I, I, I, I]),
    "pemp_comm_module": (I, [P, P, I, I, I, I, I, I, I, I, P, P, I, P, P, P, SZ, P]),
    "pemp_debug_mpa_path": (I, [I]),
    "pemp_debug_cosine_path": (I, [I]),
    "pemp_debug_pool_path": (I, [I]),
    "pemp_debug_bwd_path": (I, [I]),
    "pemp_meta_proto_attn": (I, [P, LL, P, P, P, LL, I, I, I, I, I, F, P, P, P, P, SZ, P]),
    "pemp_cosine_match": (I, [P, LL, P, P, I, I, I, I, I, F, P, P, P, P]),
    "pemp_upsample_argmax": (I, [P, I, I, I, I, I, P, P, P, P]),
    "pemp_bilinear_resize": (I, [P, I, I, I, I, I, P, P]),
    "pemp_nearest_resize_i64": (I, [P, I, I, I, I, I, P, P]),
    "pemp_map_pool_fullres_workspace_bytes": (SZ, [I, I, I, I, I]),
    "pemp_map_pool_fullres": (I, [P, LL, P, I, I, I, I, I, I, I, F, P, P, P, SZ, P]),
    "pemp_map_pool_fullres_labels_workspace_bytes": (SZ, [I, I, I, I, I, I, I]),
    "pemp_map_pool_fullres_labels": (I, [P, LL, P, I, I, I, I, I, I, I, F, P, P, P, SZ, P]),
    "pemp_bilinear_adjoint": (I, [P, I, I, I, I, I, P, P, P]),
    "pemp_panet_align_workspace_bytes": (SZ, [I, I, I, I, I, I, I, I]),
    "pemp_panet_align": (I, [P, LL, P, P, LL, P, LL, I, I, I, I, I, I, I, I, F, P, P, SZ, P]),
    "pemp_panet_align_labels": (I, [P, LL, P, P, LL, P, LL, I, I, I, I, I, I, I, I, F, P, P, SZ, P]),
    "pemp_prior_mask_workspace_bytes": (SZ, [I, I, I, I, I, I]),
    "pemp_prior_mask": (I, [P, P, P, I, I, I, I, I, I, P, P, P, SZ, P]),
    "pemp_iou_hist": (I, [P, P, P, I, LL, I, P, P]),
}


class PempError(RuntimeError):
    """A libpemp_b200 call returned a non-zero status."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pemp_b200.build` (or __graft_entry__.build()). "
                "pemp_b200 has no CPU / PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.pemp_abi_version() != 1:
            raise RuntimeError("libpemp_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def strerror(code):
    return lib().pemp_strerror(code).decode()


def check(code, what):
    """Raise like the reference does (ValueError for argument errors, `networks/pemp_stage1.py:36`)."""
    if code == 0:
        return
    msg = f"{what}: {strerror(code)} (status {code})"
    if code < 0:
        raise ValueError(msg)
    raise PempError(msg)
