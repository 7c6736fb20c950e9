"""ctypes binding of libocrb200.so (include/ocrb200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libocrb200.so")


class OcrbError(RuntimeError):
    pass


_lib = None

class ChainLinear(ctypes.Structure):
    """`ocrb_chain_linear` of include/ocrb200.h: one linear of an ocrb_skinny_chain_bf16 call."""
    _fields_ = [("X", c_void_p), ("ldx", c_int64), ("W", c_void_p), ("ldw", c_int64), ("D", c_void_p), ("ldd", c_int64),
                ("N", c_int32), ("K", c_int32), ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int64),
                ("epilogue", c_int32), ("eps", c_float), ("norm_w", c_void_p)]


class ChainAttention(ctypes.Structure):
    """`ocrb_chain_attention`: one layer's paged decode attention inside a plan."""
    _fields_ = [("qkv", c_void_p), ("ldqkv", c_int64), ("k_cache", c_void_p), ("v_cache", c_void_p), ("n_cache_pages", c_int32),
                ("block_table", c_void_p), ("max_pages", c_int32), ("ctx_len", c_void_p), ("page_size", c_int32),
                ("n_q", c_int32), ("n_kv", c_int32), ("hd", c_int32), ("cosT", c_void_p), ("sinT", c_void_p),
                ("scale", c_float), ("out", c_void_p), ("ldo", c_int64), ("split_ws", c_void_p), ("n_splits", c_int32)]


class ChainOp(ctypes.Structure):
    """`ocrb_chain_op`: kind 0 = linear (`lin`), 1 = attention (`att`)."""
    _fields_ = [("kind", c_int32), ("reserved", c_int32), ("lin", ChainLinear), ("att", ChainAttention)]


_P = c_void_p
_I = c_int32
_L = c_int64
_F = c_float

# name -> argtypes (all return int unless listed in _SPECIAL)
_SIGS = {
    "ocrb_levenshtein_batch": [_P, _P, _P, _P, _I, _I, _P, _P, _P],
    "ocrb_lcs_align_batch": [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "ocrb_rgb2gray_u8": [_P, _P, _I, _I, _I, _P],
    "ocrb_clahe_u8": [_P, _P, _I, _I, _I, _P, _P],
    "ocrb_adaptive_gauss_thresh_u8": [_P, _P, _I, _I, _I, _P],
    "ocrb_high_contrast_u8": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ocrb_binarize_u8": [_P, _P, _I, _I, _I, _I, _P],
    "ocrb_sharpen3x3_u8": [_P, _P, _I, _I, _I, _I, _P],
    "ocrb_remove_lines_mask_u8": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "ocrb_inpaint_telea_u8": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ocrb_nlm_denoise_u8": [_P, _P, _P, _I, _I, _I, _I, _P],
    "ocrb_denoise_tables_host": [_P, _P, _P, _P, _P],
    "ocrb_deskew_angle": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "ocrb_warp_affine_cubic_u8": [_P, _P, _I, _I, _I, _I, _P, _P],
    "ocrb_smart_resize_host": [_I, _I, _I, _L, _L, _P, _P],
    "ocrb_resize_bicubic_aa_u8": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ocrb_normalize_patchify": [_P, _P, _I, _I, _I, _I, _P, _I, _P],
    "ocrb_gemm_bf16": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _P, _L, _I, _P],
    "ocrb_skinny_gemm_bf16": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _P, _L, _I, _P, _F, _P, _P],
    "ocrb_skinny_chain_bf16": [_P, _I, _I, _P, _P],
    "ocrb_chain_plan_build": [_P, _I, _I, _P, _P, _P],
    "ocrb_chain_plan_run": [_P, _I, _I, _P, _P],
    "ocrb_rmsnorm_bf16": [_P, _L, _P, _P, _L, _I, _I, _F, _P],
    "ocrb_rope_vision": [_P, _I, _I, _I, _P, _P, _P],
    "ocrb_rope_text": [_P, _L, _P, _L, _I, _I, _I, _I, _P, _P, _P],
    "ocrb_attention_varlen": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    "ocrb_flash_attention_bf16": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "ocrb_kv_write_prefill": [_P, _L, _P, _L, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "ocrb_decode_attention": [_P, _L, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _F, _P, _L, _P, _I, _P],
    "ocrb_argmax_step": [_P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P],
    "ocrb_embed_gather": [_P, _P, _P, _I, _I, _P],
    "ocrb_rows_copy": [_P, _L, _P, _P, _L, _P, _I, _I, _P],
    "ocrb_residual_add_bf16": [_P, _L, _P, _L, _I, _I, _P],
    "ocrb_comm_ipc_handle": [_P, _P, _P],
    "ocrb_comm_ipc_open": [_P, _L, _P],
    "ocrb_allreduce_residual_bf16": [_P, _L, _P, _P, _I, _I, _P, _I, _I, _L, _P],
    "ocrb_skinny_rowparallel_tp_bf16": [_P, _L, _P, _L, _I, _I, _I, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P],
    "ocrb_tp_argmax_step": [_P, _L, _I, _I, _P, _P, _I, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P],
    "ocrb_decode_rope_table": [_P, _P, _P, _I, _I, _P, _P, _P],
}

EXPORTS = ["ocrb_version", "ocrb_last_error", "ocrb_launch_count", "ocrb_launch_count_reset",
           "ocrb_skinny_workspace_bytes", "ocrb_chain_workspace_bytes", "ocrb_chain_plan_bytes", "ocrb_inpaint_workspace_bytes",
           "ocrb_skinny_rowparallel_tp_was_fused"] + list(_SIGS)


def load():
    """Load the library (once).  Raises OcrbError when it is missing: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OcrbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C handwritten-ocr_b200/csrc`). This package has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.ocrb_version.restype = c_int32
    L.ocrb_last_error.restype = ctypes.c_char_p
    L.ocrb_launch_count.restype = c_uint64
    L.ocrb_launch_count_reset.restype = None
    L.ocrb_skinny_workspace_bytes.restype = c_int64
    L.ocrb_chain_workspace_bytes.restype = c_int64
    L.ocrb_chain_plan_bytes.restype = c_int64
    L.ocrb_chain_plan_bytes.argtypes = [c_int32]
    L.ocrb_inpaint_workspace_bytes.restype = c_int64
    L.ocrb_skinny_rowparallel_tp_was_fused.restype = c_int32
    L.ocrb_skinny_rowparallel_tp_was_fused.argtypes = []
    L.ocrb_inpaint_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    for name, args in _SIGS.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = c_int32
    _lib = L
    return L


_NVTX = os.environ.get("OCRB_NVTX", "0") == "1"


def call(name: str, *args):
    """One C-ABI call.  OCRB_NVTX=1 wraps every call in an NVTX range named after the entry point (visible to
    `ncu --nvtx` / Nsight Systems timelines); the engine adds phase ranges around them (engine.py)."""
    L = load()
    if _NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(L, name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        rc = getattr(L, name)(*args)
    if rc != 0:
        raise OcrbError(f"{name} failed ({rc}): {L.ocrb_last_error().decode(errors='replace')}")


class nvtx_range:
    """`with nvtx_range("vision"):` -- a phase range when OCRB_NVTX=1, nothing otherwise."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array; None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().ocrb_launch_count())


def launch_count_reset() -> None:
    load().ocrb_launch_count_reset()
