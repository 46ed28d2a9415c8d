"""ctypes loader for the C-ABI library (include/vast_b200.h).  Fails loudly: there is no CPU or
PyTorch fallback behind any op -- if the .so is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libvast_b200.so")

_lib = None

i64, i32, f32, u64, vp, sz = C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/vast_b200.h exactly (tests/test_abi.py checks it).
SIGNATURES = {
    "vast_version": (i32, []),
    "vast_last_error_string": (C.c_char_p, []),
    "vast_sm_count": (i32, []),
    "vast_timing_enable": (i32, [i32]),
    "vast_timing_read": (i32, [vp, vp, i32]),
    "vast_pool_concat": (i32, [vp, i64, i64, i64, i32, vp, i64, i64, i64, i32, vp, i64, i64, i32, i64, vp, i32, i64, vp]),
    "vast_pool_concat_bwd": (i32, [vp, i64, i64, vp, i64, i64, i64, i32, vp, i64, i64, i64, i32, vp, i64, i64, i32, vp]),
    "vast_l2norm": (i32, [vp, i32, i64, i64, i64, f32, vp, i64, vp, i64, vp, vp]),
    "vast_l2norm_bwd": (i32, [vp, i64, vp, i64, vp, i64, i64, f32, vp, i64, vp]),
    "vast_project_normalize_workspace_bytes": (sz, [i64, i64]),
    "vast_project_normalize": (i32, [vp, vp, i64, i64, i64, vp, f32, vp, i64, vp, i64, vp, vp, sz, vp]),
    "vast_match_head": (i32, [vp, vp, i64, i64, i64, vp, vp, vp, f32, f32, f32, f32, f32, vp, vp, vp]),
    "vast_pack_pair": (i32, [vp, vp, i32, i64, i64, i64, vp, vp]),
    "vast_pack_pair_push": (i32, [vp, vp, i32, i64, i64, i64, i64, vp, vp, i32, vp]),
    "vast_pack_pair_push_signal": (i32, [vp, vp, i32, i64, i64, i64, i64, vp, vp, i32, vp, i32, vp, vp]),
    "vast_wait_arrivals": (i32, [vp, i32, vp, vp]),
    "vast_omc_workspace_bytes": (sz, [i64, i64, i64, i32, i32]),
    "vast_omc_step": (i32, [vp, i64, i64, i64, i64, f32, vp, f32, f32, u64, u64, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "vast_omc_step_local": (i32, [vp, vp, i32, i64, vp, i64, i64, f32, vp, f32, f32, u64, u64, vp, vp, i32, vp, vp, vp, vp, vp, vp,
                                  vp, sz, vp]),
    "vast_gather_rows_concat3": (i32, [vp, vp, vp, vp, i64, vp, vp, i64, vp, vp, i64, i64, vp, vp, vp, vp]),
    "vast_gather_rows_concat3_peer": (i32, [vp, vp, vp, vp, i64, vp, vp, i32, i64, i64, vp, vp, i64, vp, vp, vp, vp]),
    "vast_pull_row_grads_workspace_bytes": (sz, [i64]),
    "vast_pull_row_grads": (i32, [vp, vp, i32, i64, i64, i64, i32, vp, vp, vp, sz, vp]),
    "vast_sim_operand_cols": (i64, [i64, i32]),
    "vast_sim_pack_operand": (i32, [vp, i32, i64, i64, i64, i32, i32, vp, vp]),
    "vast_sim_topk_workspace_bytes": (sz, [i64, i64, i64, i64]),
    "vast_sim_topk": (i32, [vp, vp, i64, i64, i64, i64, i64, vp, vp, sz, vp]),
    "vast_sim_topk_bounded": (i32, [vp, vp, i64, i64, i64, i64, i64, vp, vp, vp, vp, sz, vp]),
    "vast_rank_of_gt_workspace_bytes": (sz, [i64, i64, i64]),
    "vast_rank_of_gt": (i32, [vp, i64, vp, i64, i64, i64, i64, vp, vp, i64, i64, i64, vp, f32, vp, vp, sz, vp]),
    "vast_topk_merge": (i32, [vp, i64, i64, i64, i64, vp, vp]),
    "vast_topk_unpack": (i32, [vp, i64, vp, vp, vp]),
    "vast_rescore_f64": (i32, [vp, i64, vp, i64, i64, i64, vp, i64, i64, vp, vp]),
    "vast_exact_topk_rows": (i32, [vp, i64, vp, i64, i64, i64, vp, i64, i64, i64, vp, vp, vp]),
    "vast_dense_topk": (i32, [vp, i64, i64, i64, i64, i32, vp, vp, vp]),
    "vast_dense_rank_of_gt": (i32, [vp, i64, i64, i64, i32, vp, vp, i64, vp, vp]),
    "vast_bucket_by_video_workspace_bytes": (sz, [i64, i64]),
    "vast_bucket_by_video": (i32, [vp, vp, i64, i64, vp, vp, vp, sz, vp]),
    "vast_scatter_scores": (i32, [vp, vp, vp, i64, vp, i64, vp]),
    "vast_gemm_nt_workspace_bytes": (sz, [i64, i64, i64]),
    "vast_gemm_nt": (i32, [vp, i64, vp, i64, i32, i64, i64, i64, f32, vp, i64, vp, sz, vp]),
    "vast_gemm_nn": (i32, [vp, i64, i32, vp, i64, i32, i64, i64, i64, f32, vp, i64, vp, sz, vp]),
}

F32, BF16, F16 = 0, 1, 2
SIM_BF16, SIM_FP32X3, SIM_FP32X2 = 0, 1, 2


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"vast_b200: native library {LIB_PATH} is missing; build it with `python -m vast_b200.build` "
                "(there is no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().vast_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"vast_b200.{what} failed (status {status}): {msg}")


def dtype_code(t) -> int:
    import torch
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    if t == torch.float16:
        return F16
    raise TypeError(f"vast_b200: unsupported dtype {t}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vast_b200: tensors must live on a CUDA device (no CPU fallback)")
