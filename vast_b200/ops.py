"""Thin Python wrappers, one per C-ABI entry point (include/vast_b200.h).  PyTorch is used for
device memory and streams only; every op runs the hand-written sm_100a kernels in
vast_b200/_C/libvast_b200.so and raises if that library (or a CUDA device) is missing."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, dtype_code, lib, ptr, require_cuda, stream_ptr


def _ws(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns >= 512-byte aligned blocks
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def kernel_timing(enable: bool) -> None:
    """Bracket every library kernel launch with CUDA events (benchmark aid)."""
    check(lib().vast_timing_enable(int(enable)), "timing_enable")


def kernel_timing_read(max_entries: int = 8192):
    """[(kernel name, milliseconds), ...] for the launches recorded since the last read (synchronises)."""
    import ctypes
    ms = (ctypes.c_float * max_entries)()
    names = ctypes.create_string_buffer(48 * max_entries)
    n = lib().vast_timing_read(ms, names, max_entries)
    if n < 0:
        check(n, "timing_read")
    return [(names.raw[48 * i:48 * i + 48].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n)]


# ------------------------------------------------------------------ generic GEMM (test / building block)
def gemm_nt(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """C[m, n] = alpha * sum_k a[m, k] b[n, k]; a, b bf16 or fp16 (same dtype), C fp32."""
    require_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1] and a.dtype == b.dtype
    assert a.stride(1) == 1 and b.stride(1) == 1
    m, k = a.shape
    n = b.shape[0]
    c = torch.empty(m, n, dtype=torch.float32, device=a.device)
    nbytes = lib().vast_gemm_nt_workspace_bytes(m, n, k)
    ws = _ws(nbytes, a.device)
    check(lib().vast_gemm_nt(ptr(a), a.stride(0), ptr(b), b.stride(0), dtype_code(a.dtype), m, n, k, float(alpha),
                             ptr(c), c.stride(0), ptr(ws), ws.numel(), stream_ptr()), "gemm_nt")
    return c


def gemm_nn(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """C[m, n] = alpha * sum_k a[m, k] b[k, n]; b row-major (MN-major tensor-core operand, no transpose);
    a, b bf16 or fp16, formats may differ; C fp32."""
    require_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[0]
    assert a.stride(1) == 1 and b.stride(1) == 1
    m, k = a.shape
    n = b.shape[1]
    c = torch.empty(m, n, dtype=torch.float32, device=a.device)
    ws = _ws(lib().vast_gemm_nt_workspace_bytes(m, n, k), a.device)
    check(lib().vast_gemm_nn(ptr(a), a.stride(0), dtype_code(a.dtype), ptr(b), b.stride(0), dtype_code(b.dtype), m, n, k,
                             float(alpha), ptr(c), c.stride(0), ptr(ws), ws.numel(), stream_ptr()), "gemm_nn")
    return c


# ------------------------------------------------------------------ feature build
def pool_concat(vision=None, audio=None, subtitle=None, vision_mode=0, audio_mode=1, out_dtype=None):
    """pool_vision/audio/text_for_contra + torch.cat(dim=1) (general_module.py:426-449, vast.py:269-275).
    vision [bs,n,tok,c], audio [bs,n,tok,c], subtitle [bs,tok,c]; mode 0 = token 0, 1 = token mean."""
    xs = [x for x in (vision, audio, subtitle) if x is not None]
    assert xs, "pool_concat: give at least one modality"
    require_cuda(*xs)
    dt = xs[0].dtype
    assert all(x.dtype == dt and x.is_contiguous() for x in xs)
    bs = xs[0].shape[0]
    width = sum(x.shape[-1] for x in xs)
    out = torch.empty(bs, width, dtype=out_dtype or dt, device=xs[0].device)
    v, a, s = vision, audio, subtitle
    check(lib().vast_pool_concat(
        ptr(v), *(v.shape[1:] if v is not None else (0, 0, 0)), int(vision_mode),
        ptr(a), *(a.shape[1:] if a is not None else (0, 0, 0)), int(audio_mode),
        ptr(s), *(s.shape[1:] if s is not None else (0, 0)),
        dtype_code(dt), bs, ptr(out), dtype_code(out.dtype), out.stride(0), stream_ptr()), "pool_concat")
    return out


def pool_concat_bwd(grad_out, vision_shape=None, audio_shape=None, subtitle_shape=None, vision_mode=0, audio_mode=1,
                    dtype=torch.float32):
    require_cuda(grad_out)
    g = grad_out.contiguous().float()
    bs = g.shape[0]
    dev = g.device
    gv = torch.empty(vision_shape, dtype=dtype, device=dev) if vision_shape is not None else None
    ga = torch.empty(audio_shape, dtype=dtype, device=dev) if audio_shape is not None else None
    gs = torch.empty(subtitle_shape, dtype=dtype, device=dev) if subtitle_shape is not None else None
    check(lib().vast_pool_concat_bwd(
        ptr(g), g.stride(0), bs,
        ptr(gv), *(vision_shape[1:] if gv is not None else (0, 0, 0)), int(vision_mode),
        ptr(ga), *(audio_shape[1:] if ga is not None else (0, 0, 0)), int(audio_mode),
        ptr(gs), *(subtitle_shape[1:] if gs is not None else (0, 0)),
        dtype_code(dtype), stream_ptr()), "pool_concat_bwd")
    return gv, ga, gs


def l2norm(x: torch.Tensor, eps: float = 1e-12, want_f32=True, out16: torch.Tensor | None = None, want_inv=False):
    """F.normalize(x, dim=-1).  Returns (y_f32 or None, inv_norm or None); optionally also writes the
    bf16 copy into `out16` (any row stride, e.g. the all-gather send slot)."""
    require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, dim = x.shape
    y = torch.empty(rows, dim, dtype=torch.float32, device=x.device) if want_f32 else None
    inv = torch.empty(rows, dtype=torch.float32, device=x.device) if want_inv else None
    if out16 is not None:
        assert out16.dtype == torch.bfloat16 and out16.stride(1) == 1 and out16.shape[0] == rows
    check(lib().vast_l2norm(ptr(x), dtype_code(x.dtype), rows, dim, x.stride(0), float(eps),
                            ptr(y), dim, ptr(out16), out16.stride(0) if out16 is not None else 0,
                            ptr(inv), stream_ptr()), "l2norm")
    return y, inv


def l2norm_bwd(grad_y, y, inv_norm, eps: float = 1e-12):
    require_cuda(grad_y, y, inv_norm)
    g = grad_y.contiguous().float()
    rows, dim = y.shape
    gx = torch.empty_like(y)
    check(lib().vast_l2norm_bwd(ptr(g), g.stride(0), ptr(y), y.stride(0), ptr(inv_norm), rows, dim, float(eps),
                                ptr(gx), gx.stride(0), stream_ptr()), "l2norm_bwd")
    return gx


def pair_sim(a_dtype, b_dtype, mode: str | None) -> int:
    """packed-operand mode of a GEMM on tensors of these dtypes: 'bf16' forces the plain cast (bf16-in / fp32-
    accumulate); otherwise values must survive exactly or to fp32 grade -- bf16 is exact in one bf16 term, fp16 needs two
    terms (3 products), fp32 three (6 products)."""
    if mode == "bf16":
        return _lib.SIM_BF16
    if torch.float32 in (a_dtype, b_dtype):
        return _lib.SIM_FP32X3
    if torch.float16 in (a_dtype, b_dtype):
        return _lib.SIM_FP32X2
    return _lib.SIM_BF16


def _pack(t: torch.Tensor, sim: int, as_query: bool) -> torch.Tensor:
    if sim == _lib.SIM_BF16 and t.dtype == torch.bfloat16 and t.is_contiguous():
        return t                                   # already the operand: no copy
    x = t if t.dtype in (torch.float32, torch.bfloat16) else t.float()
    return sim_pack_operand(x.contiguous(), sim, as_query)


def project_normalize(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, eps: float = 1e-12,
                      mode: str | None = None, out16: torch.Tensor | None = None, w_op: torch.Tensor | None = None):
    """F.normalize(x @ weight.T + bias, dim=-1) in one tensor-core kernel (vast_project_normalize).
    x [rows, K], weight [D, K]; mode None: fp32-grade for fp32 / fp16 inputs (bf16 splits), exact bf16 products for bf16
    inputs; 'bf16': inputs rounded to bf16 (bf16-in / fp32-accumulate).  out16: optional bf16 [rows, >= D] slot (any row
    stride) that receives the bf16 copy.  w_op: the weight already packed (cache it across calls).
    Returns (feat f32 [rows, D], inv_norm f32 [rows], w_op)."""
    require_cuda(x, weight)
    assert x.dim() == 2 and weight.dim() == 2 and x.shape[1] == weight.shape[1] and x.shape[1] % 8 == 0
    rows, dim_out = x.shape[0], weight.shape[0]
    sim = pair_sim(x.dtype, weight.dtype, mode)
    x_op = _pack(x, sim, True)
    if w_op is None or w_op.shape[1] != x_op.shape[1]:
        w_op = _pack(weight.detach(), sim, False)
    y = torch.empty(rows, dim_out, dtype=torch.float32, device=x.device)
    inv = torch.empty(rows, dtype=torch.float32, device=x.device)
    b = None if bias is None else bias.detach().float().contiguous()
    if out16 is not None:
        assert out16.dtype == torch.bfloat16 and out16.stride(1) == 1 and out16.shape[0] == rows and out16.shape[1] >= dim_out
    ws = _ws(lib().vast_project_normalize_workspace_bytes(rows, dim_out), x.device)
    check(lib().vast_project_normalize(ptr(x_op), ptr(w_op), rows, dim_out, x_op.shape[1], ptr(b), float(eps), ptr(y), dim_out,
                                       ptr(out16), out16.stride(0) if out16 is not None else 0, ptr(inv), ptr(ws),
                                       ws.numel(), stream_ptr()), "project_normalize")
    return y, inv, w_op


def match_head(cls: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, u0: torch.Tensor, u1: torch.Tensor, sum_u0: float,
               sum_u1: float, v0: float, v1: float, eps: float, mode: str | None = None, w1_op: torch.Tensor | None = None,
               want_logits: bool = False):
    """softmax(Linear2(LayerNorm(GELU(Linear1(cls)))))[:, 1] in one tensor-core kernel (vast_match_head); the LayerNorm
    and the 2-row Linear are folded into u_c / sum_u_c / v_c (see include/vast_b200.h).  Returns (score [b], logits
    [b, 2] | None, w1_op)."""
    require_cuda(cls, w1, b1, u0, u1)
    assert cls.dim() == 2 and cls.shape[1] == w1.shape[1] and cls.shape[1] % 8 == 0
    rows, hidden = cls.shape[0], w1.shape[0]
    sim = pair_sim(cls.dtype, w1.dtype, mode)
    c_op = _pack(cls, sim, True)
    if w1_op is None or w1_op.shape[1] != c_op.shape[1]:
        w1_op = _pack(w1.detach(), sim, False)
    score = torch.empty(rows, dtype=torch.float32, device=cls.device)
    logits = torch.empty(rows, 2, dtype=torch.float32, device=cls.device) if want_logits else None
    check(lib().vast_match_head(ptr(c_op), ptr(w1_op), rows, hidden, c_op.shape[1], ptr(b1), ptr(u0), ptr(u1), float(sum_u0),
                                float(sum_u1), float(v0), float(v1), float(eps), ptr(score), ptr(logits), stream_ptr()),
          "match_head")
    return score, logits, w1_op


def pack_pair(feat_t: torch.Tensor, feat_cond: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[bs, 2D] bf16 = (feat_t | feat_cond): the single all-gather payload."""
    require_cuda(feat_t, feat_cond)
    assert feat_t.shape == feat_cond.shape and feat_t.dtype == feat_cond.dtype
    ft, fc = feat_t.contiguous(), feat_cond.contiguous()
    bs, dim = ft.shape
    if out is None:
        out = torch.empty(bs, 2 * dim, dtype=torch.bfloat16, device=ft.device)
    assert out.is_contiguous() and out.shape == (bs, 2 * dim) and out.dtype == torch.bfloat16
    check(lib().vast_pack_pair(ptr(ft), ptr(fc), dtype_code(ft.dtype), bs, dim, dim, ptr(out), stream_ptr()), "pack_pair")
    return out


# ------------------------------------------------------------------ contrastive step
OMC_TWO_PASS = 1
OMC_SEPARATE_ROW_STATS = 2
OMC_ASSUME_IN_RANGE = 4
OMC_WORKSPACE_CLEAN = 8


def _omc_flags(two_pass: bool, separate_row_stats: bool | None, buffers: dict | None = None) -> int:
    import os
    if separate_row_stats is None:
        separate_row_stats = os.environ.get("VAST_OMC_SEPARATE_ROW_STATS", "0") == "1"
    assume = os.environ.get("VAST_OMC_ASSUME_IN_RANGE", "0") == "1"
    # a workspace handed back by an earlier call of the same shape was left clean by that step
    clean = buffers is not None and buffers.get("_clean", False)
    return (OMC_TWO_PASS if two_pass else 0) | (OMC_SEPARATE_ROW_STATS if separate_row_stats else 0) | \
        (OMC_ASSUME_IN_RANGE if assume else 0) | (OMC_WORKSPACE_CLEAN if clean else 0)


def omc_ws_cap_bytes() -> int:
    """workspace cap of one vast_omc_step call (VAST_OMC_MAX_WS_GB, default 32 GiB; the library refuses beyond it)."""
    import os
    try:
        gb = float(os.environ.get("VAST_OMC_MAX_WS_GB", "32"))
    except ValueError:
        gb = 32.0
    return int((gb if gb > 0 else 32.0) * (1 << 30))


def _omc_step_chunked(pack, bs, row_offset, contra_temp, label_smoothing, weight_floor, seed, offset, need_sample, need_grad,
                      debug_noise, want_lse, two_pass, step_counter, separate_row_stats):
    """The step on row chunks when the O(bs * n_total) Pt workspace would exceed the cap: every chunk is a block of the
    local rows with its own row_offset (the same kernels, the same Philox words -- they are keyed by the global row), the
    loss / d tau are the row-weighted means of the chunks', gradients scale by chunk / bs.  O(chunk * n_total) workspace."""
    n_total, dim = pack.shape[0], pack.shape[1] // 2
    cap = omc_ws_cap_bytes()
    rows = bs
    while rows > 128 and lib().vast_omc_workspace_bytes(rows, n_total, dim, int(need_sample), int(need_grad)) > cap:
        rows = max(128, (rows // 2 + 127) // 128 * 128)
    if lib().vast_omc_workspace_bytes(rows, n_total, dim, int(need_sample), int(need_grad)) > cap:
        raise RuntimeError(f"vast_b200.omc_step: even {rows} rows x {n_total} columns exceed the workspace cap "
                           f"({cap / 2 ** 30:.1f} GiB, VAST_OMC_MAX_WS_GB)")
    dev = pack.device
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    gtemp = torch.zeros(1, dtype=torch.float32, device=dev) if need_grad else None
    gc = torch.empty(bs, dim, dtype=torch.float32, device=dev) if need_grad else None
    gt = torch.empty(bs, dim, dtype=torch.float32, device=dev) if need_grad else None
    neg = torch.empty(2, bs, dtype=torch.int64, device=dev) if need_sample else None
    lse = torch.empty(2, bs, dtype=torch.float32, device=dev) if want_lse else None
    ctr0 = step_counter.clone() if step_counter is not None else None
    for r0 in range(0, bs, rows):
        r1 = min(r0 + rows, bs)
        if ctr0 is not None:
            step_counter.copy_(ctr0)           # every chunk of one step draws with the same counter
        o = omc_step(pack, r1 - r0, row_offset + r0, contra_temp, label_smoothing, weight_floor, seed, offset, need_sample,
                     need_grad, None if debug_noise is None else debug_noise[:, r0:r1].contiguous(), want_lse, None, two_pass,
                     step_counter, separate_row_stats)
        w = (r1 - r0) / bs
        loss += w * o["loss"]
        if need_grad:
            gc[r0:r1] = w * o["grad_cond"]
            gt[r0:r1] = w * o["grad_t"]
            gtemp += w * o["grad_temp"]
        if need_sample:
            neg[:, r0:r1] = o["neg_idx"]
        if want_lse:
            lse[:, r0:r1] = o["lse"]
    return dict(loss=loss, neg_idx=neg, grad_cond=gc, grad_t=gt, grad_temp=gtemp, lse=lse, _ws=(None, None), _clean=False)


def omc_step(pack: torch.Tensor, bs: int, row_offset: int, contra_temp, label_smoothing: float = 0.1,
             weight_floor: float = 1e-4, seed: int = 0, offset: int = 0, need_sample: bool = True,
             need_grad: bool = True, debug_noise: torch.Tensor | None = None, want_lse: bool = False,
             buffers: dict | None = None, two_pass: bool = False, step_counter: torch.Tensor | None = None,
             separate_row_stats: bool | None = None):
    """Fused OMC step (vast.py:405-440 + backward) on the packed, gathered features.
    Returns dict(loss[1], neg_idx[2,bs] | None, grad_cond, grad_t, grad_temp | None, lse | None).
    `buffers` (a dict returned by an earlier call with the same shapes/flags) re-uses outputs + workspace.
    two_pass: evaluate the logits twice (VAST_OMC_TWO_PASS) instead of the default single pass with the
    on-device fallback; debug_noise [2, bs, n_total] (Exp(1) variates) switches to the reference-literal
    per-element race argmax_j w_j / E_j for index-exact tests.
    step_counter: optional device int64[1]; the Philox offset used is offset + step_counter[0] and the step
    increments it (so a CUDA-graph replay of the step draws fresh noise).
    separate_row_stats: run the row statistics / hard-negative draw as their own kernel instead of inside the dQ
    GEMM's epilogue (identical results; default from VAST_OMC_SEPARATE_ROW_STATS, off)."""
    require_cuda(pack)
    assert pack.dtype == torch.bfloat16 and pack.is_contiguous() and pack.dim() == 2 and pack.shape[1] % 2 == 0
    n_total, dim = pack.shape[0], pack.shape[1] // 2
    dev = pack.device
    if debug_noise is not None:
        assert debug_noise.shape == (2, bs, n_total) and debug_noise.dtype == torch.float32 and debug_noise.is_contiguous()
    if buffers is None and lib().vast_omc_workspace_bytes(bs, n_total, dim, int(need_sample), int(need_grad)) > omc_ws_cap_bytes():
        return _omc_step_chunked(pack, bs, row_offset, contra_temp, label_smoothing, weight_floor, seed, offset, need_sample,
                                 need_grad, debug_noise, want_lse, two_pass, step_counter, separate_row_stats)
    if buffers is not None:
        loss, neg, gc, gt, gtemp, lse = (buffers[k] for k in ("loss", "neg_idx", "grad_cond", "grad_t", "grad_temp", "lse"))
        ws = buffers["_ws"][0]
        assert (neg is not None) == need_sample and (gc is not None) == need_grad and (lse is not None) == want_lse
        assert gc is None or gc.shape == (bs, dim)
    else:
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        neg = torch.empty(2, bs, dtype=torch.int64, device=dev) if need_sample else None
        gc = torch.empty(bs, dim, dtype=torch.float32, device=dev) if need_grad else None
        gt = torch.empty(bs, dim, dtype=torch.float32, device=dev) if need_grad else None
        gtemp = torch.empty(1, dtype=torch.float32, device=dev) if need_grad else None
        lse = torch.empty(2, bs, dtype=torch.float32, device=dev) if want_lse else None
        ws = _ws(lib().vast_omc_workspace_bytes(bs, n_total, dim, int(need_sample), int(need_grad)), dev)
    temp_dev = None
    if isinstance(contra_temp, torch.Tensor):  # device scalar: read by the kernels, no host sync
        require_cuda(contra_temp)
        temp_dev = contra_temp.detach().reshape(1).float()
        contra_temp = 0.0
    check(lib().vast_omc_step(ptr(pack), bs, n_total, dim, row_offset, float(contra_temp), ptr(temp_dev), float(label_smoothing),
                              float(weight_floor), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1),
                              ptr(step_counter), ptr(debug_noise), _omc_flags(two_pass, separate_row_stats, buffers), ptr(loss), ptr(neg), ptr(gc), ptr(gt),
                              ptr(gtemp), ptr(lse),
                              ptr(ws), ws.numel(), stream_ptr()), "omc_step")
    return dict(loss=loss, neg_idx=neg, grad_cond=gc, grad_t=gt, grad_temp=gtemp, lse=lse, _ws=(ws, temp_dev), _clean=True)


def omc_step_local(feat_t: torch.Tensor, feat_cond: torch.Tensor, contra_temp, label_smoothing: float = 0.1,
                   weight_floor: float = 1e-4, seed: int = 0, offset: int = 0, need_sample: bool = True,
                   need_grad: bool = True, debug_noise: torch.Tensor | None = None, want_lse: bool = False,
                   buffers: dict | None = None, two_pass: bool = False, step_counter: torch.Tensor | None = None,
                   separate_row_stats: bool | None = None):
    """The fused OMC step for ONE rank straight from the feature blocks feat_t / feat_cond [bs, D] (fp32 / bf16 /
    fp16): packing and the step's first kernel are a single pass over the features (vast_omc_step_local).
    Same returns as `omc_step` plus "pack" (the [bs, 2D] bf16 operand it wrote)."""
    require_cuda(feat_t, feat_cond)
    assert feat_t.shape == feat_cond.shape and feat_t.dtype == feat_cond.dtype and feat_t.dim() == 2
    ft, fc = feat_t.contiguous(), feat_cond.contiguous()
    bs, dim = ft.shape
    dev = ft.device
    if debug_noise is not None:
        assert debug_noise.shape == (2, bs, bs) and debug_noise.dtype == torch.float32 and debug_noise.is_contiguous()
    if buffers is None and lib().vast_omc_workspace_bytes(bs, bs, dim, int(need_sample), int(need_grad)) > omc_ws_cap_bytes():
        pack = pack_pair(ft, fc)               # beyond the workspace cap: packed once, then the step in row chunks
        out = omc_step(pack, bs, 0, contra_temp, label_smoothing, weight_floor, seed, offset, need_sample, need_grad, debug_noise,
                       want_lse, None, two_pass, step_counter, separate_row_stats)
        out["pack"] = pack
        return out
    if buffers is not None:
        loss, neg, gc, gt, gtemp, lse, pack = (buffers[k] for k in ("loss", "neg_idx", "grad_cond", "grad_t", "grad_temp",
                                                                   "lse", "pack"))
        ws = buffers["_ws"][0]
        assert (neg is not None) == need_sample and (gc is not None) == need_grad and (lse is not None) == want_lse
        assert pack.shape == (bs, 2 * dim)
    else:
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        neg = torch.empty(2, bs, dtype=torch.int64, device=dev) if need_sample else None
        gc = torch.empty(bs, dim, dtype=torch.float32, device=dev) if need_grad else None
        gt = torch.empty(bs, dim, dtype=torch.float32, device=dev) if need_grad else None
        gtemp = torch.empty(1, dtype=torch.float32, device=dev) if need_grad else None
        lse = torch.empty(2, bs, dtype=torch.float32, device=dev) if want_lse else None
        pack = torch.empty(bs, 2 * dim, dtype=torch.bfloat16, device=dev)
        ws = _ws(lib().vast_omc_workspace_bytes(bs, bs, dim, int(need_sample), int(need_grad)), dev)
    temp_dev = None
    if isinstance(contra_temp, torch.Tensor):
        require_cuda(contra_temp)
        temp_dev = contra_temp.detach().reshape(1).float()
        contra_temp = 0.0
    check(lib().vast_omc_step_local(ptr(ft), ptr(fc), dtype_code(ft.dtype), ft.stride(0), ptr(pack), bs, dim,
                                    float(contra_temp), ptr(temp_dev), float(label_smoothing), float(weight_floor),
                                    int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), ptr(step_counter),
                                    ptr(debug_noise), _omc_flags(two_pass, separate_row_stats, buffers), ptr(loss), ptr(neg), ptr(gc),
                                    ptr(gt), ptr(gtemp), ptr(lse), ptr(ws), ws.numel(), stream_ptr()), "omc_step_local")
    return dict(loss=loss, neg_idx=neg, grad_cond=gc, grad_t=gt, grad_temp=gtemp, lse=lse, pack=pack, _ws=(ws, temp_dev, ft, fc),
                _clean=True)


def local_step_ok(feat_t: torch.Tensor) -> bool:
    """vast_omc_step_local needs 16-byte vectors: D a multiple of 8 (contiguous rows are then 16-byte aligned)."""
    return feat_t.dim() == 2 and feat_t.shape[1] % 8 == 0 and feat_t.data_ptr() % 16 == 0


def gather_rows_concat3(ids_local, mask_local, ids_all, mask_all, cond_local, cond_all, neg_text, neg_cond):
    """vast.py:432-448: (input_ids_1 [3bs,L], attention_mask_1 [3bs,L], condition_feats [3bs,S,H])."""
    require_cuda(ids_local, mask_local, ids_all, mask_all, cond_local, cond_all, neg_text, neg_cond)
    bs, L = ids_local.shape
    n_total = ids_all.shape[0]
    il, ml = ids_local.contiguous(), mask_local.contiguous()
    ia, ma = ids_all.contiguous(), mask_all.contiguous()
    cl, ca = cond_local.contiguous(), cond_all.contiguous()
    assert il.dtype == torch.int64 and ml.dtype == torch.int64 and ia.dtype == torch.int64 and ma.dtype == torch.int64
    assert cl.dtype == ca.dtype and cl.shape[1:] == ca.shape[1:]
    row_bytes = cl[0].numel() * cl.element_size()
    ids_out = torch.empty(3 * bs, L, dtype=torch.int64, device=il.device)
    mask_out = torch.empty(3 * bs, L, dtype=torch.int64, device=il.device)
    cond_out = torch.empty((3 * bs,) + tuple(cl.shape[1:]), dtype=cl.dtype, device=cl.device)
    check(lib().vast_gather_rows_concat3(ptr(il), ptr(ml), ptr(ia), ptr(ma), L, ptr(cl), ptr(ca), row_bytes,
                                         ptr(neg_text.contiguous()), ptr(neg_cond.contiguous()), bs, n_total,
                                         ptr(ids_out), ptr(mask_out), ptr(cond_out), stream_ptr()), "gather_rows_concat3")
    return ids_out, mask_out, cond_out


def gather_rows_concat3_peer(ids_local, mask_local, ids_all, mask_all, cond_local, peer_ptrs, rows_per_rank: int, neg_text,
                             neg_cond):
    """`gather_rows_concat3` with the negative condition rows read from their owner ranks' symmetric-memory blocks
    (peer_ptrs: one device address per rank, rank order): vast_gather_rows_concat3_peer."""
    import ctypes
    require_cuda(ids_local, mask_local, ids_all, mask_all, cond_local, neg_text, neg_cond)
    bs, L = ids_local.shape
    il, ml, ia, ma = ids_local.contiguous(), mask_local.contiguous(), ids_all.contiguous(), mask_all.contiguous()
    cl = cond_local.contiguous()
    row_bytes = cl[0].numel() * cl.element_size()
    world = len(peer_ptrs)
    ids_out = torch.empty(3 * bs, L, dtype=torch.int64, device=il.device)
    mask_out = torch.empty(3 * bs, L, dtype=torch.int64, device=il.device)
    cond_out = torch.empty((3 * bs,) + tuple(cl.shape[1:]), dtype=cl.dtype, device=cl.device)
    arr = (ctypes.c_void_p * world)(*[int(p) for p in peer_ptrs])
    check(lib().vast_gather_rows_concat3_peer(ptr(il), ptr(ml), ptr(ia), ptr(ma), L, ptr(cl), arr, world, rows_per_rank, row_bytes,
                                              ptr(neg_text.contiguous()), ptr(neg_cond.contiguous()), bs, ptr(ids_out),
                                              ptr(mask_out), ptr(cond_out), stream_ptr()), "gather_rows_concat3_peer")
    return ids_out, mask_out, cond_out


def pull_row_grads(requests: torch.Tensor, peer_ptrs, bs: int, row0: int, like: torch.Tensor, base_grad: torch.Tensor | None = None):
    """Gradient of this rank's rows from the requesters' symmetric gradient blocks (vast_pull_row_grads): out[l] =
    base_grad[l] + sum over requests e with requests[e] == row0 + l of block[e // bs][e % bs].  `like` [bs, ...] gives
    shape / dtype of a block."""
    import ctypes
    require_cuda(requests, like)
    assert requests.dtype == torch.int64 and requests.is_contiguous()
    world = len(peer_ptrs)
    out = torch.empty((bs,) + tuple(like.shape[1:]), dtype=like.dtype, device=like.device)
    row_bytes = out[0].numel() * out.element_size()
    if base_grad is not None:
        base_grad = base_grad.contiguous()
        assert base_grad.shape == out.shape and base_grad.dtype == out.dtype
    ws = _ws(lib().vast_pull_row_grads_workspace_bytes(bs), like.device)
    arr = (ctypes.c_void_p * world)(*[int(p) for p in peer_ptrs])
    check(lib().vast_pull_row_grads(ptr(requests), arr, world, bs, row0, row_bytes, dtype_code(like.dtype), ptr(base_grad),
                                    ptr(out), ptr(ws), ws.numel(), stream_ptr()), "pull_row_grads")
    return out


# ------------------------------------------------------------------ retrieval scoring
SIM_BF16, SIM_FP32X3, SIM_FP32X2 = _lib.SIM_BF16, _lib.SIM_FP32X3, _lib.SIM_FP32X2
TOPK_MAX = 64


def sim_pack_operand(x: torch.Tensor, mode: int = SIM_BF16, as_query: bool = True) -> torch.Tensor:
    """16-bit tensor-core operand of x [rows, dim]: bf16 cast, or the 6-block 3-term split (fp32 grade)."""
    require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype in (torch.float32, torch.bfloat16)
    rows, dim = x.shape
    cols = lib().vast_sim_operand_cols(dim, mode)
    out = torch.empty(rows, cols, dtype=torch.bfloat16, device=x.device)
    check(lib().vast_sim_pack_operand(ptr(x), dtype_code(x.dtype), rows, dim, x.stride(0), mode, int(as_query),
                                      ptr(out), stream_ptr()), "sim_pack_operand")
    return out


def sim_topk(q_op: torch.Tensor, k_op: torch.Tensor, k: int, col_offset: int = 0, bounds_in: torch.Tensor | None = None,
             want_bounds: bool = False):
    """Streaming similarity + top-k on packed operands; returns sortable keys [n_q, k] (int64 storage).
    bounds_in [n_q] int32 (orderable score bits, 0 = none): proven lower bounds of every row's final k-th score over
    ALL columns of the problem -- the lists then start warm (vast_sim_topk_bounded) and may hold fewer than k keys.
    want_bounds: also return the bounds this call proved ([n_q] int32)."""
    require_cuda(q_op, k_op)
    assert q_op.dtype == torch.bfloat16 and k_op.dtype == torch.bfloat16 and q_op.is_contiguous() and k_op.is_contiguous()
    assert q_op.shape[1] == k_op.shape[1]
    n_q, cols = q_op.shape
    n_k = k_op.shape[0]
    keys = torch.empty(n_q, k, dtype=torch.int64, device=q_op.device)
    nbytes = lib().vast_sim_topk_workspace_bytes(n_q, n_k, cols, k)
    ws = _ws(nbytes, q_op.device)
    if bounds_in is None and not want_bounds:
        check(lib().vast_sim_topk(ptr(q_op), ptr(k_op), n_q, n_k, cols, k, col_offset, ptr(keys), ptr(ws), ws.numel(),
                                  stream_ptr()), "sim_topk")
        return keys
    if bounds_in is not None:
        require_cuda(bounds_in)
        assert bounds_in.dtype == torch.int32 and bounds_in.is_contiguous() and bounds_in.numel() == n_q
    bout = torch.empty(n_q, dtype=torch.int32, device=q_op.device) if want_bounds else None
    check(lib().vast_sim_topk_bounded(ptr(q_op), ptr(k_op), n_q, n_k, cols, k, col_offset, ptr(bounds_in), ptr(bout),
                                      ptr(keys), ptr(ws), ws.numel(), stream_ptr()), "sim_topk_bounded")
    return (keys, bout) if want_bounds else keys


def rank_of_gt(q: torch.Tensor, kk: torch.Tensor, gt_col: torch.Tensor, col_lo: int = 0, n_k: int | None = None,
               mode: int | None = None, delta_rel: float = 2.0 ** -11) -> torch.Tensor:
    """Streaming exact rank of the ground truth over key rows [col_lo, col_lo + n_k) of `kk` (vast_rank_of_gt):
    #{j : s_ij > s_i,gt or (s_ij == s_i,gt and j < gt_col[i])}, s = exact fp32-feature similarities.  int32 [n_q]."""
    require_cuda(q, kk, gt_col)
    assert q.dtype == torch.float32 and kk.dtype == torch.float32 and q.stride(1) == 1 and kk.stride(1) == 1
    n_q, dim = q.shape
    n_total = kk.shape[0]
    n_k = n_total - col_lo if n_k is None else n_k
    mode = SIM_FP32X2 if mode is None else mode
    gt = gt_col.int().contiguous()
    out = torch.zeros(n_q, dtype=torch.int32, device=q.device)
    if n_k <= 0 or n_q == 0:
        return out
    q_op = sim_pack_operand(q, mode, True)
    k_op = sim_pack_operand(kk[col_lo:col_lo + n_k], mode, False)
    cols = q_op.shape[1]
    ws = _ws(lib().vast_rank_of_gt_workspace_bytes(n_q, n_k, cols), q.device)
    check(lib().vast_rank_of_gt(ptr(q), q.stride(0), ptr(kk), kk.stride(0), n_q, n_total, dim, ptr(q_op), ptr(k_op), cols,
                                col_lo, n_k, ptr(gt), float(delta_rel), ptr(out), ptr(ws), ws.numel(), stream_ptr()),
          "rank_of_gt")
    return out


def topk_merge(keys_parts: torch.Tensor, k_out: int) -> torch.Tensor:
    """[parts, n_q, k_in] keys -> [n_q, k_out] (score desc, index asc)."""
    require_cuda(keys_parts)
    assert keys_parts.dim() == 3 and keys_parts.dtype == torch.int64 and keys_parts.is_contiguous()
    parts, n_q, k_in = keys_parts.shape
    out = torch.empty(n_q, k_out, dtype=torch.int64, device=keys_parts.device)
    check(lib().vast_topk_merge(ptr(keys_parts), parts, n_q, k_in, k_out, ptr(out), stream_ptr()), "topk_merge")
    return out


def topk_unpack(keys: torch.Tensor):
    require_cuda(keys)
    keys = keys.contiguous()
    vals = torch.empty(keys.shape, dtype=torch.float32, device=keys.device)
    idx = torch.empty(keys.shape, dtype=torch.int32, device=keys.device)
    check(lib().vast_topk_unpack(ptr(keys), keys.numel(), ptr(vals), ptr(idx), stream_ptr()), "topk_unpack")
    return vals, idx


def rescore_f64(q: torch.Tensor, kk: torch.Tensor, idx: torch.Tensor, key_offset: int = 0):
    """Exact fp64 scores of candidate lists, rows re-sorted by (score desc, index asc).  idx is modified in place."""
    require_cuda(q, kk, idx)
    assert q.dtype == torch.float32 and kk.dtype == torch.float32 and idx.dtype == torch.int32 and idx.is_contiguous()
    assert q.stride(1) == 1 and kk.stride(1) == 1
    n_q, k = idx.shape
    score = torch.empty(n_q, k, dtype=torch.float64, device=q.device)
    check(lib().vast_rescore_f64(ptr(q), q.stride(0), ptr(kk), kk.stride(0), n_q, q.shape[1], ptr(idx), k, key_offset,
                                 ptr(score), stream_ptr()), "rescore_f64")
    return idx, score


def exact_topk_rows(q, kk, rows_list, k, idx_out, score_out, col_offset: int = 0):
    require_cuda(q, kk, rows_list, idx_out, score_out)
    assert rows_list.dtype == torch.int32 and idx_out.dtype == torch.int32 and score_out.dtype == torch.float64
    check(lib().vast_exact_topk_rows(ptr(q), q.stride(0), ptr(kk), kk.stride(0), kk.shape[0], q.shape[1],
                                     ptr(rows_list), rows_list.numel(), k, col_offset, ptr(idx_out), ptr(score_out),
                                     stream_ptr()), "exact_topk_rows")


def dense_topk(score: torch.Tensor, k: int, axis: int = 1):
    """top-k of a materialised fp32 matrix, ties -> lower index.  axis=1: [n_rows,k]; axis=0: [k,n_cols]."""
    require_cuda(score)
    assert score.dtype == torch.float32 and score.dim() == 2 and score.stride(1) == 1
    n_rows, n_cols = score.shape
    shape = (n_rows, k) if axis == 1 else (k, n_cols)
    idx = torch.empty(shape, dtype=torch.int32, device=score.device)
    vals = torch.empty(shape, dtype=torch.float32, device=score.device)
    check(lib().vast_dense_topk(ptr(score), n_rows, n_cols, score.stride(0), k, axis, ptr(idx), ptr(vals), stream_ptr()),
          "dense_topk")
    return vals, idx


def dense_rank_of_gt(score: torch.Tensor, gt_row: torch.Tensor, gt_col: torch.Tensor, axis: int = 1) -> torch.Tensor:
    require_cuda(score, gt_row, gt_col)
    assert score.dtype == torch.float32 and score.stride(1) == 1
    gt_row, gt_col = gt_row.int().contiguous(), gt_col.int().contiguous()
    out = torch.empty(gt_row.numel(), dtype=torch.int32, device=score.device)
    check(lib().vast_dense_rank_of_gt(ptr(score), score.shape[0], score.shape[1], score.stride(0), axis, ptr(gt_row),
                                      ptr(gt_col), gt_row.numel(), ptr(out), stream_ptr()), "dense_rank_of_gt")
    return out


def bucket_by_video(text_idx: torch.Tensor, video_idx: torch.Tensor, n_videos: int):
    """Candidate pairs -> CSR by video: (offsets [n_videos+1] int32, texts [n_valid_pairs...] int32 ascending per video)."""
    require_cuda(text_idx, video_idx)
    t, v = text_idx.int().contiguous().reshape(-1), video_idx.int().contiguous().reshape(-1)
    n = t.numel()
    offsets = torch.empty(n_videos + 1, dtype=torch.int32, device=t.device)
    texts = torch.full((max(n, 1),), -1, dtype=torch.int32, device=t.device)
    ws = _ws(lib().vast_bucket_by_video_workspace_bytes(n, n_videos), t.device)
    check(lib().vast_bucket_by_video(ptr(t), ptr(v), n, n_videos, ptr(offsets), ptr(texts), ptr(ws), ws.numel(),
                                     stream_ptr()), "bucket_by_video")
    return offsets, texts[:n]


def scatter_scores(text_idx, video_idx, scores, out: torch.Tensor) -> torch.Tensor:
    require_cuda(text_idx, video_idx, scores, out)
    assert out.dtype == torch.float32 and out.stride(1) == 1
    t, v = text_idx.int().contiguous().reshape(-1), video_idx.int().contiguous().reshape(-1)
    s = scores.float().contiguous().reshape(-1)
    check(lib().vast_scatter_scores(ptr(t), ptr(v), ptr(s), t.numel(), ptr(out), out.stride(0), stream_ptr()),
          "scatter_scores")
    return out


def gemm_nt_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp32-grade a @ b.T for fp32 inputs on the bf16 tensor cores: 3-term bf16 splits, six leading cross
    products accumulated in fp32 (replaces the fp32 `torch.matmul` of evaluation_mm.py:223)."""
    return gemm_nt(sim_pack_operand(a.contiguous(), SIM_FP32X3, True), sim_pack_operand(b.contiguous(), SIM_FP32X3, False))
