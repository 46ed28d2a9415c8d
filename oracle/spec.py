"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp64) of the reference's
cross-modal contrastive + retrieval-scoring hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker; the product
(vast_b200/) never imports anything under oracle/.

Parity status: the reference ships NO tests, golden vectors or fixtures for this
path, so the oracle is pinned by *running the reference
itself* in the authoring container: oracle/make_golden.py drives the reference's
own `VAST.forward_ret`, `refine_score_matrix`, `compute_metric_ret`,
`concat_all_gather` and `ddp_allgather` on seeded inputs and commits
inputs + outputs under tests/golden/; tests/test_oracle_golden.py checks every
function below against those vectors.

Every function cites the reference lines it restates (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# feature build: pool -> concat -> Linear -> L2 normalise
# ----------------------------------------------------------------------------


def pool_vision_for_contra(feature, encoder_type="evaclip"):
    """model/general_module.py:426-435. feature [b, n, tokens, c] -> [b, c].
    clip/evaclip: cls token (index 0) then frame mean; swin: token mean then frame mean."""
    f = np.asarray(feature, dtype=np.float64)
    if encoder_type.startswith("clip") or encoder_type.startswith("evaclip"):
        f = f[:, :, 0]
    elif encoder_type.startswith("swin"):
        f = f.mean(axis=2)
    return f.mean(axis=1)


def pool_text_for_contra(feature):
    """model/general_module.py:438-439. [b, L, c] -> cls token [b, c]."""
    return np.asarray(feature, dtype=np.float64)[:, 0]


def pool_audio_for_contra(feature, encoder_type="beats"):
    """model/general_module.py:441-449. ast: cls token; beats: token mean; then clip mean."""
    f = np.asarray(feature, dtype=np.float64)
    if encoder_type.startswith("ast"):
        f = f[:, :, 0]
    elif encoder_type.startswith("beats"):
        f = f.mean(axis=2)
    else:
        raise NotImplementedError
    return f.mean(axis=1)


def l2_normalize(x, eps=1e-12):
    """F.normalize(x, dim=-1) as used at model/vast.py:225-278: x / max(||x||_2, eps)."""
    x = np.asarray(x, dtype=np.float64)
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
    return x / np.maximum(n, eps)


def fuse_feature(pooled, weight, bias=None):
    """model/vast.py:269-279 (feat_vas; :254-266 for va/vs, :221-247 single modality):
    torch.cat(pooled, dim=1) -> Linear(weight [D, sum_dims], bias) -> F.normalize."""
    x = np.concatenate([np.asarray(p, dtype=np.float64) for p in pooled], axis=1)
    y = x @ np.asarray(weight, dtype=np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    return l2_normalize(y)


def match_head(cls, w1, b1, gamma, beta, w2, b2, eps=1e-12):
    """model/general_module.py:34-42 `Match_head.forward`: linear2(layernorm(gelu(linear1(cls)))) with the reference's
    erf GELU (:13-17) and torch LayerNorm (biased variance, eps inside the root); returns (logits [b, 2],
    score [b] = softmax(logits, 1)[:, 1] as used by compute_slice_scores, model/vast.py:378)."""
    from math import erf, sqrt
    x = np.asarray(cls, dtype=np.float64) @ np.asarray(w1, dtype=np.float64).T + np.asarray(b1, dtype=np.float64)
    g = x * 0.5 * (1.0 + np.vectorize(erf)(x / sqrt(2.0)))
    mu = g.mean(axis=1, keepdims=True)
    var = ((g - mu) ** 2).mean(axis=1, keepdims=True)
    h = (g - mu) / np.sqrt(var + eps) * np.asarray(gamma, dtype=np.float64) + np.asarray(beta, dtype=np.float64)
    z = h @ np.asarray(w2, dtype=np.float64).T + np.asarray(b2, dtype=np.float64)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return z, (e / e.sum(axis=1, keepdims=True))[:, 1]


# ----------------------------------------------------------------------------
# OMC / ITC loss (label-smoothed bidirectional softmax CE) and its backward
# ----------------------------------------------------------------------------


def _lse(z):
    m = z.max(axis=1, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(axis=1, keepdims=True)))[:, 0]


def omc_loss(feat_cond, feat_t, feat_t_all, feat_cond_all, contra_temp, rank=0,
             label_smoothing=0.1, grad_out=1.0):
    """model/vast.py:405-415 + autograd.

    sim_cond2t = feat_cond @ feat_t_all.T / temp ; sim_t2cond = feat_t @ feat_cond_all.T / temp
    targets = rank*bs + arange(bs) ; loss = (CE_eps(sim_cond2t) + CE_eps(sim_t2cond)) / 2
    CE_eps(z)_i = lse_i - (1-eps) z_iy - (eps/C) sum_j z_ij   (== F.cross_entropy(label_smoothing=eps))
    Gradients flow only to the local rows and to temp (gathered side is no_grad,
    utils/distributed.py:50).
    """
    fc = np.asarray(feat_cond, dtype=np.float64)
    ft = np.asarray(feat_t, dtype=np.float64)
    kt = np.asarray(feat_t_all, dtype=np.float64)
    kc = np.asarray(feat_cond_all, dtype=np.float64)
    tau = float(contra_temp)
    eps = float(label_smoothing)
    bs = fc.shape[0]
    C = kt.shape[0]
    y = rank * bs + np.arange(bs)
    out = {}
    total = 0.0
    dtau = 0.0
    for name, q, k in (("cond2t", fc, kt), ("t2cond", ft, kc)):
        z = (q @ k.T) / tau
        lse = _lse(z)
        zy = z[np.arange(bs), y]
        ce = lse - (1.0 - eps) * zy - (eps / C) * z.sum(axis=1)
        total += ce.mean()
        p = np.exp(z - lse[:, None])
        ysm = np.full_like(p, eps / C)
        ysm[np.arange(bs), y] += 1.0 - eps
        dz = (p - ysm) * (grad_out / (2.0 * bs))
        out["grad_" + ("cond" if name == "cond2t" else "t")] = (dz @ k) / tau
        dtau += -(dz * z).sum() / tau
        out["sim_" + name] = z
        out["lse_" + name] = lse
    out["loss"] = total / 2.0
    out["grad_temp"] = dtau
    return out


def hardneg_weights(sim, rank, floor=1e-4):
    """model/vast.py:423-427: softmax(sim, 1) + 1e-4, own block's diagonal zeroed."""
    z = np.asarray(sim, dtype=np.float64)
    bs = z.shape[0]
    w = np.exp(z - _lse(z)[:, None]) + floor
    w[np.arange(bs), rank * bs + np.arange(bs)] = 0.0
    return w


def exp_race_sample(weights, expo):
    """model/vast.py:430,437: torch.multinomial(w[b], 1).  ATen draws
    argmax_j w_j / E_j with E_j ~ Exp(1) (same distribution as inverse-CDF
    sampling); given the noise E explicitly the draw is a deterministic argmax
    (first index on ties)."""
    key = np.asarray(weights, dtype=np.float64) / np.asarray(expo, dtype=np.float64)
    return key.argmax(axis=1)


# Philox4x32-10 counter-based generator (Salmon et al. 2011); restated so the
# CPU can reproduce the exact uniforms the CUDA sampler draws.
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs uint32 arrays (broadcastable). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def sampler_bits(seed, offset, stream, rows, cols, row0=0):
    """The 32-bit Philox words the CUDA hard-negative sampler (vast_b200/csrc/omc.cu, EpiSoft)
    draws for CHUNK `col` (32 consecutive logit columns) of row `row`, direction `stream`
    (0 = cond2t, 1 = t2cond):
    counter = (col >> 2, row0 + row, offset lo32, (offset hi32 << 1) | stream),
    key = (seed lo32, seed hi32), word = col & 3.  row0 = rank * bs (global row)."""
    r = (np.arange(rows, dtype=np.uint32) + np.uint32(row0))[:, None]
    cg = (np.arange((cols + 3) // 4, dtype=np.uint32))[None, :]
    c3 = ((((int(offset) >> 32) << 1) | int(stream)) & 0xFFFFFFFF)
    out = philox4x32_10(cg, r, np.uint32(int(offset) & 0xFFFFFFFF), np.uint32(c3),
                        int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return np.stack(out, axis=-1).reshape(rows, -1)[:, :cols]


def sampler_expo(seed, offset, stream, rows, cols, row0=0):
    """Exp(1) race noise of the CUDA sampler: v = (x + 0.5) 2^-32, E = -log1p(-v)."""
    v = (sampler_bits(seed, offset, stream, rows, cols, row0).astype(np.float64) + 0.5) * (2.0 ** -32)
    return -np.log1p(-v)


def sampler_tail_units(seed, offset, stream, rows, row0=0):
    """The three uniforms (within-chunk, mixture, uniform-component) of the CUDA sampler's finalize step:
    Philox counter = (0xFFFFFFFF, row0 + row, offset lo32, (offset hi32 << 1) | stream), words 0..2,
    u = (x + 0.5) 2^-32."""
    r = np.arange(rows, dtype=np.uint32) + np.uint32(row0)
    c3 = ((((int(offset) >> 32) << 1) | int(stream)) & 0xFFFFFFFF)
    out = philox4x32_10(np.full(rows, 0xFFFFFFFF, dtype=np.uint32), r, np.uint32(int(offset) & 0xFFFFFFFF), np.uint32(c3),
                        int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return (np.stack(out[:3], axis=-1).astype(np.float64) + 0.5) * (2.0 ** -32)


def hardneg_hier_sample(sim, rank, chunk_expo, units, floor=1e-4, chunk=32):
    """The production hard-negative draw of vast_b200 (csrc/omc.cu: EpiSoft<false> + row finalize), an
    exact re-formulation of `torch.multinomial(softmax(sim) + floor with the positive zeroed, 1)`
    (model/vast.py:423-440):
      * w_j = p_j + floor (j != target) is the mixture  (1 - p_t) * [p_j / (1 - p_t)]  +  floor (N-1) * [1 / (N-1)];
      * the softmax component is drawn by the exponential race BETWEEN chunks of `chunk` columns
        (argmax_c  sum_{j in c} p_j / E_c: the minimum of independent E_j / p_j over a chunk is
        Exp(sum p_j) and its argmin is independent of it) followed by inverse CDF inside the chunk.
    sim [bs, N] logits; chunk_expo [bs, ceil(N/chunk)] Exp(1); units [bs, 3] uniforms.  Returns
    (index [bs], used_softmax_component [bs] bool, winning chunk [bs])."""
    z = np.asarray(sim, dtype=np.float64)
    bs, n = z.shape
    p = np.exp(z - _lse(z)[:, None])
    tcol = rank * bs + np.arange(bs)
    pt = p[np.arange(bs), tcol].copy()
    p[np.arange(bs), tcol] = 0.0
    nch = (n + chunk - 1) // chunk
    pad = np.zeros((bs, nch * chunk))
    pad[:, :n] = p
    cs = pad.reshape(bs, nch, chunk).sum(axis=2)
    key = cs / np.asarray(chunk_expo, dtype=np.float64)[:, :nch]
    win = key.argmax(axis=1)
    out = np.zeros(bs, dtype=np.int64)
    used_a = np.zeros(bs, dtype=bool)
    for i in range(bs):
        w_a = 1.0 - pt[i]
        w_b = floor * (n - 1)
        take_a = cs[i, win[i]] > 0 and units[i, 1] * (w_a + w_b) < w_a
        if n == 1:
            out[i] = -1
        elif take_a:
            seg = pad[i, win[i] * chunk:(win[i] + 1) * chunk]
            pre = np.cumsum(seg)
            hit = np.nonzero((seg > 0) & (pre >= units[i, 0] * pre[-1]))[0]
            j = hit[0] if hit.size else np.nonzero(seg > 0)[0][-1]
            out[i] = win[i] * chunk + j
            used_a[i] = True
        else:
            j = min(int(units[i, 2] * (n - 1)), n - 2)
            out[i] = j + 1 if j >= tcol[i] else j
    return out, used_a, win


def gather_negatives(condition_feats, condition_feats_collate, input_ids, attention_mask,
                     input_ids_collate, attention_mask_collate, neg_idx_t2cond, neg_idx_cond2t):
    """model/vast.py:429-448: rows picked by the sampled indices and the 3-way concats
    input_ids_1 = cat(ids, ids, ids_neg), attention_mask_1 likewise,
    condition_feats = cat(cond, cond_neg, cond)."""
    cond_neg = np.asarray(condition_feats_collate)[np.asarray(neg_idx_t2cond)]
    ids_neg = np.asarray(input_ids_collate)[np.asarray(neg_idx_cond2t)]
    att_neg = np.asarray(attention_mask_collate)[np.asarray(neg_idx_cond2t)]
    ids1 = np.concatenate([input_ids, input_ids, ids_neg], axis=0)
    att1 = np.concatenate([attention_mask, attention_mask, att_neg], axis=0)
    cond3 = np.concatenate([condition_feats, cond_neg, condition_feats], axis=0)
    return ids1, att1, cond3


# ----------------------------------------------------------------------------
# retrieval: similarity, top-k, metrics, ITM re-rank bookkeeping
# ----------------------------------------------------------------------------


def score_matrix(feat_t, feat_cond):
    """evaluation/evaluation_mm.py:223: feat_t @ feat_cond.T (no temperature)."""
    return np.asarray(feat_t, dtype=np.float64) @ np.asarray(feat_cond, dtype=np.float64).T


def score_matrix_f64_lane_order(feat_t, feat_cond, chunk=256):
    """Exact-mode scores in the *defined* summation order of the CUDA fp64 re-score
    kernel (vast_b200/csrc/retrieval.cu, rescore_f64): products of two fp32 values are
    exact in fp64; lane l of a warp accumulates k = l, l+32, ... in increasing k;
    lanes are combined by the xor-butterfly 16, 8, 4, 2, 1.  Bit-identical to the GPU."""
    a = np.asarray(feat_t, dtype=np.float32).astype(np.float64)
    b = np.asarray(feat_cond, dtype=np.float32).astype(np.float64)
    nt, d = a.shape
    nv = b.shape[0]
    dp = (d + 31) // 32 * 32
    if dp != d:
        a = np.pad(a, ((0, 0), (0, dp - d)))
        b = np.pad(b, ((0, 0), (0, dp - d)))
    out = np.empty((nt, nv), dtype=np.float64)
    a3 = a.reshape(nt, dp // 32, 32)
    b3 = b.reshape(nv, dp // 32, 32)
    for i0 in range(0, nt, chunk):
        ai = a3[i0:i0 + chunk]
        acc = np.zeros((ai.shape[0], nv, 32), dtype=np.float64)
        for s in range(dp // 32):
            acc = acc + ai[:, None, s, :] * b3[None, :, s, :]
        for h in (16, 8, 4, 2, 1):
            acc = acc[..., :h] + acc[..., h:2 * h]
        out[i0:i0 + chunk] = acc[..., 0]
    return out


def topk_ties(score, k, axis=1):
    """evaluation/evaluation_mm.py:257,259 `topk(k)`.  torch leaves tie order unspecified;
    ours is (score descending, index ascending).  Returns (values, indices)."""
    s = np.asarray(score)
    if axis == 0:
        v, i = topk_ties(s.T, k, axis=1)
        return v.T, i.T
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, axis=1), order


def rank_of_gt(score, gt_col):
    """Rank the gt column would get in a stable descending sort of each row
    (evaluation_mm.py:333-338: `indice_matrix[i].index(gt_indice)`), lower index first on ties."""
    s = np.asarray(score)
    g = s[np.arange(s.shape[0]), gt_col][:, None]
    cols = np.arange(s.shape[1])[None, :]
    return ((s > g) | ((s == g) & (cols < np.asarray(gt_col)[:, None]))).sum(axis=1)


def compute_metric_ret(score_matrix_, ids, ids_txt, direction="forward"):
    """evaluation/evaluation_mm.py:326-380, same return dict (values rounded to 0.1).
    forward: rank of ids.index(ids_txt[i]) in row i; backward: min rank over all texts whose
    id equals the video id (multi-caption), ranking down columns."""
    s = np.asarray(score_matrix_)
    assert s.shape == (len(ids_txt), len(ids))
    first = {}
    for j, v in enumerate(ids):
        first.setdefault(v, j)  # list.index returns the first occurrence
    if direction == "forward":
        gt = np.array([first[t] for t in ids_txt])
        rank = rank_of_gt(s, gt)
        n = len(ids_txt)
        tag = "forward"
    else:
        st = s.T
        rank = np.empty(len(ids), dtype=np.int64)
        ids_txt_arr = list(ids_txt)
        for i, v in enumerate(ids):
            gts = [t for t, tv in enumerate(ids_txt_arr) if tv == v]
            rank[i] = min(rank_of_gt(st[i:i + 1], np.array([g]))[0] for g in gts)
        n = len(ids)
        tag = "backward"
    r1 = int((rank < 1).sum()) / n
    r5 = int((rank < 5).sum()) / n
    r10 = int((rank < 10).sum()) / n
    return {
        f"{tag}_r1": round(r1 * 100, 1),
        f"{tag}_recall": f"{round(r1 * 100, 1)}/{round(r5 * 100, 1)}/{round(r10 * 100, 1)}",
        f"{tag}_ravg": round((r1 + r5 + r10) / 3 * 100, 1),
    }


def refine_score_matrix(condition_feats_per_rank, input_ids, attention_mask, score_matrix_t_cond,
                        slice_scorer, itm_rerank_num, direction="forward", small_batch=25):
    """evaluation/evaluation_mm.py:253-319 restated for ALL ranks at once.

    condition_feats_per_rank: list (one entry per rank) of [Nv_r, S, H] arrays (each rank
    holds only its own videos, :274-286).  For every video column with at least one
    candidate, the candidate texts (top-k by row for 'forward', by column for 'backward')
    are scored by `slice_scorer(cond[b,S,H], ids[b,L], mask[b,L]) -> [b]` in chunks of 25
    (:302) and scattered; everything else stays 0.  Returns the [Nt, Nv] matrix the
    final ddp_allgather(...).T assembles (:317)."""
    s = np.asarray(score_matrix_t_cond)
    nt, nv = s.shape
    k = itm_rerank_num
    mask = np.zeros((nt, nv), dtype=bool)
    if direction == "forward":
        _, idx = topk_ties(s, k, axis=1)
        mask[np.arange(nt)[:, None], idx] = True
    else:
        _, idx = topk_ties(s, k, axis=0)
        mask[idx, np.arange(nv)[None, :]] = True
    out = np.zeros((nt, nv), dtype=np.float64)
    col = 0
    for cond in condition_feats_per_rank:
        for i in range(len(cond)):
            rows = np.nonzero(mask[:, col])[0]
            if len(rows):
                vals = []
                for c0 in range(0, len(rows), small_batch):
                    r = rows[c0:c0 + small_batch]
                    c = np.broadcast_to(np.asarray(cond[i])[None], (len(r),) + np.asarray(cond[i]).shape)
                    vals.append(np.asarray(slice_scorer(c, np.asarray(input_ids)[r], np.asarray(attention_mask)[r])))
                out[rows, col] = np.concatenate(vals)
            col += 1
    assert col == nv
    return out


def evaluate_ret(feat_t, input_ids, attention_mask, feat_cond, condition_feats_per_rank, ids, ids_txt,
                 slice_scorer, itm_rerank_num, bidirection=False):
    """evaluation/evaluation_mm.py:171-251 after the per-batch collection: per sub-task the ITC metrics on
    feat_t @ feat_cond.T (:223-232, keys renamed forward->video, backward->txt :226,229) and the ITM metrics on the
    re-ranked matrix (:236-249).  feat_cond / condition_feats_per_rank: dicts keyed by sub-task."""
    val_log = {}
    scores = {}
    for task, fc in feat_cond.items():
        scores[task] = score_matrix(feat_t, fc)
        log = {k.replace("forward", "video"): v for k, v in compute_metric_ret(scores[task], ids, ids_txt, "forward").items()}
        if bidirection:
            log.update({k.replace("backward", "txt"): v
                        for k, v in compute_metric_ret(scores[task], ids, ids_txt, "backward").items()})
        val_log[f"ret_itc_{task}"] = log
    for task in feat_cond:
        r = refine_score_matrix(condition_feats_per_rank[task], input_ids, attention_mask, scores[task], slice_scorer,
                                itm_rerank_num, "forward")
        log = {k.replace("forward", "video"): v for k, v in compute_metric_ret(r, ids, ids_txt, "forward").items()}
        if bidirection:
            r = refine_score_matrix(condition_feats_per_rank[task], input_ids, attention_mask, scores[task],
                                    slice_scorer, itm_rerank_num, "backward")
            log.update({k.replace("backward", "txt"): v for k, v in compute_metric_ret(r, ids, ids_txt, "backward").items()})
        val_log[f"ret_itm_{task}"] = log
    return val_log


# ----------------------------------------------------------------------------
# collectives (semantics only)
# ----------------------------------------------------------------------------


def concat_all_gather(per_rank):
    """utils/distributed.py:50-66: concatenate equal-shaped per-rank tensors in rank order."""
    return np.concatenate([np.asarray(x) for x in per_rank], axis=0)


def ddp_allgather(per_rank):
    """utils/distributed.py:133-149: ragged all-gather (pad to max, gather, trim) ==
    concatenation of the ragged per-rank tensors in rank order."""
    return np.concatenate([np.asarray(x) for x in per_rank], axis=0)
