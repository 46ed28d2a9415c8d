"""Retrieval evaluation step (evaluation/evaluation_mm.py:171-380) on the CUDA path.

Drop-in functions keep the reference's names, arguments and return values:
    compute_metric_ret(score_matrix, ids, ids_txt, direction)            evaluation_mm.py:326
    refine_score_matrix(condition_feats, input_ids, attention_mask,
                        score_matrix_t_cond, model, itm_rerank_num, direction)   evaluation_mm.py:253
    evaluate_ret(model, tasks, val_loader, global_step)                  evaluation_mm.py:171
New streaming entry points never build the [Nt, Nv] matrix:
    retrieval_topk, recall_from_feats, refine_candidates."""
from __future__ import annotations

import torch

from . import ops
from .distributed import _rank, _world, all_gather_ids, all_gather_list, ddp_allgather


# ------------------------------------------------------------------ streaming similarity + top-k
def _shard_bounds(n: int, rank: int, world: int):
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


@torch.no_grad()
def retrieval_topk(feat_t: torch.Tensor, feat_cond: torch.Tensor, k: int, mode: str = "bf16", shard=None,
                   shard_mode: str = "rows", warm_bounds: bool = True, _delta_scale: float = 1.0):
    """For every row of feat_t [Nt, D] the k best rows of feat_cond [Nv, D] by (score desc, index asc).

    mode "bf16": bf16 inputs, fp32 accumulate (tensor cores) -- scores are the bf16-mode similarities.
    mode "fp32": exact ranking of the fp32 similarities: tensor-core shortlist from 3-term bf16 splits,
                 fp64 re-score in a fixed summation order, proof check (shortlist cut-off + error bound below
                 the k-th exact score) and brute-force fp64 fallback for rows that cannot be proven.
    shard (rank, world): split the work over the ranks of the evaluation; both feature matrices are the FULL ones
                 (the reference all-gathers them, evaluation_mm.py:212,222).  Returns (values f32, indices i32) [Nt, k],
                 identical on every rank and identical to the unsharded result.
    shard_mode "rows" (default): every rank scores its slice of the QUERY rows against all columns and the finished
                 lists are all-gathered -- no merge, and every row's list is built once instead of once per rank;
                 "cols": every rank scores all rows against its slice of the columns (the layout of the ITM stage,
                 SURVEY 8e), candidate lists are all-gathered and merged.  Top-k list work does not shrink with the
                 column count, so "cols" scales worse (measured at W=8, cfg5: 2.8 ms vs 1.7 ms)."""
    assert feat_t.dim() == 2 and feat_cond.dim() == 2 and feat_t.shape[1] == feat_cond.shape[1]
    nt, nv = feat_t.shape[0], feat_cond.shape[0]
    rank, world = shard if shard is not None else (0, 1)
    lo, hi = _shard_bounds(nv, rank, world)
    exact = mode in ("fp32", "fp32x3")
    if not exact and mode != "bf16":
        raise ValueError(mode)
    # exact mode: the shortlist comes from 2-term bf16 splits (3 tensor-core products; "fp32x3": 3 terms, 6 products,
    # twice the tensor work) -- either way it is re-scored in fp64 and PROVEN complete with the same error bound
    sim_mode = (ops.SIM_FP32X3 if mode == "fp32x3" else ops.SIM_FP32X2) if exact else ops.SIM_BF16
    kl = k if not exact else min(ops.TOPK_MAX, max(2 * k, k + 16))
    if kl > ops.TOPK_MAX or k > ops.TOPK_MAX:
        raise RuntimeError(f"retrieval_topk: k={k} exceeds the supported maximum {ops.TOPK_MAX}")
    ft = feat_t.float().contiguous() if exact else feat_t.contiguous()
    fc = feat_cond.float().contiguous() if exact else feat_cond.contiguous()
    if shard_mode == "rows" and world > 1:
        import torch.distributed as dist
        per = (nt + world - 1) // world
        rlo, rhi = min(rank * per, nt), min((rank + 1) * per, nt)
        mine = torch.zeros(per, kl, dtype=torch.int64, device=ft.device)
        if rhi > rlo:
            q_op = ops.sim_pack_operand(ft[rlo:rhi], sim_mode, True)
            k_op = ops.sim_pack_operand(fc, sim_mode, False)
            mine[:rhi - rlo] = ops.sim_topk(q_op, k_op, kl)
        allk = torch.empty(world * per, kl, dtype=torch.int64, device=ft.device)
        dist.all_gather_into_tensor(allk, mine)
        keys = allk[:nt].contiguous()
    else:
        if shard_mode not in ("cols", "rows"):
            raise ValueError(shard_mode)
        q_op = ops.sim_pack_operand(ft, sim_mode, True)
        k_op = ops.sim_pack_operand(fc[lo:hi], sim_mode, False) if hi > lo else None
        bounds = None
        if world >= 4 and warm_bounds:    # measured (cfg5, ranks emulated): pays from 4 shards on; at 2 phase A costs what it saves
            # Column shards start WARM (each rank owns its videos).  A cold shard does k (1 + ln(n / (W k)))
            # list insertions per row -- almost as many as the whole problem -- so the list work would not shrink with W.
            # Phase A: this rank scans its slice of the QUERY rows against its own columns (1 / W^2 of the problem); the
            # k-th score found is a proven lower bound of the row's global k-th score.  The [Nt] bounds are
            # all-gathered (4 bytes per row) and phase B admits only scores at or above them: ~k insertions per row.
            import torch.distributed as dist
            per = (nt + world - 1) // world
            rlo, rhi = min(rank * per, nt), min((rank + 1) * per, nt)
            mine_b = torch.zeros(per, dtype=torch.int32, device=ft.device)
            if rhi > rlo and k_op is not None and hi - lo >= kl:
                _, b = ops.sim_topk(q_op[rlo:rhi], k_op, kl, col_offset=lo, want_bounds=True)
                mine_b[:rhi - rlo] = b
            allb = torch.empty(world * per, dtype=torch.int32, device=ft.device)
            dist.all_gather_into_tensor(allb, mine_b)
            bounds = allb[:nt].contiguous()
        if k_op is not None:
            keys = ops.sim_topk(q_op, k_op, kl, col_offset=lo, bounds_in=bounds)
        else:
            keys = torch.zeros(nt, kl, dtype=torch.int64, device=ft.device)
        if world > 1:
            # merge fused into the exchange: rank r merges the candidates of row slice r (all-to-all: every rank sends
            # each peer only that peer's rows, (W-1)/W * Nt*k keys in and out) and the finished lists are all-gathered --
            # W/2 x less traffic than all-gathering every rank's [Nt, k] candidates, and the merge work is sharded too
            import torch.distributed as dist
            per = (nt + world - 1) // world
            send = torch.zeros(world * per, kl, dtype=torch.int64, device=keys.device)
            send[:nt] = keys
            recv = torch.empty(world, per, kl, dtype=torch.int64, device=keys.device)
            dist.all_to_all_single(recv, send)
            mine = ops.topk_merge(recv, kl)
            allk = torch.empty(world * per, kl, dtype=torch.int64, device=keys.device)
            dist.all_gather_into_tensor(allk, mine)
            keys = allk[:nt].contiguous()
    vals, idx = ops.topk_unpack(keys)
    if not exact:
        return vals, idx
    idx, s64 = ops.rescore_f64(ft, fc, idx)
    kk = min(k, nv)
    # proof: every column outside the shortlist has approx <= cutoff, hence exact <= cutoff + delta
    cutoff = vals[:, kl - 1].double()                         # -inf when the list is not full (all columns inside)
    delta = _delta_scale * (2.0 ** -11) * ft.norm(dim=1).double() * fc.norm(dim=1).max().double()
    unproven = torch.isfinite(cutoff) & ~(s64[:, kk - 1] > cutoff + delta)
    idx_k = idx[:, :k].contiguous()
    s_k = s64[:, :k].contiguous()
    rows = unproven.nonzero().flatten().int()
    if rows.numel() > 0:
        ops.exact_topk_rows(ft, fc, rows, k, idx_k, s_k)
    return s_k.float(), idx_k


@torch.no_grad()
def rank_of_gt(feat_t: torch.Tensor, feat_cond: torch.Tensor, gt_col: torch.Tensor, shard=None) -> torch.Tensor:
    """Streaming rank of the ground truth (evaluation_mm.py:333-338 without the sort / `.tolist()` /
    `list.index`): ranks[i] = #{j : s_ij > s_i,gt or (s_ij == s_i,gt and j < gt_col[i])} over the EXACT similarities of
    the fp32 features -- counted inside the similarity GEMM's epilogue, near-ties of the ground truth re-scored in fp64
    (`vast_rank_of_gt`), so the [Nt, Nv] matrix never exists and the result does not depend on tiling or GPU count.
    shard (rank, world): every rank counts over its slice of the columns; partial counts are summed (all-reduce).
    Returns int32 [Nt], identical on every rank."""
    ft, fc = feat_t.float().contiguous(), feat_cond.float().contiguous()
    rank, world = shard if shard is not None else (0, 1)
    lo, hi = _shard_bounds(fc.shape[0], rank, world)
    r = ops.rank_of_gt(ft, fc, gt_col, lo, hi - lo)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(r)
    return r


# ------------------------------------------------------------------ metrics
def _first_index(ids):
    first = {}
    for j, v in enumerate(ids):
        first.setdefault(v, j)
    return first


def _format(tag, r1, r5, r10):
    return {f'{tag}_r1': round(r1 * 100, 1),
            f'{tag}_recall': f'{round(r1 * 100, 1)}/{round(r5 * 100, 1)}/{round(r10 * 100, 1)}',
            f'{tag}_ravg': round((r1 + r5 + r10) / 3 * 100, 1)}


def _backward_pairs(ids, ids_txt):
    by_id = {}
    for t, v in enumerate(ids_txt):
        by_id.setdefault(v, []).append(t)
    rows, cols = [], []
    for i, v in enumerate(ids):
        ts = by_id.get(v, [])
        if not ts:
            raise ValueError(f"compute_metric_ret: video id {v!r} has no caption")
        rows += ts
        cols += [i] * len(ts)
    return rows, cols


@torch.no_grad()
def compute_metric_ret(score_matrix, ids, ids_txt, direction='forward'):
    """evaluation_mm.py:326-380 on a materialised [Nt, Nv] matrix; the full sort + `.tolist()` + `list.index`
    is replaced by a rank-of-gt count kernel (lower index first on ties)."""
    assert tuple(score_matrix.shape) == (len(ids_txt), len(ids))
    score = score_matrix.float().contiguous()
    dev = score.device
    if direction == 'forward':
        first = _first_index(ids)
        gt_col = torch.tensor([first[t] for t in ids_txt], dtype=torch.int32, device=dev)
        gt_row = torch.arange(len(ids_txt), dtype=torch.int32, device=dev)
        rank = ops.dense_rank_of_gt(score, gt_row, gt_col, axis=1)
        n = len(ids_txt)
    else:
        rows, cols = _backward_pairs(ids, ids_txt)
        r = ops.dense_rank_of_gt(score, torch.tensor(rows, dtype=torch.int32, device=dev),
                                 torch.tensor(cols, dtype=torch.int32, device=dev), axis=0)
        rank = torch.full((len(ids),), 2 ** 30, dtype=torch.int32, device=dev)
        rank = rank.scatter_reduce(0, torch.tensor(cols, dtype=torch.int64, device=dev), r, reduce='amin')
        n = len(ids)
    counts = torch.stack([(rank < 1).sum(), (rank < 5).sum(), (rank < 10).sum()]).tolist()  # one host read
    return _format(direction, counts[0] / n, counts[1] / n, counts[2] / n)


@torch.no_grad()
def recall_from_feats(feat_t, feat_cond, ids, ids_txt, direction='forward', mode='bf16', shard=None, method='topk'):
    """R@1/5/10 without the score matrix.  method 'topk': R@K only needs the top-10 lists (the reference reports nothing
    else: median / mean rank are computed and dropped, evaluation_mm.py:343-351); method 'rank': the reference's own
    formulation -- the rank of every ground truth (`rank_of_gt`, exact fp32 semantics), then `rank < K`."""
    dev = feat_t.device
    if method == 'rank':
        if direction == 'forward':
            first = _first_index(ids)
            gt = torch.tensor([first[t] for t in ids_txt], dtype=torch.int32, device=dev)
            rank = rank_of_gt(feat_t, feat_cond, gt, shard)
            n = len(ids_txt)
        else:
            rows, cols = _backward_pairs(ids, ids_txt)                  # (text, video) ground-truth pairs
            cols_t = torch.tensor(cols, dtype=torch.int64, device=dev)
            r = rank_of_gt(feat_cond.float()[cols_t], feat_t, torch.tensor(rows, dtype=torch.int32, device=dev), shard)
            rank = torch.full((len(ids),), 2 ** 30, dtype=torch.int32, device=dev)
            rank = rank.scatter_reduce(0, cols_t, r, reduce='amin')
            n = len(ids)
        counts = torch.stack([(rank < 1).sum(), (rank < 5).sum(), (rank < 10).sum()]).tolist()
        return _format(direction, counts[0] / n, counts[1] / n, counts[2] / n)
    if direction == 'forward':
        _, idx = retrieval_topk(feat_t, feat_cond, min(10, feat_cond.shape[0]), mode, shard)
        first = _first_index(ids)
        gt = torch.tensor([first[t] for t in ids_txt], dtype=torch.int32, device=dev)
        hit = idx == gt[:, None]
        n = len(ids_txt)
    else:
        _, idx = retrieval_topk(feat_cond, feat_t, min(10, feat_t.shape[0]), mode, shard)   # top texts per video
        txt_id = {v: i for i, v in enumerate(dict.fromkeys(ids_txt))}
        vid_code = torch.tensor([txt_id.get(v, -2) for v in ids], dtype=torch.int64, device=dev)
        txt_code = torch.tensor([txt_id[v] for v in ids_txt], dtype=torch.int64, device=dev)
        hit = txt_code[idx.clamp_min(0).long()] == vid_code[:, None]
        hit &= idx >= 0
        n = len(ids)
    pos = torch.where(hit.any(dim=1), hit.float().argmax(dim=1), torch.full((hit.shape[0],), 10 ** 6, device=dev))
    counts = torch.stack([(pos < 1).sum(), (pos < 5).sum(), (pos < 10).sum()]).tolist()
    return _format(direction, counts[0] / n, counts[1] / n, counts[2] / n)


# ------------------------------------------------------------------ ITM re-rank bookkeeping
@torch.no_grad()
def _rerank_pairs(condition_feats, input_ids, attention_mask, text_idx, video_local, model, small_batch, pair_batch=8192):
    """Score (text, local video) candidate pairs with model.compute_slice_scores (model/vast.py:373-380),
    per video in chunks of `small_batch` like evaluation_mm.py:292-311 -- or, when the model offers
    `compute_pair_scores(condition_feats, video_idx [b], input_ids [b, L], attention_mask [b, L]) -> [b]`, in batches of
    `pair_batch` pairs across videos.  Returns (texts, videos, scores)."""
    nv_local = condition_feats.shape[0]
    if nv_local == 0 or text_idx.numel() == 0:   # a rank that owns no videos (ragged evaluation shards): nothing to score
        dev = condition_feats.device
        return (torch.empty(0, dtype=torch.int32, device=dev), torch.empty(0, dtype=torch.int32, device=dev),
                torch.empty(0, dtype=torch.float32, device=dev))
    offsets, texts = ops.bucket_by_video(text_idx, video_local, nv_local)
    if hasattr(model, "compute_pair_scores"):
        # a scorer that takes (video index, text) PAIRS batches across videos -- `pair_batch` pairs per call
        # instead of <= 25 texts of one video -- and can project a video's cross-attention K/V once for all its pairs.
        # The pairs arrive grouped by video (ascending text inside a video), i.e. in the order the reference scores them.
        n_pairs = offsets[-1:].long()                                       # device scalar: no host read on this path
        cap = text_idx.numel()
        vids_all = torch.searchsorted(offsets[1:].long().contiguous(), torch.arange(cap, device=offsets.device), right=True).int()
        valid = torch.arange(cap, device=offsets.device) < n_pairs
        vids_all = torch.where(valid, vids_all.clamp_max(nv_local - 1), torch.zeros_like(vids_all))
        rows = torch.where(valid, texts.long().clamp_min(0), torch.zeros_like(texts.long()))
        out_scores = torch.zeros(cap, dtype=torch.float32, device=condition_feats.device)
        for c in range(0, cap, pair_batch):
            sl = slice(c, min(c + pair_batch, cap))
            out_scores[sl] = model.compute_pair_scores(condition_feats, vids_all[sl].long(), input_ids[rows[sl]],
                                                       attention_mask[rows[sl]]).float()
        keep = valid.nonzero().flatten()                                     # (sizes the outputs: one host read, as below)
        return texts[keep], vids_all[keep], out_scores[keep]
    off = offsets.cpu().tolist()                       # one host read for the whole re-rank
    out_scores = torch.empty(off[-1], dtype=torch.float32, device=condition_feats.device)
    vids = torch.empty(off[-1], dtype=torch.int32, device=condition_feats.device)
    for i in range(nv_local):
        b, e = off[i], off[i + 1]
        if e == b:
            continue
        rows = texts[b:e].long()
        cur_ids, cur_mask = input_ids[rows], attention_mask[rows]
        cond = condition_feats[i].unsqueeze(0).expand(e - b, -1, -1)
        for c in range(0, e - b, small_batch):
            out_scores[b + c:b + min(c + small_batch, e - b)] = model.compute_slice_scores(
                cond[c:c + small_batch], cur_ids[c:c + small_batch], cur_mask[c:c + small_batch]).float()
        vids[b:e] = i
    return texts[:off[-1]], vids, out_scores


@torch.no_grad()
def refine_score_matrix(condition_feats, input_ids, attention_mask, score_matrix_t_cond, model, itm_rerank_num,
                        direction='forward', small_batch=25):
    """Drop-in for evaluation_mm.py:253-319.  The Nt*k scalar writes into a dense int64 mask, the host-side
    `sum(mask[:, i] == 1)` and the boolean-mask gathers are replaced by: top-k kernel (ties -> lower
    index) -> counting sort of the candidate pairs by video -> scatter of the ITM scores."""
    k = itm_rerank_num
    score = score_matrix_t_cond.float().contiguous()
    nt, nv = score.shape
    dev = score.device
    if direction == 'forward':
        _, idx = ops.dense_topk(score, min(k, nv), axis=1)                     # [nt, k] video per text
        t_idx = torch.arange(nt, dtype=torch.int32, device=dev)[:, None].expand_as(idx).reshape(-1)
        v_idx = idx.reshape(-1)
    else:
        _, idx = ops.dense_topk(score, min(k, nt), axis=0)                     # [k, nv] text per video
        t_idx = idx.reshape(-1)
        v_idx = torch.arange(nv, dtype=torch.int32, device=dev)[None, :].expand_as(idx).reshape(-1)
    rank = _rank()
    cur_length = condition_feats.shape[0]
    length_ls = all_gather_list(cur_length)
    start = sum(length_ls[:rank])
    local = (v_idx >= start) & (v_idx < start + cur_length) & (t_idx >= 0)
    v_local = torch.where(local, v_idx - start, torch.full_like(v_idx, -1))
    texts, vids, scores = _rerank_pairs(condition_feats, input_ids, attention_mask, t_idx, v_local, model, small_batch)
    cur_new = torch.zeros(nt, cur_length, dtype=torch.float32, device=dev)
    if texts.numel() > 0:
        ops.scatter_scores(texts, vids, scores, cur_new)
    out = ddp_allgather(cur_new.T.contiguous()).T
    return out.to(score_matrix_t_cond.dtype)


# ------------------------------------------------------------------ streaming re-rank (no [Nt, Nv] matrix anywhere)
@torch.no_grad()
def refine_candidates(condition_feats, input_ids, attention_mask, feat_t, feat_cond, model, itm_rerank_num,
                      direction='forward', mode='fp32', small_batch=25, pair_batch=8192):
    """`refine_score_matrix` (evaluation_mm.py:253-319) without the score matrix: candidates come from the streaming
    top-k over the FEATURES, the refined scores stay in the [rows, k] layout of the candidate lists.

    condition_feats [Nv_local, S, H] are this rank's videos (rank order = column order, as in the reference);
    input_ids / attention_mask [Nt, L], feat_t [Nt, D], feat_cond [Nv, D] are the gathered, full tensors.
    Returns (idx int32 [rows, k], itm f32 [rows, k]), identical on every rank:
      forward : rows = texts,  idx[t] = the k best videos of text t (score desc, index asc), itm = P(match)
      backward: rows = videos, idx[v] = the k best texts of video v
    Empty slots (fewer than k columns) have idx = -1 and itm = 0.  The dense matrix of the reference is
    `zeros(Nt, Nv)` with itm scattered at (t, idx[t, j]) / (idx[v, j], v): `recall_from_candidates` evaluates
    R@K on exactly that matrix without building it.  Every rank re-ranks the pairs whose video it owns, in the
    reference's per-video mini-batches; the [rows, k] score blocks are summed across ranks (each slot is written
    by exactly one rank) instead of the dense transposed all-gather of evaluation_mm.py:317."""
    k = itm_rerank_num
    nt, nv = feat_t.shape[0], feat_cond.shape[0]
    dev = feat_t.device
    rank, world = _rank(), _world()
    shard = (rank, world) if world > 1 else None
    if direction == 'forward':
        _, idx = retrieval_topk(feat_t, feat_cond, min(k, nv), mode, shard)            # [nt, k] video per text
        t_idx = torch.arange(nt, dtype=torch.int32, device=dev)[:, None].expand_as(idx).reshape(-1)
        v_idx = idx.reshape(-1)
    else:
        _, idx = retrieval_topk(feat_cond, feat_t, min(k, nt), mode, shard)            # [nv, k] text per video
        v_idx = torch.arange(nv, dtype=torch.int32, device=dev)[:, None].expand_as(idx).reshape(-1)
        t_idx = idx.reshape(-1)
    cur_length = condition_feats.shape[0]
    length_ls = all_gather_list(cur_length)
    start = sum(length_ls[:rank])
    local = (v_idx >= start) & (v_idx < start + cur_length) & (t_idx >= 0)
    v_local = torch.where(local, v_idx - start, torch.full_like(v_idx, -1))
    texts, vids, scores = _rerank_pairs(condition_feats, input_ids, attention_mask, t_idx, v_local, model, small_batch,
                                        pair_batch)
    # back into the list layout: both sides hold the same set of unique (text, video) pairs -> sort both by the pair key
    itm = torch.zeros(idx.numel(), dtype=torch.float32, device=dev)
    slots = local.nonzero().flatten()
    if slots.numel() > 0:
        key_slot = t_idx[slots].long() * nv + v_idx[slots].long()
        key_pair = texts.long() * nv + (vids.long() + start)
        itm[slots[torch.argsort(key_slot)]] = scores[torch.argsort(key_pair)]
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(itm)
    return idx, itm.view_as(idx)


def _rank_in_sparse_row(idx, itm, gt):
    """Rank of column gt[r] in row r of the matrix that is zero except itm[r, j] at column idx[r, j] (ties -> lower
    index first, like the dense kernel).  idx/itm [R, k], gt [R] -> int64 [R]."""
    gtc = gt[:, None].to(idx.dtype)
    valid = idx >= 0
    at_gt = valid & (idx == gtc)
    s_gt = (itm * at_gt).sum(dim=1, keepdim=True)                       # 0 when gt is not a candidate
    ahead = valid & ~at_gt & ((itm > s_gt) | ((itm == s_gt) & (idx < gtc)))
    r = ahead.sum(dim=1)
    # columns outside the list hold 0: they tie with gt only if s_gt == 0, and then the lower indices come first
    listed_below = (valid & (idx < gtc)).sum(dim=1)
    return r + torch.where(s_gt[:, 0] == 0, gt.long() - listed_below, torch.zeros_like(r))


@torch.no_grad()
def recall_from_candidates(idx, itm, ids, ids_txt, direction='forward'):
    """`compute_metric_ret` (evaluation_mm.py:326-380) on the refined matrix given as candidate lists
    (`refine_candidates`): same dict, no [Nt, Nv] matrix."""
    dev = idx.device
    if direction == 'forward':
        first = _first_index(ids)
        gt = torch.tensor([first[t] for t in ids_txt], dtype=torch.int64, device=dev)
        rank = _rank_in_sparse_row(idx, itm, gt)
        n = len(ids_txt)
    else:
        rows, cols = _backward_pairs(ids, ids_txt)                      # (text, video) ground-truth pairs
        cols_t = torch.tensor(cols, dtype=torch.int64, device=dev)
        r = _rank_in_sparse_row(idx[cols_t], itm[cols_t], torch.tensor(rows, dtype=torch.int64, device=dev))
        rank = torch.full((len(ids),), 2 ** 40, dtype=torch.int64, device=dev)
        rank = rank.scatter_reduce(0, cols_t, r, reduce='amin')
        n = len(ids)
    counts = torch.stack([(rank < 1).sum(), (rank < 5).sum(), (rank < 10).sum()]).tolist()  # one host read
    return _format(direction, counts[0] / n, counts[1] / n, counts[2] / n)


# above this many score-matrix entries evaluate_ret switches to the streaming path on its own (1 GiB of fp32)
STREAMING_THRESHOLD = 1 << 28


@torch.no_grad()
def evaluate_ret(model, tasks, val_loader, global_step, streaming=None):
    """Drop-in for evaluation_mm.py:171-251 (same val_log keys: ret_itc_{task}, ret_itm_{task}).

    streaming=False: the reference's flow on a materialised [Nt, Nv] fp32 matrix (drop-in `refine_score_matrix` /
    `compute_metric_ret`).  streaming=True: the matrix is never built -- exact-mode streaming top-k for the ITC
    metrics (`recall_from_feats`), candidate lists + ITM scores in list layout for the re-rank (`refine_candidates`,
    `recall_from_candidates`); the work is sharded over the ranks instead of repeated on each.  Default (None):
    streaming once Nt * Nv exceeds STREAMING_THRESHOLD (cfg5 would need 40 GB)."""
    val_log = {}
    ids, ids_txt, input_ids, attention_mask, feat_t = [], [], [], [], []
    subtasks = tasks.split('%')[1:]
    store = {f'feat_cond_{t}': [] for t in subtasks}
    store.update({f'condition_feats_{t}': [] for t in subtasks})
    for batch in val_loader:
        ev = model(batch, tasks, compute_loss=False)
        feat_t.append(ev['feat_t'])
        input_ids.append(ev['input_ids'])
        attention_mask.append(ev['attention_mask'])
        ids += list(batch['ids'])
        if 'ids_txt' in batch:
            if isinstance(batch['ids_txt'][0], list):
                ids_txt += [j for i in batch['ids_txt'] for j in i]
            else:
                ids_txt += list(batch['ids_txt'])
        else:
            ids_txt += list(batch['ids'])
        for t in subtasks:
            store[f'feat_cond_{t}'].append(ev[f'feat_cond_{t}'])
            store[f'condition_feats_{t}'].append(ev[f'condition_feats_{t}'])
    ids = all_gather_ids(ids)          # evaluation_mm.py:208-209 without the pickle round trip
    ids_txt = all_gather_ids(ids_txt)
    feat_t = ddp_allgather(torch.cat(feat_t, dim=0))
    input_ids = ddp_allgather(torch.cat(input_ids, dim=0))
    attention_mask = ddp_allgather(torch.cat(attention_mask, dim=0))
    bidir = bool(getattr(model.config, 'ret_bidirection_evaluation', False))
    if streaming is None:
        streaming = feat_t.shape[0] * len(ids) > STREAMING_THRESHOLD
    if streaming:
        shard = (_rank(), _world()) if _world() > 1 else None
        dirs = (('forward', 'video'),) + ((('backward', 'txt'),) if bidir else ())
        feat_cond = {}
        for t in subtasks:
            feat_cond[t] = ddp_allgather(torch.cat(store[f'feat_cond_{t}'], dim=0)).float()
            log = {}
            for d, name in dirs:
                log.update({k_.replace(d, name): v for k_, v in
                            recall_from_feats(feat_t.float(), feat_cond[t], ids, ids_txt, d, mode='fp32', shard=shard).items()})
            val_log[f'ret_itc_{t}'] = log
        for t in subtasks:
            cond = torch.cat(store[f'condition_feats_{t}'], dim=0)
            log = {}
            for d, name in dirs:
                idx, itm = refine_candidates(cond, input_ids, attention_mask, feat_t.float(), feat_cond[t], model,
                                             model.config.itm_rerank_num, direction=d, mode='fp32')
                log.update({k_.replace(d, name): v for k_, v in recall_from_candidates(idx, itm, ids, ids_txt, d).items()})
            val_log[f'ret_itm_{t}'] = log
        return val_log
    scores = {}
    for t in subtasks:
        fc = ddp_allgather(torch.cat(store[f'feat_cond_{t}'], dim=0))
        # fp32 like the reference (no autocast here, evaluation_mm.py:170,223); kept dense because the
        # drop-in refine_score_matrix signature takes the matrix -- use recall_from_feats /
        # retrieval_topk for the streaming path at scale.
        scores[t] = ops.gemm_nt_f32(feat_t.float(), fc.float())
        log = {k_.replace('forward', 'video'): v for k_, v in compute_metric_ret(scores[t], ids, ids_txt, 'forward').items()}
        if bidir:
            log.update({k_.replace('backward', 'txt'): v
                        for k_, v in compute_metric_ret(scores[t], ids, ids_txt, 'backward').items()})
        val_log[f'ret_itc_{t}'] = log
    for t in subtasks:
        cond = torch.cat(store[f'condition_feats_{t}'], dim=0)
        k = model.config.itm_rerank_num
        sm = refine_score_matrix(cond, input_ids, attention_mask, scores[t], model, k, direction='forward')
        log = {k_.replace('forward', 'video'): v for k_, v in compute_metric_ret(sm, ids, ids_txt, 'forward').items()}
        if bidir:
            sm = refine_score_matrix(cond, input_ids, attention_mask, scores[t], model, k, direction='backward')
            log.update({k_.replace('backward', 'txt'): v
                        for k_, v in compute_metric_ret(sm, ids, ids_txt, 'backward').items()})
        val_log[f'ret_itm_{t}'] = log
    return val_log
