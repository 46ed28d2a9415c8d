"""OMC/ITC contrastive loss + hard-negative sampling + negative gather: the body of
`VAST.forward_ret` (model/vast.py:383-464) on the fused CUDA path.

`forward_ret` below keeps the reference signature and return values, so it can be bound to a VAST
module in place of the reference method."""
from __future__ import annotations

import itertools

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .distributed import _rank, _world, all_gather_with_grad, concat_all_gather, exchange_rows

# fetch only the sampled negative rows of `condition_feats` from their owner ranks instead of
# all-gathering the whole [N, S, 768] tensor with gradient (model/vast.py:422).  Same values, same gradients.
NEGATIVE_ROW_EXCHANGE = True

_call_counter = itertools.count()


class _OmcFn(torch.autograd.Function):
    """loss, negatives and unit gradients in one fused forward; backward scales the unit gradients
    (gradients reach only the local feature rows and the temperature, exactly like the reference:
    the gathered operands come from a no_grad all-gather, utils/distributed.py:50)."""

    @staticmethod
    def forward(ctx, feat_cond, feat_t, contra_temp, rank, world_size, label_smoothing, weight_floor, seed, offset,
                need_sample, debug_noise):
        bs, dim = feat_t.shape
        pg = None
        if world_size > 1 and dim % 8 == 0:
            from .peer import packed_gather
            pg = packed_gather(bs, dim, feat_t.device)
        need_grad = any(ctx.needs_input_grad[:3])
        temp = contra_temp.detach() if isinstance(contra_temp, torch.Tensor) else contra_temp
        out = None
        if world_size == 1 and ops.local_step_ok(feat_t) and ops.local_step_ok(feat_cond):
            # one rank: nothing to gather, so packing is fused into the step's first kernel
            out = ops.omc_step_local(feat_t.detach(), feat_cond.detach(), temp, label_smoothing, weight_floor, seed, offset,
                                     need_sample, need_grad, debug_noise)
        elif pg is not None:
            # pack + all-gather in ONE kernel: rows go straight into every rank's gathered buffer over NVLink
            pack = pg.gather(feat_t.detach(), feat_cond.detach())
        else:
            local = ops.pack_pair(feat_t.detach(), feat_cond.detach())
            if world_size > 1:
                pack = torch.empty(world_size * bs, 2 * dim, dtype=torch.bfloat16, device=local.device)
                dist.all_gather_into_tensor(pack, local)  # ONE collective for both feature blocks
            else:
                pack = local
        if out is None:
            out = ops.omc_step(pack, bs, rank * bs, temp, label_smoothing, weight_floor, seed, offset, need_sample, need_grad,
                               debug_noise)
        if need_grad:
            ctx.save_for_backward(out["grad_cond"], out["grad_t"], out["grad_temp"])
        ctx.temp_shape = contra_temp.shape if isinstance(contra_temp, torch.Tensor) else None
        ctx.dtypes = (feat_cond.dtype, feat_t.dtype)
        loss = out["loss"].reshape(())
        if need_sample:
            neg_text, neg_cond = out["neg_idx"][0], out["neg_idx"][1]
            ctx.mark_non_differentiable(neg_text, neg_cond)
            return loss, neg_text, neg_cond
        return loss, None, None

    @staticmethod
    def backward(ctx, g, _a, _b):
        need = ctx.needs_input_grad
        want_temp = ctx.temp_shape is not None and need[2]
        src = [t for t, w in zip(ctx.saved_tensors, (need[0], need[1], want_temp)) if w]
        scaled = iter(torch._foreach_mul(src, g) if src else ())       # one launch for all three products
        d_cond = next(scaled).to(ctx.dtypes[0]) if need[0] else None
        d_t = next(scaled).to(ctx.dtypes[1]) if need[1] else None
        d_temp = next(scaled).reshape(ctx.temp_shape) if want_temp else None
        return d_cond, d_t, d_temp, None, None, None, None, None, None, None, None


def omc_loss_and_negatives(feat_cond, feat_t, contra_temp, *, rank=None, world_size=None, label_smoothing=0.1,
                           weight_floor=1e-4, generator=None, need_negatives=True, debug_noise=None):
    """Fused replacement of model/vast.py:404-440.

    Returns (loss, neg_idx_cond2t, neg_idx_t2cond):
      loss            scalar, differentiable in feat_cond, feat_t, contra_temp
      neg_idx_cond2t  [bs] int64: negative TEXT row (in the all-gathered order) per local condition row
      neg_idx_t2cond  [bs] int64: negative CONDITION row per local text row
    Indices stay on the device (the reference does 2*bs `.item()` host syncs here)."""
    rank = _rank() if rank is None else rank
    world_size = _world() if world_size is None else world_size
    seed = generator.initial_seed() if generator is not None else torch.initial_seed()
    offset = next(_call_counter)
    return _OmcFn.apply(feat_cond, feat_t, contra_temp, rank, world_size, float(label_smoothing), float(weight_floor),
                        seed, offset, need_negatives, debug_noise)


class _GatherConcat3(torch.autograd.Function):
    """condition_feats = cat(cond, cond_all[neg], cond) (vast.py:432-433,448) with its gradient."""

    @staticmethod
    def forward(ctx, cond_local, cond_all, ids_local, mask_local, ids_all, mask_all, neg_text, neg_cond):
        ids1, att1, cond3 = ops.gather_rows_concat3(ids_local, mask_local, ids_all, mask_all, cond_local, cond_all,
                                                    neg_text, neg_cond)
        ctx.save_for_backward(neg_cond)
        ctx.n_all = cond_all.shape[0]
        ctx.mark_non_differentiable(ids1, att1)
        return ids1, att1, cond3

    @staticmethod
    def backward(ctx, _g1, _g2, g):
        (neg_cond,) = ctx.saved_tensors
        bs = neg_cond.shape[0]
        g_local = g_all = None
        if ctx.needs_input_grad[0]:
            g_local = g[:bs] + g[2 * bs:]
        if ctx.needs_input_grad[1]:
            g_all = torch.zeros((ctx.n_all,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
            g_all.index_add_(0, neg_cond, g[bs:2 * bs])
        return g_local, g_all, None, None, None, None, None, None


class _PeerGatherConcat3(torch.autograd.Function):
    """cat(cond, cond_all[neg_cond], cond) with the negative rows read from their owners over peer memory and the
    gradients pulled back by the owners: no [N, S, H] all-gather, no index exchange in the forward, no
    host synchronisation in either direction."""

    @staticmethod
    def forward(ctx, cond_local, ids_local, mask_local, ids_all, mask_all, neg_text, neg_cond, pr):
        ptrs = pr.publish(cond_local.detach(), "fwd")
        ids1, att1, cond3 = ops.gather_rows_concat3_peer(ids_local, mask_local, ids_all, mask_all, cond_local.detach(), ptrs,
                                                         pr.bs, neg_text, neg_cond)
        ctx.save_for_backward(neg_cond)
        ctx.pr = pr
        ctx.mark_non_differentiable(ids1, att1)
        return ids1, att1, cond3

    @staticmethod
    def backward(ctx, _g1, _g2, g):
        (neg_cond,) = ctx.saved_tensors
        pr = ctx.pr
        bs = neg_cond.shape[0]
        g = g.contiguous()
        base = g[:bs] + g[2 * bs:]                                   # the two copies of the local rows
        ptrs = pr.publish(g[bs:2 * bs], "bwd")
        req = torch.empty(pr.world * bs, dtype=torch.int64, device=g.device)
        dist.all_gather_into_tensor(req, neg_cond.contiguous())
        gx = ops.pull_row_grads(req, ptrs, bs, pr.rank * bs, g[:bs], base_grad=base)
        return gx, None, None, None, None, None, None, None


def gather_negatives_peer(cond_local, ids_local, mask_local, ids_all, mask_all, neg_text, neg_cond):
    """`gather_negatives` for world_size > 1 without ever gathering condition_feats: returns None when peer memory is
    unavailable (the caller falls back to `exchange_rows` + `gather_negatives`)."""
    from .peer import peer_rows
    pr = peer_rows(cond_local.shape[0], cond_local.shape[1:], cond_local.dtype, cond_local.device)
    if pr is None or (cond_local[0].numel() * cond_local.element_size()) % 16 != 0:
        return None
    return _PeerGatherConcat3.apply(cond_local, ids_local, mask_local, ids_all, mask_all, neg_text, neg_cond, pr)


def gather_negatives(cond_local, cond_all, ids_local, mask_local, ids_all, mask_all, neg_text, neg_cond):
    """(input_ids_1 [3bs,L], attention_mask_1 [3bs,L], condition_feats [3bs,S,H])  -- vast.py:429-448."""
    return _GatherConcat3.apply(cond_local, cond_all, ids_local, mask_local, ids_all, mask_all, neg_text, neg_cond)


def forward_ret(self, batch, task, compute_loss=True):
    """Drop-in for `VAST.forward_ret` (model/vast.py:383-483): same arguments, same returned dicts.
    `self` needs what the reference method uses: batch_get, contra_temp, itm_ratio,
    multimodal_encoder.bert, itm_head."""
    if isinstance(batch.raw_captions[0], list):
        batch.raw_captions = [i for j in batch.raw_captions for i in j]
    subtasks = task.split('%')[1:]
    if not compute_loss:
        evaluation_dict = {}
        evaluation_dict['feat_t'] = self.batch_get(batch, 'feat_t')
        caption_tokens = self.batch_get(batch, 'caption_tokens')
        evaluation_dict['input_ids'] = caption_tokens.input_ids
        evaluation_dict['attention_mask'] = caption_tokens.attention_mask
        for t in subtasks:
            assert t in ['tv', 'ta', 'tva', 'tvs', 'tvas']
            evaluation_dict[f'feat_cond_{t}'] = self.batch_get(batch, f'feat_{t[1:]}')
            evaluation_dict[f'condition_feats_{t}'] = self.batch_get(batch, f'condition_feats_{t[1:]}')
        return evaluation_dict

    loss_itc, loss_itm = [], []
    feat_t = self.batch_get(batch, 'feat_t')
    caption_tokens = self.batch_get(batch, 'caption_tokens')
    input_ids, attention_mask = caption_tokens.input_ids, caption_tokens.attention_mask
    input_ids_collate = concat_all_gather(input_ids)
    attention_mask_collate = concat_all_gather(attention_mask)
    for t in subtasks:
        assert t in ['tv', 'ta', 'tva', 'tvs', 'tvas']
        feat_cond = self.batch_get(batch, f'feat_{t[1:]}')
        loss, neg_text, neg_cond = omc_loss_and_negatives(feat_cond, feat_t, self.contra_temp)
        loss_itc.append(loss)
        condition_feats = self.batch_get(batch, f'condition_feats_{t[1:]}')
        peer_out = None
        if NEGATIVE_ROW_EXCHANGE and _world() > 1:
            # the sampled rows come straight from their owners' symmetric-memory blocks into the [3bs, S, H] ITM input
            peer_out = gather_negatives_peer(condition_feats, input_ids, attention_mask, input_ids_collate,
                                             attention_mask_collate, neg_text, neg_cond)
        if peer_out is not None:
            input_ids_1, attention_mask_1, condition_feats_3 = peer_out
        elif NEGATIVE_ROW_EXCHANGE and _world() > 1:
            cond_neg = exchange_rows(condition_feats, neg_cond)          # [bs, S, H]: only the sampled rows travel
            own = torch.arange(cond_neg.shape[0], dtype=torch.int64, device=cond_neg.device)
            input_ids_1, attention_mask_1, condition_feats_3 = gather_negatives(
                condition_feats, cond_neg, input_ids, attention_mask, input_ids_collate, attention_mask_collate,
                neg_text, own)
        else:
            condition_feats_collate = all_gather_with_grad(condition_feats)
            input_ids_1, attention_mask_1, condition_feats_3 = gather_negatives(
                condition_feats, condition_feats_collate, input_ids, attention_mask, input_ids_collate,
                attention_mask_collate, neg_text, neg_cond)
        output = self.multimodal_encoder.bert(input_ids=input_ids_1, attention_mask=attention_mask_1,
                                              encoder_hidden_states=condition_feats_3).last_hidden_state
        batch_size = neg_cond.shape[0]
        logits = self.itm_head(output[:, 0].half())
        ground_truth = torch.zeros(batch_size * 3, dtype=torch.long, device=logits.device)
        ground_truth[:batch_size] = 1
        loss_itm.append(self.itm_ratio * F.cross_entropy(logits, ground_truth))
    return {'loss_itc': sum(loss_itc) / len(loss_itc), 'loss_itm': sum(loss_itm) / len(loss_itm)}
