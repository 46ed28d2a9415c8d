"""Contrastive feature build (model/vast.py:221-279 `batch_get('feat_*')`): pool -> concat -> Linear ->
F.normalize, with the pool/concat and normalise steps as fused CUDA kernels (the Linear stays a library GEMM)."""
from __future__ import annotations

import torch

from . import ops


class _PoolConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vision, audio, subtitle, vision_mode, audio_mode):
        ctx.shapes = tuple(None if x is None else tuple(x.shape) for x in (vision, audio, subtitle))
        ctx.modes = (vision_mode, audio_mode)
        ctx.dtype = next(x.dtype for x in (vision, audio, subtitle) if x is not None)
        return ops.pool_concat(None if vision is None else vision.contiguous(),
                               None if audio is None else audio.contiguous(),
                               None if subtitle is None else subtitle.contiguous(), vision_mode, audio_mode)

    @staticmethod
    def backward(ctx, grad):
        sv, sa, ss = ctx.shapes
        gv, ga, gs = ops.pool_concat_bwd(grad, sv, sa, ss, ctx.modes[0], ctx.modes[1], dtype=ctx.dtype)
        return gv, ga, gs, None, None


def pool_concat(vision=None, audio=None, subtitle=None, vision_encoder_type="evaclip", audio_encoder_type="beats"):
    """pool_vision_for_contra / pool_audio_for_contra / pool_text_for_contra (general_module.py:426-449)
    + torch.cat(dim=1) (vast.py:254,264,275)."""
    vmode = 1 if vision_encoder_type.startswith("swin") else 0
    if audio is not None and not (audio_encoder_type.startswith("ast") or audio_encoder_type.startswith("beats")):
        raise NotImplementedError(audio_encoder_type)
    amode = 0 if audio_encoder_type.startswith("ast") else 1
    return _PoolConcat.apply(vision, audio, subtitle, vmode, amode)


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        y, inv = ops.l2norm(x.contiguous(), eps, want_inv=True)
        ctx.save_for_backward(y, inv)
        ctx.eps = eps
        ctx.in_dtype = x.dtype
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, grad):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(grad, y, inv, ctx.eps).to(ctx.in_dtype), None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, dim=-1) for [rows, dim] features (vast.py:225-278)."""
    return _L2Norm.apply(x, eps)


_w_cache: dict = {}


def _weight_operand(weight: torch.Tensor, sim: int) -> torch.Tensor:
    """packed tensor-core operand of a weight matrix, rebuilt only when the parameter changed (an optimizer step bumps
    `_version`)."""
    key = (id(weight), sim)
    stamp = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.dtype)
    hit = _w_cache.get(key)
    if hit is None or hit[0] != stamp:
        hit = (stamp, ops._pack(weight.detach(), sim, False))
        _w_cache[key] = hit
    return hit[1]


class _ProjectNormalize(torch.autograd.Function):
    """F.normalize(F.linear(x, W, b), dim=-1) (vast.py:221-279) with the forward in ONE tensor-core kernel; the backward
    is vast_l2norm_bwd followed by the Linear's two library GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, mode):
        w_op = _weight_operand(weight, ops.pair_sim(x.dtype, weight.dtype, mode))
        y, inv, _ = ops.project_normalize(x.detach(), weight.detach(), None if bias is None else bias.detach(), eps, mode,
                                          w_op=w_op)
        ctx.save_for_backward(x, weight, y, inv)
        ctx.eps, ctx.has_bias = eps, bias is not None
        return y

    @staticmethod
    def backward(ctx, grad):
        x, weight, y, inv = ctx.saved_tensors
        gu = ops.l2norm_bwd(grad, y, inv, ctx.eps)                    # d loss / d (x W^T + b)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (gu.to(weight.dtype) @ weight).to(x.dtype)
        if ctx.needs_input_grad[1]:
            gw = (gu.t().to(x.dtype) @ x).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gu.sum(dim=0)
        return gx, gw, gb, None, None


def _linear_of(head):
    lin = getattr(head, "linear", head)                               # Contra_head wraps a bias-free nn.Linear
    return lin if isinstance(lin, torch.nn.Linear) else None


def project_normalize(head: torch.nn.Module, pooled: torch.Tensor, eps: float = 1e-12, mode: str | None = None) -> torch.Tensor:
    """F.normalize(head(pooled), dim=-1) for a Contra_head / nn.Linear head: the fused projection."""
    lin = _linear_of(head)
    return _ProjectNormalize.apply(pooled, lin.weight, lin.bias, eps, mode)


def build_feature(head: torch.nn.Module, vision=None, audio=None, subtitle=None, vision_encoder_type="evaclip",
                  audio_encoder_type="beats", fused: bool | None = None) -> torch.Tensor:
    """feat_{v,a,s,va,vs,vas} (vast.py:221-279): pooled = pool_concat(...); feat = normalize(head(pooled)).
    fused=True (needs a Linear / Contra_head with in_features % 8 == 0): projection, bias and normalisation run as one
    tensor-core kernel (`vast_project_normalize`); default: the head runs as a library GEMM followed by
    `vast_l2norm`.  Measured device time per call (scripts/proj_bench.py, CUDA-graph replays, K = 2944 -> D = 1024):
    the fused kernel pays a fixed ~18 us (one long K loop per tile, ticketed exchange, two passes over the accumulator)
    and loses to cuBLAS + vast_l2norm at these sizes, so it is opt-in (VAST_FUSED_PROJECTION=1 flips the default)."""
    import os
    pooled = pool_concat(vision, audio, subtitle, vision_encoder_type, audio_encoder_type)
    lin = _linear_of(head)
    if fused is None:
        fused = os.environ.get("VAST_FUSED_PROJECTION", "0") == "1" and lin is not None and lin.in_features % 8 == 0
    if fused:
        return project_normalize(head, pooled)
    return l2_normalize(head(pooled).float())


# ------------------------------------------------------------------ Match_head (general_module.py:34-42) fused
def _match_constants(head):
    """(u0, u1, sum u0, sum u1, v0, v1) of a Match_head: LayerNorm + Linear(2) folded (see include/vast_b200.h)."""
    ln, l2 = head.layernorm, head.linear2

    def build():
        with torch.no_grad():
            w2, gamma, beta = l2.weight.float(), ln.weight.float(), ln.bias.float()
            u = (w2 * gamma[None, :]).contiguous()
            v = w2 @ beta + l2.bias.float()
            host = torch.cat([u.sum(dim=1), v]).tolist()             # one host read per weight update
        return u[0].contiguous(), u[1].contiguous(), host[0], host[1], host[2], host[3]
    stamp_src = (l2.weight, ln.weight, ln.bias, l2.bias)
    key = (id(head), "match")
    stamp = tuple((t.data_ptr(), t._version) for t in stamp_src)
    hit = _w_cache.get(key)
    if hit is None or hit[0] != stamp:
        hit = (stamp, build())
        _w_cache[key] = hit
    return hit[1]


@torch.no_grad()
def match_head_scores(head: torch.nn.Module, cls: torch.Tensor, mode: str | None = None, want_logits: bool = False):
    """`F.softmax(head(cls), dim=1)[:, 1]` for a reference Match_head (general_module.py:34-42; model/vast.py:378) as ONE
    tensor-core kernel.  cls [b, hidden].  mode None: fp32-grade for fp32 / fp16 inputs, 'bf16': bf16-in / fp32-acc."""
    u0, u1, su0, su1, v0, v1 = _match_constants(head)
    w1 = head.linear1.weight
    w1_op = _weight_operand(w1, ops.pair_sim(cls.dtype, w1.dtype, mode))
    score, logits, _ = ops.match_head(cls, w1, head.linear1.bias.detach().float().contiguous(), u0, u1, su0, su1, v0, v1,
                                      float(head.layernorm.eps), mode, w1_op, want_logits)
    return (score, logits) if want_logits else score


def compute_slice_scores(self, slice_multimodal_vision_input, slice_input_ids, slice_attention_mask):
    """Drop-in for `VAST.compute_slice_scores` (model/vast.py:373-380): the cross-encoder is called exactly like the
    reference; Match_head + softmax[:, 1] run as the fused kernel."""
    out = self.multimodal_encoder.bert(input_ids=slice_input_ids, attention_mask=slice_attention_mask,
                                       encoder_hidden_states=slice_multimodal_vision_input).last_hidden_state
    return match_head_scores(self.itm_head, out[:, 0])


# ------------------------------------------------------------------ batch_get drop-in (model/vast.py:82-83, 221-312)
_FEAT_SPECS = {
    # key: (head attribute, (vision, audio, subtitle/caption) source keys)
    "feat_v": ("contra_head_v", ("vision_output", None, None)),
    "feat_a": ("contra_head_a", (None, "audio_output", None)),
    "feat_s": ("contra_head_s", (None, None, "subtitle_output")),
    "feat_t": ("contra_head_t", (None, None, "caption_output")),
    "feat_va": ("contra_head_va", ("vision_output", "audio_output", None)),
    "feat_vs": ("contra_head_vs", ("vision_output", None, "subtitle_output")),
    "feat_vas": ("contra_head_vas", ("vision_output", "audio_output", "subtitle_output")),
}
_CAPTION_FEATS = {"feat_t_omni_caption": "omni_caption_tokens", "feat_t_vision_caption": "vision_caption_tokens",
                  "feat_t_audio_caption": "audio_caption_tokens"}
FEATURE_KEYS = tuple(_FEAT_SPECS) + tuple(_CAPTION_FEATS)


def batch_get(self, batch, key):
    """Drop-in for `VAST.batch_get` (model/vast.py:82-312) on the contrastive-feature keys
    feat_t / feat_v / feat_a / feat_s / feat_va / feat_vs / feat_vas / feat_t_{omni,vision,audio}_caption: the same
    memoisation (`if key in batch: return batch[key]`, :82-83; the result is stored under batch[key]), the same inputs
    (encoder outputs fetched through `self.batch_get`), pool -> concat -> contra head -> F.normalize on the fused
    kernels (`build_feature`).  Every other key (tokenisers, encoders, condition_feats_*) is the reference's business:
    it is forwarded to the method this one replaced (`install` keeps it as `_batch_get_reference`)."""
    if key in batch:
        return batch[key]
    cfg = getattr(self, "config", None)
    vtype = getattr(cfg, "vision_encoder_type", "evaclip")
    atype = getattr(cfg, "audio_encoder_type", "beats")
    if key in _FEAT_SPECS:
        head, (kv, ka, ks) = _FEAT_SPECS[key]
        vis = self.batch_get(batch, kv) if kv else None
        aud = self.batch_get(batch, ka) if ka else None
        sub = self.batch_get(batch, ks) if ks else None
        batch[key] = build_feature(getattr(self, head), vis, aud, sub, vtype, atype)
        return batch[key]
    if key in _CAPTION_FEATS:
        tokens = self.batch_get(batch, _CAPTION_FEATS[key])
        hidden = self.multimodal_encoder.bert(input_ids=tokens.input_ids, attention_mask=tokens.attention_mask).last_hidden_state
        batch[key] = build_feature(self.contra_head_t, subtitle=hidden)
        return batch[key]
    ref = getattr(type(self), "_batch_get_reference", None)
    if ref is None:
        raise KeyError(f"vast_b200.batch_get: key {key!r} is not a contrastive feature and no reference batch_get was "
                       "installed (vast_b200.install(VAST) keeps the original method for every other key)")
    return ref(self, batch, key)


def install(model_cls):
    """Bind the drop-ins onto the reference model class in one call:
        import vast_b200; vast_b200.install(VAST)
    `forward_ret` (model/vast.py:383-483), `batch_get` (feature keys, :221-312) and `compute_slice_scores` (:373-380:
    fused Match_head) are replaced; the reference's own
    `batch_get` stays reachable for every other key."""
    from .contrastive import forward_ret
    if getattr(model_cls, "batch_get", None) is not batch_get:
        model_cls._batch_get_reference = getattr(model_cls, "batch_get", None)
        model_cls.batch_get = batch_get
    model_cls.forward_ret = forward_ret
    model_cls.compute_slice_scores = compute_slice_scores
    return model_cls
