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


def build_feature(head: torch.nn.Module, vision=None, audio=None, subtitle=None, vision_encoder_type="evaclip",
                  audio_encoder_type="beats") -> torch.Tensor:
    """feat_{v,a,s,va,vs,vas} (vast.py:221-279): pooled = pool_concat(...); feat = normalize(head(pooled))."""
    pooled = pool_concat(vision, audio, subtitle, vision_encoder_type, audio_encoder_type)
    return l2_normalize(head(pooled).float())


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
    `forward_ret` (model/vast.py:383-483) and `batch_get` (feature keys, :221-312) are replaced; the reference's own
    `batch_get` stays reachable for every other key."""
    from .contrastive import forward_ret
    if getattr(model_cls, "batch_get", None) is not batch_get:
        model_cls._batch_get_reference = getattr(model_cls, "batch_get", None)
        model_cls.batch_get = batch_get
    model_cls.forward_ret = forward_ret
    return model_cls
