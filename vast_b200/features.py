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
