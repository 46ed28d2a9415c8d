"""The two fused heads either side of the contrastive path (vast_project_normalize, vast_match_head) through the C-ABI,
against the oracle and the reference's golden outputs."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import spec

pytestmark = pytest.mark.gpu


def cu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t.to(dtype) if dtype is not None else t


class MatchHead(nn.Module):
    """same structure / attribute names as the reference's Match_head (general_module.py:34-42)"""

    def __init__(self, g, tag):
        super().__init__()
        h = g[f"{tag}_w1"].shape[0]
        self.linear1, self.layernorm, self.linear2 = nn.Linear(h, h), nn.LayerNorm(h, eps=1e-12), nn.Linear(h, 2)
        with torch.no_grad():
            self.linear1.weight.copy_(torch.from_numpy(g[f"{tag}_w1"]))
            self.linear1.bias.copy_(torch.from_numpy(g[f"{tag}_b1"]))
            self.layernorm.weight.copy_(torch.from_numpy(g[f"{tag}_gamma"]))
            self.layernorm.bias.copy_(torch.from_numpy(g[f"{tag}_beta"]))
            self.linear2.weight.copy_(torch.from_numpy(g[f"{tag}_w2"]))
            self.linear2.bias.copy_(torch.from_numpy(g[f"{tag}_b2"]))

    def forward(self, x):
        h = self.linear1(x)
        return self.linear2(self.layernorm(h * 0.5 * (1.0 + torch.erf(h / 2.0 ** 0.5))))


@pytest.mark.parametrize("tag", ["h768", "h96"])
def test_match_head_golden(golden, tag):
    """fp32 inputs (fp32-grade split operands): score and logits == the reference's Match_head + softmax[:, 1]."""
    import vast_b200
    g = golden("match_head")
    head = MatchHead(g, tag).cuda()
    score, logits = vast_b200.match_head_scores(head, cu(g[f"{tag}_cls"]), want_logits=True)
    np.testing.assert_allclose(logits.cpu().numpy(), g[f"{tag}_logits"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(score.cpu().numpy(), g[f"{tag}_score"], rtol=2e-4, atol=2e-6)
    # bf16-in / fp32-accumulate mode against the oracle on bf16-rounded cls and W1
    sb = vast_b200.match_head_scores(head, cu(g[f"{tag}_cls"]), mode="bf16")
    rb = lambda k: torch.from_numpy(g[f"{tag}_{k}"]).bfloat16().float().numpy()
    _, want = spec.match_head(rb("cls"), rb("w1"), *(g[f"{tag}_{k}"] for k in ("b1", "gamma", "beta", "w2", "b2")))
    np.testing.assert_allclose(sb.cpu().numpy(), want, rtol=1e-3, atol=1e-5)


def test_match_head_large_batch_and_dropin(golden):
    """thousands of pairs per call (the batched re-rank) and the compute_slice_scores drop-in."""
    import vast_b200
    g = golden("match_head")
    head = MatchHead(g, "h768").cuda()
    gen = torch.Generator().manual_seed(5)
    cls = torch.randn(5000, 768, generator=gen) * 0.7
    got = vast_b200.match_head_scores(head, cls.cuda())
    _, want = spec.match_head(cls.numpy(), *(g[f"h768_{k}"] for k in ("w1", "b1", "gamma", "beta", "w2", "b2")))
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-4, atol=2e-6)
    for dt, tol in ((torch.float16, 2e-3), (torch.bfloat16, 2e-2)):   # 16-bit cls tokens: exact products of the rounded inputs
        got16 = vast_b200.match_head_scores(head, cls.to(dt).cuda())
        _, w16 = spec.match_head(cls.to(dt).float().numpy(), *(g[f"h768_{k}"] for k in ("w1", "b1", "gamma", "beta", "w2", "b2")))
        np.testing.assert_allclose(got16.cpu().numpy(), w16, rtol=tol, atol=tol * 1e-2)

    class Model:
        compute_slice_scores = vast_b200.compute_slice_scores

        def __init__(self):
            self.itm_head = head
            self.multimodal_encoder = types.SimpleNamespace(bert=self._bert)

        @staticmethod
        def _bert(input_ids, attention_mask, encoder_hidden_states):
            h = encoder_hidden_states[:, :input_ids.shape[1], :768].float() + 0.01 * input_ids[..., None].float() * attention_mask[..., None]
            return types.SimpleNamespace(last_hidden_state=h)
    cond = torch.randn(25, 12, 768, generator=gen).cuda()
    ids = torch.randint(0, 100, (25, 8), generator=gen).cuda()
    mask = torch.ones(25, 8, dtype=torch.int64).cuda()
    s = Model().compute_slice_scores(cond, ids, mask)
    ref = torch.softmax(head(Model._bert(ids, mask, cond).last_hidden_state[:, 0]), dim=1)[:, 1]
    assert torch.allclose(s, ref, rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("rows,k,d,bias", [(5, 96, 32, True), (512, 2944, 1024, True), (300, 768, 520, False), (4096, 2944, 1024, True),
                                           (129, 8, 8, True)])
def test_project_normalize_vs_oracle(rows, k, d, bias):
    """vast_project_normalize (Linear + bias + F.normalize in one kernel) vs the oracle's fuse_feature chain, fp32-grade
    from fp32 inputs; bf16 slot and inv_norm outputs; rows / columns that do not fill tiles."""
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(rows + k)
    x = torch.randn(rows, k, generator=gen)
    w = torch.randn(d, k, generator=gen) * 0.05
    b = torch.randn(d, generator=gen) * 0.1 if bias else None
    slot = torch.zeros(rows, 2 * d, dtype=torch.bfloat16, device="cuda")
    y, inv, _ = ops.project_normalize(x.cuda(), w.cuda(), None if b is None else b.cuda(), out16=slot[:, d:])
    u = x.double().numpy() @ w.double().numpy().T + (0 if b is None else b.double().numpy())
    want = spec.l2_normalize(u)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(inv.cpu().numpy(), 1.0 / np.sqrt((u * u).sum(1)), rtol=1e-4)
    assert torch.equal(slot[:, d:], y.bfloat16()) and not slot[:, :d].any()
    # bf16-in / fp32-accumulate mode vs the oracle on bf16-rounded operands
    yb, _, _ = ops.project_normalize(x.cuda().bfloat16(), w.cuda().bfloat16(), None if b is None else b.cuda())
    ub = x.bfloat16().double().numpy() @ w.bfloat16().double().numpy().T + (0 if b is None else b.double().numpy())
    np.testing.assert_allclose(yb.cpu().numpy(), spec.l2_normalize(ub), rtol=2e-4, atol=2e-6)


def test_build_feature_fused_equals_unfused_with_grads(golden):
    """build_feature through the fused projection == the unfused chain (library GEMM + vast_l2norm) and the reference's
    golden feat_vas; gradients w.r.t. encoder outputs, weight and bias agree with torch autograd of the same chain."""
    import vast_b200
    g = golden("features")
    outs = {}
    for fused in (True, False):
        lin = nn.Linear(g["weight"].shape[1], g["weight"].shape[0]).cuda()
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(g["weight"]))
            lin.bias.copy_(torch.from_numpy(g["bias"]))
        vis, aud, sub = (cu(g[k]).requires_grad_() for k in ("vision", "audio", "subtitle"))
        feat = vast_b200.build_feature(lin, vis, aud, sub, "evaclip01_giant", "beats", fused=fused)
        np.testing.assert_allclose(feat.detach().cpu().numpy(), g["feat_vas"], rtol=2e-4, atol=2e-6)
        go = torch.randn(feat.shape, generator=torch.Generator().manual_seed(3)).cuda()
        feat.backward(go)
        outs[fused] = [t.grad.clone() for t in (vis, aud, sub, lin.weight, lin.bias)]
    for a, b in zip(outs[True], outs[False]):
        assert torch.allclose(a, b, rtol=2e-3, atol=1e-6)
