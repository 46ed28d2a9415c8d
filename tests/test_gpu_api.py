"""Python drop-in surface (autograd wrappers, forward_ret, feature build) on the GPU vs oracle / golden."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import spec

pytestmark = pytest.mark.gpu


def bf16_round(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).bfloat16().float().numpy()


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_omc_autograd_matches_oracle(golden):
    import vast_b200
    g = golden("omc_w1")
    ft = torch.from_numpy(g["feat_t"]).cuda().requires_grad_()
    fc = torch.from_numpy(g["feat_cond"]).cuda().requires_grad_()
    temp = nn.Parameter(torch.tensor(0.07, device="cuda"))
    loss, neg_text, neg_cond = vast_b200.omc_loss_and_negatives(fc, ft, temp, rank=0, world_size=1)
    (3.0 * loss).backward()   # non-unit upstream gradient
    o = spec.omc_loss(bf16_round(g["feat_cond"]), bf16_round(g["feat_t"]), bf16_round(g["feat_t"]),
                      bf16_round(g["feat_cond"]), 0.07, grad_out=3.0)
    assert abs(loss.item() - o["loss"]) < 1e-3 * o["loss"]
    assert rel(ft.grad.cpu().numpy(), o["grad_t"]) < 1e-3
    assert rel(fc.grad.cpu().numpy(), o["grad_cond"]) < 1e-3
    assert abs(temp.grad.item() - o["grad_temp"]) < 1e-3 * abs(o["grad_temp"])
    assert neg_text.dtype == torch.int64 and neg_text.shape == (64,) and not neg_text.requires_grad
    assert (neg_text.cpu() != torch.arange(64)).all() and (neg_cond.cpu() != torch.arange(64)).all()


def test_graphed_step_matches_eager_and_oracle(golden):
    """vast_b200.OmcGraphStep: the fused step as one CUDA-graph launch -- same loss / gradients as the oracle, fresh
    negatives every replay, live temperature parameter, and a loud error for gradients that were overwritten."""
    import vast_b200
    g = golden("omc_w1")
    temp = nn.Parameter(torch.tensor(0.07, device="cuda"))
    step = vast_b200.OmcGraphStep(64, g["feat_t"].shape[1], temp, rank=0, world_size=1)
    tb, cb = bf16_round(g["feat_t"]), bf16_round(g["feat_cond"])
    negs = []
    for it, tau in enumerate((0.07, 0.07, 0.05)):
        with torch.no_grad():
            temp.fill_(tau)                      # in-place update (what an optimizer does) is seen by the graph
        temp.grad = None
        ft = torch.from_numpy(g["feat_t"]).cuda().requires_grad_()
        fc = torch.from_numpy(g["feat_cond"]).cuda().requires_grad_()
        loss, neg_text, neg_cond = step(fc, ft)
        (2.0 * loss).backward()
        o = spec.omc_loss(cb, tb, tb, cb, tau, grad_out=2.0)
        assert abs(loss.item() - o["loss"]) < 1e-3 * o["loss"]
        assert rel(ft.grad.cpu().numpy(), o["grad_t"]) < 1e-3
        assert rel(fc.grad.cpu().numpy(), o["grad_cond"]) < 1e-3
        assert abs(temp.grad.item() - o["grad_temp"]) < 1e-3 * abs(o["grad_temp"])
        assert (neg_text.cpu() != torch.arange(64)).all() and (neg_cond.cpu() != torch.arange(64)).all()
        negs.append(neg_text.clone())
    assert not torch.equal(negs[0], negs[1])
    # a loader may fill the step's own input buffers (no device-to-device copy in the call)
    fc_in, ft_in = step.inputs()
    ft_in.copy_(torch.from_numpy(g["feat_t"]))
    fc_in.copy_(torch.from_numpy(g["feat_cond"]))
    loss_in, _, _ = step(fc_in, ft_in)
    assert abs(loss_in.item() - o["loss"]) < 1e-3 * o["loss"]
    # outputs of call i are overwritten by call i+2: backward of the stale loss must raise
    ft = torch.from_numpy(g["feat_t"]).cuda().requires_grad_()
    fc = torch.from_numpy(g["feat_cond"]).cuda().requires_grad_()
    stale, _, _ = step(fc, ft)
    step(fc, ft)
    step(fc, ft)
    with pytest.raises(RuntimeError):
        stale.backward()
    with pytest.raises(RuntimeError):
        step(fc[:10], ft[:10])


class _Stub(nn.Module):
    """Stand-in for the VAST module: what forward_ret touches (model/vast.py:383-464)."""

    def __init__(self, hidden=16):
        super().__init__()
        self.contra_temp = nn.Parameter(torch.tensor(0.07))
        self.itm_ratio = 0.1
        gen = torch.Generator().manual_seed(0)
        self.emb = nn.Parameter(torch.randn(30522, hidden, generator=gen) * 0.1)
        self.proj = nn.Parameter(torch.randn(hidden, hidden, generator=gen) * 0.3)
        self.w = nn.Parameter(torch.randn(hidden, 2, generator=torch.Generator().manual_seed(1)))
        self.hidden = hidden
        self.seen = None
        self.multimodal_encoder = types.SimpleNamespace(bert=self._bert)
        self.itm_head = lambda x: x.float() @ self.w

    def _bert(self, input_ids=None, attention_mask=None, encoder_hidden_states=None):
        self.seen = (input_ids, attention_mask, encoder_hidden_states)
        x = self.emb[input_ids] * attention_mask.unsqueeze(-1).to(self.emb.dtype)
        ctx = encoder_hidden_states.float().mean(dim=1, keepdim=True)[..., :self.hidden]
        return types.SimpleNamespace(last_hidden_state=torch.tanh((x + ctx) @ self.proj))

    def batch_get(self, batch, key):
        return batch[key]


class _Batch(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def test_forward_ret_dropin(golden):
    import vast_b200
    g = golden("omc_w1")
    bs = 64
    m = _Stub().cuda()
    m.forward_ret = types.MethodType(vast_b200.forward_ret, m)
    ft = torch.from_numpy(g["feat_t"]).cuda().requires_grad_()
    fc = torch.from_numpy(g["feat_cond"]).cuda().requires_grad_()
    cond = torch.from_numpy(g["cond"]).cuda().requires_grad_()
    ids, mask = torch.from_numpy(g["input_ids"]).cuda(), torch.from_numpy(g["attention_mask"]).cuda()
    batch = _Batch(raw_captions=["x"] * bs, feat_t=ft, feat_vas=fc, condition_feats_vas=cond,
                   caption_tokens=_Batch(input_ids=ids, attention_mask=mask))
    out = m.forward_ret(batch, "ret%tvas", compute_loss=True)
    assert set(out) == {"loss_itc", "loss_itm"}
    assert abs(out["loss_itc"].item() - float(g["loss_itc"])) < 1e-2 * float(g["loss_itc"])
    (out["loss_itc"] + out["loss_itm"]).backward()
    assert rel(ft.grad.cpu().numpy(), g["grad_t"]) < 2e-2       # ITM does not reach feat_t
    assert cond.grad is not None and cond.grad.abs().sum().item() > 0
    ids1, att1, cond3 = m.seen
    assert ids1.shape == (3 * bs, ids.shape[1]) and cond3.shape == (3 * bs,) + tuple(cond.shape[1:])
    assert torch.equal(ids1[:bs], ids) and torch.equal(ids1[bs:2 * bs], ids) and torch.equal(cond3[2 * bs:], cond.detach())
    assert torch.equal(cond3[:bs], cond.detach())
    # the negative rows are real rows of the collated tensors and never the positive
    neg_rows = cond3[bs:2 * bs]
    match = (neg_rows[:, None] == cond.detach()[None]).flatten(2).all(-1)
    assert (match.sum(1) >= 1).all() and not match.diagonal().any()
    ev = m.forward_ret(batch, "ret%tvas", compute_loss=False)
    assert set(ev) == {"feat_t", "input_ids", "attention_mask", "feat_cond_tvas", "condition_feats_tvas"}


def test_build_feature_golden_and_grad(golden):
    import vast_b200
    g = golden("features")
    lin = nn.Linear(g["weight"].shape[1], g["weight"].shape[0]).cuda()
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(g["weight"]))
        lin.bias.copy_(torch.from_numpy(g["bias"]))
    vis = torch.from_numpy(g["vision"]).cuda().requires_grad_()
    aud = torch.from_numpy(g["audio"]).cuda().requires_grad_()
    sub = torch.from_numpy(g["subtitle"]).cuda().requires_grad_()
    feat = vast_b200.build_feature(lin, vis, aud, sub, "evaclip01_giant", "beats")
    np.testing.assert_allclose(feat.detach().cpu().numpy(), g["feat_vas"], rtol=2e-4, atol=2e-6)
    go = torch.randn(feat.shape, generator=torch.Generator().manual_seed(3)).cuda()
    feat.backward(go)
    # torch reference of the same chain
    v2, a2, s2 = (torch.from_numpy(g[k]).requires_grad_() for k in ("vision", "audio", "subtitle"))
    lin_c = nn.Linear(g["weight"].shape[1], g["weight"].shape[0])
    with torch.no_grad():
        lin_c.weight.copy_(torch.from_numpy(g["weight"]))
        lin_c.bias.copy_(torch.from_numpy(g["bias"]))
    ref = torch.nn.functional.normalize(lin_c(torch.cat([v2[:, :, 0].mean(1), a2.mean(2).mean(1), s2[:, 0]], 1)), dim=-1)
    ref.backward(go.cpu())
    for got, want in ((vis.grad, v2.grad), (aud.grad, a2.grad), (sub.grad, s2.grad)):
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-3, atol=1e-6)


def test_batch_get_dropin_golden_memoised_and_delegating(golden):
    """vast_b200.batch_get bound to a model class in place of VAST.batch_get (model/vast.py:82-312): feat_vas equals
    the reference's own batch_get('feat_vas') golden, results are memoised in the batch (:82-83), encoder outputs come
    through self.batch_get, every non-feature key goes to the method it replaced, and the single-modality / caption
    keys follow the reference chain pool -> head -> normalize."""
    import vast_b200
    g = golden("features")
    calls = []

    class Contra(nn.Module):            # general_module.py:26-31
        def __init__(self, i, o):
            super().__init__()
            self.linear = nn.Linear(i, o, bias=False)

        def forward(self, x):
            return self.linear(x)

    class Bert:
        def __call__(self, input_ids, attention_mask):
            calls.append("bert")
            h = torch.nn.functional.one_hot(input_ids % 24, 24).float().cuda() * attention_mask[..., None].float()
            return types.SimpleNamespace(last_hidden_state=h)

    class Model:
        config = types.SimpleNamespace(vision_encoder_type="evaclip01_giant", audio_encoder_type="beats")

        def batch_get(self, batch, key):                      # the "reference" method being replaced
            calls.append(key)
            if key == "vision_output":
                batch[key] = torch.from_numpy(g["vision"]).cuda()
            elif key == "omni_caption_tokens":
                batch[key] = types.SimpleNamespace(input_ids=torch.arange(30).reshape(5, 6).cuda(),
                                                   attention_mask=torch.ones(5, 6, dtype=torch.int64).cuda())
            else:
                raise KeyError(key)
            return batch[key]

    vast_b200.install(Model)
    assert Model.batch_get is vast_b200.batch_get and Model.forward_ret is vast_b200.forward_ret
    m = Model()
    lin = nn.Linear(g["weight"].shape[1], g["weight"].shape[0]).cuda()
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(g["weight"]))
        lin.bias.copy_(torch.from_numpy(g["bias"]))
    m.contra_head_vas = lin
    m.contra_head_v = Contra(48, 32).cuda()
    m.contra_head_t = Contra(24, 32).cuda()
    m.multimodal_encoder = types.SimpleNamespace(bert=Bert())
    batch = {"audio_output": torch.from_numpy(g["audio"]).cuda(), "subtitle_output": torch.from_numpy(g["subtitle"]).cuda()}
    feat = m.batch_get(batch, "feat_vas")
    np.testing.assert_allclose(feat.detach().cpu().numpy(), g["feat_vas"], rtol=2e-4, atol=2e-6)
    assert calls == ["vision_output"]                          # fetched through the replaced method, exactly once
    assert m.batch_get(batch, "feat_vas") is feat and "feat_vas" in batch and calls == ["vision_output"]   # memoised
    fv = m.batch_get(batch, "feat_v")                          # reuses the memoised vision_output
    want = torch.nn.functional.normalize(m.contra_head_v(torch.from_numpy(g["pool_v"]).cuda()), dim=-1)
    np.testing.assert_allclose(fv.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=2e-4, atol=2e-6)
    assert calls == ["vision_output"]
    ft = m.batch_get(batch, "feat_t_omni_caption")             # tokens -> bert -> cls -> contra_head_t -> normalize
    hid = Bert()(batch["omni_caption_tokens"].input_ids, batch["omni_caption_tokens"].attention_mask).last_hidden_state
    want = torch.nn.functional.normalize(m.contra_head_t(hid[:, 0]), dim=-1)
    np.testing.assert_allclose(ft.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=2e-4, atol=2e-6)
    with pytest.raises(KeyError):
        m.batch_get(batch, "no_such_key")                      # delegated: the reference method's own error
    # preset keys are returned verbatim like the reference (:82-83)
    sentinel = torch.zeros(1)
    assert m.batch_get({"feat_t": sentinel}, "feat_t") is sentinel
