"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL
reference (/root/reference, read-only) on seeded synthetic inputs.

Run in the authoring container only:   python -m oracle.make_golden
The reference has no tests/fixtures of its own for this path, so these vectors
(outputs of the reference's own functions) are what pins the oracle and the CUDA path.

What is driven (reference file:line):
  VAST.forward_ret                 model/vast.py:383-464   (loss_itc, grads, ITM inputs = sampled negatives)
  VAST.batch_get('feat_vas')       model/vast.py:269-279   (+ pool_* general_module.py:426-449)
  evaluation_mm.compute_metric_ret evaluation/evaluation_mm.py:326-380
  evaluation_mm.refine_score_matrix evaluation/evaluation_mm.py:253-319
  concat_all_gather / ddp_allgather / all_gather_with_grad   utils/distributed.py:33-66,133-149
  evaluation_mm.evaluate_ret       evaluation/evaluation_mm.py:171-251   (val_log of a stubbed two-task evaluation)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_loader as R  # noqa: E402


def synth_feats(n, d, seed, noise=0.8):
    """t = randn, c = t + 0.8 randn, both L2-normalised."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n, d, generator=g)
    c = t + noise * torch.randn(n, d, generator=g)
    return F.normalize(t, dim=-1), F.normalize(c, dim=-1)


class Recorder:
    """Wraps the stub cross-encoder to capture what forward_ret feeds the ITM head."""

    def __init__(self, inner):
        self.inner = inner
        self.calls = []

    def __call__(self, input_ids=None, attention_mask=None, encoder_hidden_states=None):
        self.calls.append((input_ids.clone(), attention_mask.clone(), encoder_hidden_states.detach().clone()))
        return self.inner(input_ids=input_ids, attention_mask=attention_mask,
                          encoder_hidden_states=encoder_hidden_states)


def run_forward_ret(ft, fc, cond, ids, mask, temp, seed, task="ret%tvas"):
    from easydict import EasyDict as edict
    m = R.make_stub_model(contra_temp=temp, hidden=cond.shape[-1])
    rec = Recorder(m.multimodal_encoder.bert)
    m.multimodal_encoder = types.SimpleNamespace(bert=rec)
    ft = ft.clone().requires_grad_()
    fc = fc.clone().requires_grad_()
    batch = edict(raw_captions=["x"] * ft.shape[0])
    batch["feat_t"] = ft
    batch["feat_" + task.split("%")[1][1:]] = fc
    batch["condition_feats_" + task.split("%")[1][1:]] = cond
    batch["caption_tokens"] = edict(input_ids=ids, attention_mask=mask)
    torch.manual_seed(seed)
    out = m.forward_ret(batch, task, compute_loss=True)
    out["loss_itc"].backward()
    ids1, att1, cond3 = rec.calls[0]
    return dict(loss_itc=out["loss_itc"].detach().numpy(), loss_itm=out["loss_itm"].detach().numpy(),
                grad_t=ft.grad.numpy(), grad_cond=fc.grad.numpy(), grad_temp=m.contra_temp.grad.numpy(),
                input_ids_1=ids1.numpy(), attention_mask_1=att1.numpy(), condition_feats_3=cond3.numpy())


def reproduce_exponentials(seed, bs, n):
    """The Exp(1) noise torch.multinomial(w,1) consumes inside forward_ret after
    torch.manual_seed(seed): one exponential_() fill of N per call, t2cond rows then cond2t rows."""
    torch.manual_seed(seed)
    e = torch.empty(2, bs, n)
    for d in range(2):
        for b in range(bs):
            e[d, b] = torch.empty(n).exponential_(1)
    return e.numpy()


def golden_omc_w1():
    bs, d, s, h, l = 64, 512, 6, 16, 10  # cfg1 feature shape; S/H small (gather is layout-only)
    ft, fc = synth_feats(bs, d, 1234)
    g = torch.Generator().manual_seed(99)
    cond = torch.randn(bs, s, h, generator=g)
    ids = torch.randint(0, 30522, (bs, l), generator=g)
    mask = (torch.rand(bs, l, generator=g) > 0.2).long()
    seed = 4321
    out = run_forward_ret(ft, fc, cond, ids, mask, 0.07, seed)
    expo = reproduce_exponentials(seed, bs, bs)
    # confirm multinomial == exp-race with this noise, and recover the sampled indices
    z1 = (fc @ ft.T / 0.07)
    z2 = (ft @ fc.T / 0.07)
    w_t2c = F.softmax(z2, dim=1) + 1e-4
    w_t2c.fill_diagonal_(0)
    w_c2t = F.softmax(z1, dim=1) + 1e-4
    w_c2t.fill_diagonal_(0)
    neg_t2c = (w_t2c / torch.from_numpy(expo[0])).argmax(dim=1)
    neg_c2t = (w_c2t / torch.from_numpy(expo[1])).argmax(dim=1)
    assert np.array_equal(out["condition_feats_3"][bs:2 * bs], cond[neg_t2c].numpy()), "exp-race != multinomial (t2cond)"
    assert np.array_equal(out["input_ids_1"][2 * bs:], ids[neg_c2t].numpy()), "exp-race != multinomial (cond2t)"
    np.savez_compressed(os.path.join(GOLD, "omc_w1.npz"), feat_t=ft.numpy(), feat_cond=fc.numpy(),
                        cond=cond.numpy(), input_ids=ids.numpy(), attention_mask=mask.numpy(),
                        contra_temp=np.float32(0.07), expo=expo, neg_t2cond=neg_t2c.numpy(),
                        neg_cond2t=neg_c2t.numpy(), **out)
    print("omc_w1: loss_itc", out["loss_itc"], "grad_temp", out["grad_temp"])


def _w2_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ns = R.load(init_pg=False)
    bs, d, s, h, l = 16, 64, 4, 16, 6
    ft_all, fc_all = synth_feats(bs * world, d, 777)
    g = torch.Generator().manual_seed(55)
    cond_all = torch.randn(bs * world, s, h, generator=g)
    ids_all = torch.randint(0, 30522, (bs * world, l), generator=g)
    mask_all = torch.ones(bs * world, l, dtype=torch.long)
    sl = slice(rank * bs, (rank + 1) * bs)
    seed = 100 + rank
    out = run_forward_ret(ft_all[sl], fc_all[sl], cond_all[sl], ids_all[sl], mask_all[sl], 0.05, seed)
    out["expo"] = reproduce_exponentials(seed, bs, bs * world)
    # collectives
    x = torch.arange(rank * 10, rank * 10 + 6, dtype=torch.float32).reshape(3, 2)
    out["concat_all_gather"] = ns.D.concat_all_gather(x).numpy()
    ragged = torch.arange((rank + 2) * 3, dtype=torch.float32).reshape(rank + 2, 3) + 100 * rank
    out["ddp_allgather"] = ns.D.ddp_allgather(ragged).numpy()
    out["all_gather_list"] = np.array([len(v) for v in ns.D.all_gather_list(list(range(rank + 1)))])
    xg = x.clone().requires_grad_()
    yg = ns.D.all_gather_with_grad(xg)
    (yg * torch.arange(yg.numel(), dtype=torch.float32).reshape(yg.shape)).sum().backward()
    out["agwg_out"] = yg.detach().numpy()
    out["agwg_grad"] = xg.grad.numpy()
    # eval: refine_score_matrix with column shards (each rank holds its own videos)
    nt, nv_r = 24, 6 + rank  # ragged shards
    nv = sum(6 + r for r in range(world))
    et, _ = synth_feats(nt, d, 31)
    _, ev = synth_feats(nv, d, 32)
    ev[:nt][: min(nt, nv)] = F.normalize(et[: min(nt, nv)] + 0.5 * ev[: min(nt, nv)], dim=-1)
    score = et @ ev.T
    g2 = torch.Generator().manual_seed(66)
    econd_all = torch.randn(nv, s, h, generator=g2)
    eids = torch.randint(0, 30522, (nt, l), generator=g2)
    emask = torch.ones(nt, l, dtype=torch.long)
    start = sum(6 + r for r in range(rank))
    m = R.make_stub_model(hidden=h)
    for direction in ("forward", "backward"):
        r = ns.E.refine_score_matrix(econd_all[start:start + nv_r], eids, emask, score, m, 4, direction=direction)
        out["refine_" + direction] = r.numpy()
    out.update(eval_feat_t=et.numpy(), eval_feat_v=ev.numpy(), eval_cond=econd_all.numpy(),
               eval_ids=eids.numpy(), eval_mask=emask.numpy())
    out.update(feat_t_all=ft_all.numpy(), feat_cond_all=fc_all.numpy(), cond_all=cond_all.numpy(),
               ids_all=ids_all.numpy(), mask_all=mask_all.numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def golden_w2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_w2_worker, args=(r, world, 29611, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in range(world))
    for p in procs:
        p.join()
    flat = {}
    shared = ("feat_t_all", "feat_cond_all", "cond_all", "ids_all", "mask_all", "eval_feat_t",
              "eval_feat_v", "eval_cond", "eval_ids", "eval_mask")
    for r in range(world):
        for k, v in res[r].items():
            if k in shared:
                flat[k] = v
            else:
                flat[f"r{r}_{k}"] = v
    flat["contra_temp"] = np.float32(0.05)
    np.savez_compressed(os.path.join(GOLD, "dist_w2.npz"), **flat)
    print("dist_w2: loss r0/r1", flat["r0_loss_itc"], flat["r1_loss_itc"])


def golden_retrieval():
    ns = R.load()
    out = {}
    # (a) 1 caption / video
    nt = nv = 96
    d, s, h, l = 64, 4, 16, 6
    et, ev = synth_feats(nt, d, 2024, noise=4.0)
    score = et @ ev.T
    ids = list(range(nv))
    out["a_feat_t"], out["a_feat_v"] = et.numpy(), ev.numpy()
    for direction in ("forward", "backward"):
        log = ns.E.compute_metric_ret(score, ids, ids, direction)
        out[f"a_metric_{direction}"] = np.array([log[f"{direction}_r1"], log[f"{direction}_ravg"]])
        out[f"a_recall_{direction}"] = np.array(log[f"{direction}_recall"])
    # (b) multi-caption: 5 texts per video, string ids
    nv_b, per = 20, 5
    nt_b = nv_b * per
    _, evb = synth_feats(nv_b, d, 2025)
    g = torch.Generator().manual_seed(5)
    etb = F.normalize(evb.repeat_interleave(per, dim=0) + 1.2 * torch.randn(nt_b, d, generator=g), dim=-1)
    score_b = etb @ evb.T
    ids_b = [f"video{i}" for i in range(nv_b)]
    ids_txt_b = [f"video{i // per}" for i in range(nt_b)]
    out["b_feat_t"], out["b_feat_v"] = etb.numpy(), evb.numpy()
    out["b_per"] = np.int64(per)
    for direction in ("forward", "backward"):
        log = ns.E.compute_metric_ret(score_b, ids_b, ids_txt_b, direction)
        out[f"b_metric_{direction}"] = np.array([log[f"{direction}_r1"], log[f"{direction}_ravg"]])
        out[f"b_recall_{direction}"] = np.array(log[f"{direction}_recall"])
    # (c) refine_score_matrix W=1, both directions; k >= 10 so R@10 does not depend on torch's unspecified
    #     tie order among the exact zeros of the refined matrix
    g2 = torch.Generator().manual_seed(8)
    cond = torch.randn(nv_b, s, h, generator=g2)
    tids = torch.randint(0, 30522, (nt_b, l), generator=g2)
    tmask = (torch.rand(nt_b, l, generator=g2) > 0.1).long()
    m = R.make_stub_model(hidden=h)
    out.update(c_cond=cond.numpy(), c_ids=tids.numpy(), c_mask=tmask.numpy())
    for direction, k in (("forward", 12), ("backward", 30)):
        r = ns.E.refine_score_matrix(cond, tids, tmask, score_b, m, k, direction=direction)
        out[f"c_refine_{direction}"] = r.numpy()
        log = ns.E.compute_metric_ret(r, ids_b, ids_txt_b, direction)
        out[f"c_recall_{direction}"] = np.array(log[f"{direction}_recall"])
    out["c_k"] = np.array([12, 30])
    np.savez_compressed(os.path.join(GOLD, "retrieval.npz"), **out)
    print("retrieval:", out["a_recall_forward"], out["b_recall_backward"], out["c_recall_forward"])


def golden_features():
    """batch_get('feat_vas') (model/vast.py:269-279) + pool_* on random encoder outputs."""
    from easydict import EasyDict as edict
    m = R.make_stub_model()
    b, n, tok, cv, na, ta, ca, ls, cs, d = 5, 3, 7, 48, 2, 9, 24, 6, 24, 32
    g = torch.Generator().manual_seed(17)
    vis = torch.randn(b, n, tok, cv, generator=g)
    aud = torch.randn(b, na, ta, ca, generator=g)
    sub = torch.randn(b, ls, cs, generator=g)
    lin = nn.Linear(cv + ca + cs, d)
    with torch.no_grad():
        lin.weight.copy_(0.05 * torch.randn(d, cv + ca + cs, generator=g))
        lin.bias.copy_(0.01 * torch.randn(d, generator=g))
    m.contra_head_vas = lin
    batch = edict()
    batch["vision_output"], batch["audio_output"], batch["subtitle_output"] = vis, aud, sub
    with torch.no_grad():
        feat = m.batch_get(batch, "feat_vas")
        pv = m.pool_vision_for_contra(vis)
        pa = m.pool_audio_for_contra(aud)
        ps = m.pool_text_for_contra(sub)
        m.config.vision_encoder_type = "swin_base"
        m.config.audio_encoder_type = "ast"
        pv_swin = m.pool_vision_for_contra(vis)
        pa_ast = m.pool_audio_for_contra(aud)
    np.savez_compressed(os.path.join(GOLD, "features.npz"), vision=vis.numpy(), audio=aud.numpy(),
                        subtitle=sub.numpy(), weight=lin.weight.detach().numpy(), bias=lin.bias.detach().numpy(),
                        feat_vas=feat.numpy(), pool_v=pv.numpy(), pool_a=pa.numpy(), pool_s=ps.numpy(),
                        pool_v_swin=pv_swin.numpy(), pool_a_ast=pa_ast.numpy())
    print("features: feat_vas", tuple(feat.shape))


def golden_match_head():
    """The reference's own Match_head (model/general_module.py:34-42: Linear -> GELU(erf) -> LayerNorm(eps 1e-12) ->
    Linear(2)) and the score `F.softmax(itm_head(cls), dim=1)[:, 1]` of compute_slice_scores (model/vast.py:378) on
    seeded cls tokens, fp32, for the real hidden size 768 and a ragged one."""
    G = R.load().G
    out = {}
    for tag, hidden, b in (("h768", 768, 37), ("h96", 96, 130)):
        torch.manual_seed(11 + hidden)
        head = G.Match_head(hidden)
        with torch.no_grad():
            head.layernorm.weight.copy_(1.0 + 0.1 * torch.randn(hidden))
            head.layernorm.bias.copy_(0.05 * torch.randn(hidden))
            cls = torch.randn(b, hidden) * 0.7
            logits = head(cls)
            score = F.softmax(logits, dim=1)[:, 1]
        n = lambda x: x.detach().numpy()
        out.update({f"{tag}_cls": n(cls), f"{tag}_w1": n(head.linear1.weight), f"{tag}_b1": n(head.linear1.bias),
                    f"{tag}_gamma": n(head.layernorm.weight), f"{tag}_beta": n(head.layernorm.bias),
                    f"{tag}_w2": n(head.linear2.weight), f"{tag}_b2": n(head.linear2.bias),
                    f"{tag}_logits": n(logits), f"{tag}_score": n(score)})
    np.savez_compressed(os.path.join(GOLD, "match_head.npz"), **out)
    print("match_head:", out["h768_score"][:4])


def golden_evaluate_ret():
    """evaluation_mm.evaluate_ret (evaluation/evaluation_mm.py:171-251) end to end: a stub model that returns
    pre-seeded evaluation dicts per batch (what VAST.forward_ret(compute_loss=False) returns, model/vast.py:468-483),
    two sub-tasks, multi-caption ids (5 texts per video, flattened at :195-201), bidirectional metrics, ITM re-rank
    with k=16 (>= 10, so R@10 is independent of torch's tie order among the zeros of the refined matrix)."""
    import json
    from easydict import EasyDict as edict
    ns = R.load()
    nv, per, d, s, h, l, nb = 24, 5, 64, 4, 16, 6, 3
    nt = nv * per
    g = torch.Generator().manual_seed(77)
    conds, feats_v = {}, {}
    for i, task in enumerate(("tv", "tvas")):
        _, feats_v[task] = synth_feats(nv, d, 3030 + i)
        conds[task] = torch.randn(nv, s, h, generator=g)
    base = 0.5 * (feats_v["tv"] + feats_v["tvas"])
    feat_t = F.normalize(base.repeat_interleave(per, dim=0) + 1.0 * torch.randn(nt, d, generator=g), dim=-1)
    tids = torch.randint(0, 30522, (nt, l), generator=g)
    tmask = (torch.rand(nt, l, generator=g) > 0.1).long()
    ids = [f"vid{i}" for i in range(nv)]
    stub = R.make_stub_model(hidden=h)
    stub.config.itm_rerank_num = 16
    stub.config.ret_bidirection_evaluation = True

    class Model:
        config = stub.config
        compute_slice_scores = stub.compute_slice_scores

        def __call__(self, batch, tasks, compute_loss=False):
            assert compute_loss is False
            return batch["ev"]

    loader = []
    vb = nv // nb
    for b in range(nb):
        vs = slice(b * vb, (b + 1) * vb)
        ts = slice(b * vb * per, (b + 1) * vb * per)
        ev = {"feat_t": feat_t[ts], "input_ids": tids[ts], "attention_mask": tmask[ts]}
        for task in ("tv", "tvas"):
            ev[f"feat_cond_{task}"] = feats_v[task][vs]
            ev[f"condition_feats_{task}"] = conds[task][vs]
        loader.append({"ids": ids[vs], "ids_txt": [[v] * per for v in ids[vs]], "ev": ev})
    log = ns.E.evaluate_ret(Model(), "ret%tv%tvas", loader, 0)
    out = dict(feat_t=feat_t.numpy(), ids_tok=tids.numpy(), mask=tmask.numpy(), per=np.int64(per), nb=np.int64(nb),
               log_json=np.array(json.dumps(log, sort_keys=True)))
    for task in ("tv", "tvas"):
        out[f"feat_v_{task}"] = feats_v[task].numpy()
        out[f"cond_{task}"] = conds[task].numpy()
    np.savez_compressed(os.path.join(GOLD, "evaluate_ret.npz"), **out)
    print("evaluate_ret:", json.dumps(log, sort_keys=True)[:300])


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    golden_w2()          # spawns its own process group first (port separate from the W=1 group)
    R.load()
    golden_omc_w1()
    golden_retrieval()
    golden_features()
    golden_match_head()
    golden_evaluate_ret()
    print("golden vectors written to", GOLD)
