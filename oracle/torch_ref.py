"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's hot path, operation for
operation (same ATen ops, same order, same RNG consumption), used
  * as the CPU baseline leg of bench.py (`cpu_baseline.kind = "port"`, and `--impl reference`:
    /root/reference itself does not exist on the GPU box), and
  * as a second oracle pinned bit-for-bit to tests/golden (same seed -> same sampled negatives).
Never imported by vast_b200/."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def itc_and_negatives(feat_cond, feat_t, feat_t_all, feat_cond_all, contra_temp, rank=0):
    """model/vast.py:405-440 for one sub-task (W emulated by passing the gathered tensors).
    Returns (loss, neg_idx_t2cond list, neg_idx_cond2t list); consumes the global torch RNG exactly
    like the reference (bs multinomial draws for t2cond, then bs for cond2t)."""
    sim_cond2t = torch.matmul(feat_cond, feat_t_all.permute(1, 0))
    sim_cond2t = sim_cond2t / contra_temp
    sim_t2cond = torch.matmul(feat_t, feat_cond_all.permute(1, 0))
    sim_t2cond = sim_t2cond / contra_temp
    bs = feat_t.size(0)
    targets = torch.linspace(rank * bs, rank * bs + bs - 1, bs, dtype=int)
    loss = (F.cross_entropy(sim_cond2t, targets, label_smoothing=0.1)
            + F.cross_entropy(sim_t2cond, targets, label_smoothing=0.1)) / 2
    with torch.no_grad():
        weights_t2cond = F.softmax(sim_t2cond, dim=1) + 1e-4
        weights_t2cond[:, rank * bs: rank * bs + bs].fill_diagonal_(0)
        weights_cond2t = F.softmax(sim_cond2t, dim=1) + 1e-4
        weights_cond2t[:, rank * bs: rank * bs + bs].fill_diagonal_(0)
    neg_t2cond = [torch.multinomial(weights_t2cond[b], 1).item() for b in range(bs)]
    neg_cond2t = [torch.multinomial(weights_cond2t[b], 1).item() for b in range(bs)]
    return loss, neg_t2cond, neg_cond2t


def contrastive_step(feat_cond, feat_t, contra_temp):
    """One fwd+bwd contrastive step at W=1 (what bench.py times on the host cores)."""
    feat_cond = feat_cond.detach().requires_grad_()
    feat_t = feat_t.detach().requires_grad_()
    temp = torch.tensor(float(contra_temp), requires_grad=True)
    loss, n1, n2 = itc_and_negatives(feat_cond, feat_t, feat_t.detach(), feat_cond.detach(), temp)
    loss.backward()
    return loss.detach(), feat_cond.grad, feat_t.grad, temp.grad, n1, n2


def retrieval_step(feat_t, feat_cond, k):
    """evaluation_mm.py:223 + :257 (+ the metric's sort, :333): dense fp32 scores, top-k, full sort."""
    score = torch.matmul(feat_t, feat_cond.permute(1, 0))
    idx = score.topk(k, dim=1)[1]
    return score, idx
