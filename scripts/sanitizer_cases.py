"""Small invocations of every kernel family of the library, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool <tool> python scripts/sanitizer_cases.py
Shapes are tiny (the tools slow kernels down 10-100x) but hit every code path: CTA pairs and lone CTAs, split rows and
columns, ragged edges, the warp-specialised top-k epilogue with its shared-memory queues, the ticketed reductions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vast_b200
from vast_b200 import ops

torch.manual_seed(0)
dev = "cuda"
nrm = torch.nn.functional.normalize


def feats(n, d):
    t = nrm(torch.randn(n, d, device=dev), dim=-1)
    return t, nrm(t + 0.8 * torch.randn(n, d, device=dev), dim=-1)


# contrastive step: gathered form (ranks emulated), single-rank symmetric form, two-pass, loss only, split-K shape
t, c = feats(512, 128)
ops.omc_step(ops.pack_pair(t, c), 256, 256, 0.07, seed=1, offset=2)
ops.omc_step(ops.pack_pair(t, c), 130, 100, 0.07, seed=1, offset=2, two_pass=True)
ops.omc_step_local(t, c, 0.07, seed=1, offset=2)
ops.omc_step_local(t[:300], c[:300], 0.07, seed=1, offset=2, need_grad=False, need_sample=False)
t2, c2 = feats(1024, 64)
ops.omc_step(ops.pack_pair(t2, c2), 128, 0, 0.07, seed=3, offset=0)
ops.omc_step(ops.pack_pair(t2, c2), 128, 0, 0.07, seed=3, offset=0, debug_noise=torch.empty(2, 128, 1024, device=dev).exponential_())
# retrieval: bf16 and exact top-k (pairs + tail splits), warm bounds, streaming rank, merges, dense drop-ins
q, v = feats(700, 64)[0], feats(3001, 64)[0]
vast_b200.retrieval_topk(q, v, 16, mode="bf16")
vast_b200.retrieval_topk(q, v, 10, mode="fp32")
vast_b200.retrieval_topk(q[:5], v[:40], 50, mode="bf16")
qo, ko = ops.sim_pack_operand(q, ops.SIM_BF16, True), ops.sim_pack_operand(v, ops.SIM_BF16, False)
keys, b = ops.sim_topk(qo, ko, 16, want_bounds=True)
ops.sim_topk(qo, ko, 16, bounds_in=b)
vast_b200.rank_of_gt(q, v, torch.arange(700, device=dev) % 3001)
s = torch.randn(100, 257, device=dev)
ops.dense_topk(s, 20, axis=1); ops.dense_topk(s, 20, axis=0)
ops.dense_rank_of_gt(s, torch.arange(100, device=dev), torch.arange(100, device=dev), axis=1)
off, texts = ops.bucket_by_video(torch.arange(300, device=dev).int(), (torch.arange(300, device=dev) % 7).int(), 7)
# feature build + heads
vis, aud, sub = torch.randn(9, 2, 5, 48, device=dev), torch.randn(9, 2, 6, 24, device=dev), torch.randn(9, 4, 24, device=dev)
lin = torch.nn.Linear(96, 40).to(dev)
f = vast_b200.build_feature(lin, vis.requires_grad_(), aud, sub)
f.sum().backward()
ops.project_normalize(torch.randn(300, 768, device=dev).bfloat16(), torch.randn(520, 768, device=dev).bfloat16())
ops.l2norm(torch.randn(33, 72, device=dev))
ops.gemm_nt(torch.randn(300, 200, device=dev).bfloat16(), torch.randn(129, 200, device=dev).bfloat16())
ops.gemm_nn(torch.randn(300, 200, device=dev).half(), torch.randn(200, 136, device=dev).half())


class Head(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.linear1, self.layernorm, self.linear2 = torch.nn.Linear(96, 96), torch.nn.LayerNorm(96, eps=1e-12), torch.nn.Linear(96, 2)


vast_b200.match_head_scores(Head().to(dev), torch.randn(300, 96, device=dev))
# gathers (local and "peer" with the ranks emulated)
blocks = [torch.randn(16, 3, 8, device=dev) for _ in range(2)]
ids = torch.randint(0, 99, (32, 4), device=dev)
neg = torch.randint(0, 32, (16,), device=dev)
ops.gather_rows_concat3(ids[:16], ids[:16], ids, ids, blocks[0], torch.cat(blocks), neg, neg)
ops.gather_rows_concat3_peer(ids[:16], ids[:16], ids, ids, blocks[0], [x.data_ptr() for x in blocks], 16, neg, neg)
ops.pull_row_grads(torch.cat([neg, neg]), [x.data_ptr() for x in blocks], 16, 0, blocks[0], base_grad=blocks[1])
torch.cuda.synchronize()
print("sanitizer cases done")
