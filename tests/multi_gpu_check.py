"""Multi-GPU correctness check (run under torchrun, one rank per GPU, NCCL):
  * omc_loss_and_negatives with the packed all-gather == the oracle evaluated on the gathered features
  * column-sharded retrieval_topk + candidate all-gather + merge == the single-GPU result
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import vast_b200
from oracle import spec

bs, d, temp = 192, 256, 0.07
n = bs * world
g = torch.Generator().manual_seed(7)
t = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, d, generator=g), dim=-1)
sl = slice(rank * bs, (rank + 1) * bs)
ft = t[sl].cuda().requires_grad_()
fc = c[sl].cuda().requires_grad_()
tau = torch.nn.Parameter(torch.tensor(temp, device="cuda"))
loss, neg_text, neg_cond = vast_b200.omc_loss_and_negatives(fc, ft, tau)      # rank / world from the process group
loss.backward()
tb, cb = t.bfloat16().float().numpy(), c.bfloat16().float().numpy()
o = spec.omc_loss(cb[sl], tb[sl], tb, cb, temp, rank=rank)
ok = abs(loss.item() - o["loss"]) < 1e-3 * abs(o["loss"])
for got, want in ((ft.grad, o["grad_t"]), (fc.grad, o["grad_cond"])):
    ok &= np.linalg.norm(got.cpu().numpy() - want) < 1e-3 * np.linalg.norm(want)
ok &= abs(tau.grad.item() - o["grad_temp"]) < 1e-3 * abs(o["grad_temp"])
tgt = torch.arange(rank * bs, (rank + 1) * bs)
ok &= bool((neg_text.cpu() != tgt).all() and (neg_cond.cpu() != tgt).all() and (neg_text.cpu() < n).all())

# fused pack + all-gather over peer memory == pack_pair + NCCL all-gather, bit for bit
from vast_b200 import ops
from vast_b200.peer import packed_gather
pg = packed_gather(bs, d, "cuda")
if pg is not None:
    for it in range(3):
        a = torch.randn(bs, d, generator=g).cuda() + rank
        b = torch.randn(bs, d, generator=g).cuda()
        got_pack = pg.gather(a, b).clone()
        ref_pack = torch.empty(n, 2 * d, dtype=torch.bfloat16, device="cuda")
        dist.all_gather_into_tensor(ref_pack, ops.pack_pair(a, b))
        ok &= bool(torch.equal(got_pack, ref_pack))
    if rank == 0:
        print("peer gather:", pg.mode, flush=True)
elif rank == 0:
    print("peer gather: unavailable", flush=True)

# the same step as one CUDA-graph launch (vast_b200.OmcGraphStep): same loss / gradients, three steps in a row
gstep = vast_b200.OmcGraphStep(bs, d, tau)
prev = None
for it in range(3):
    ft2 = t[sl].cuda().requires_grad_()
    fc2 = c[sl].cuda().requires_grad_()
    tau.grad = None
    l2, nt2, nc2 = gstep(fc2, ft2)
    l2.backward()
    ok &= abs(l2.item() - o["loss"]) < 1e-3 * abs(o["loss"])
    for got, want in ((ft2.grad, o["grad_t"]), (fc2.grad, o["grad_cond"])):
        ok &= np.linalg.norm(got.cpu().numpy() - want) < 1e-3 * np.linalg.norm(want)
    ok &= abs(tau.grad.item() - o["grad_temp"]) < 1e-3 * abs(o["grad_temp"])
    ok &= bool((nt2.cpu() != tgt).all() and (nc2.cpu() != tgt).all() and (nt2.cpu() < n).all())
    if prev is not None:
        ok &= not torch.equal(prev, nt2)      # fresh Philox noise on every replay
    prev = nt2.clone()
if rank == 0:
    print("graphed step ok:", bool(ok), flush=True)

# retrieval: sharded == single
nt, nv, k = 700, 3001, 16
q = torch.nn.functional.normalize(torch.randn(nt, d, generator=g), dim=-1).cuda()
v = torch.nn.functional.normalize(torch.randn(nv, d, generator=g), dim=-1).cuda()
for mode in ("bf16", "fp32"):
    v1, i1 = vast_b200.retrieval_topk(q, v, k, mode=mode)
    v2, i2 = vast_b200.retrieval_topk(q, v, k, mode=mode, shard=(rank, world), shard_mode="cols")
    v3, i3 = vast_b200.retrieval_topk(q, v, k, mode=mode, shard=(rank, world))          # default: query rows sharded
    ok &= bool(torch.equal(i1, i2)) and bool(torch.equal(i1, i3)) and bool(torch.equal(v1, v3))
# streaming re-rank (refine_candidates): every rank owns a slice of the videos' condition tokens; lists and ITM scores
# must equal the unsharded ones (ITM stub scores each pair independently of its mini-batch)
from vast_b200 import retrieval


class _Stub:
    def __init__(self):
        gg = torch.Generator().manual_seed(3)
        self.emb = (torch.randn(1000, 16, generator=gg) * 0.1).cuda()
        self.w = torch.randn(16, 2, generator=gg).cuda()

    def compute_slice_scores(self, cond, ids, mask):
        h = torch.tanh(self.emb[ids[:, 0]] + cond.float().mean(dim=1)[:, :16])
        return torch.softmax(h @ self.w, dim=1)[:, 1]


stub = _Stub()
nt2, nv2, k2 = 301, 8 * world + 3, 6
q2 = torch.nn.functional.normalize(torch.randn(nt2, d, generator=g), dim=-1).cuda()
v2f = torch.nn.functional.normalize(torch.randn(nv2, d, generator=g), dim=-1).cuda()
ids2 = torch.randint(0, 1000, (nt2, 5), generator=g).cuda()
mask2 = torch.ones(nt2, 5, dtype=torch.int64).cuda()
cond2 = torch.randn(nv2, 4, 16, generator=g).cuda()
per = -(-nv2 // world)
mine = slice(min(rank * per, nv2), min((rank + 1) * per, nv2))      # ragged: the last rank holds fewer videos
for direction in ("forward", "backward"):
    idx_s, itm_s = retrieval.refine_candidates(cond2[mine], ids2, mask2, q2, v2f, stub, k2, direction)
    if direction == "forward":
        _, idx_1 = vast_b200.retrieval_topk(q2, v2f, k2, mode="fp32")
        tt = torch.arange(nt2, device="cuda")[:, None].expand_as(idx_1).reshape(-1)
        vv = idx_1.reshape(-1).long()
    else:
        _, idx_1 = vast_b200.retrieval_topk(v2f, q2, k2, mode="fp32")
        vv = torch.arange(nv2, device="cuda")[:, None].expand_as(idx_1).reshape(-1)
        tt = idx_1.reshape(-1).long()
    want = stub.compute_slice_scores(cond2[vv], ids2[tt], mask2[tt]).view_as(idx_1)
    ok &= bool(torch.equal(idx_s, idx_1)) and bool(torch.allclose(itm_s, want, rtol=1e-5, atol=1e-7))
if rank == 0:
    print("streaming re-rank ok:", bool(ok), flush=True)

# id lists as sized tensor all-gathers == the reference's pickled gather, on NCCL
my_ids = [f"vid{rank}_{i}" for i in range(rank + 1)]
ok &= vast_b200.all_gather_ids(my_ids) == [j for i in vast_b200.all_gather_list(my_ids) for j in i]
ok &= vast_b200.all_gather_ids([rank * 10 + i for i in range(3)]) == [r * 10 + i for r in range(world) for i in range(3)]

# two graphed steps of the SAME shape used back to back with no host synchronisation in between (three sub-tasks per
# training step is the natural setup): each owns its pair of symmetric buffers, so a fast rank's second push can never
# land in a buffer a slower rank still reads.  Results must equal the single-step ones, every time.
gs_a = vast_b200.OmcGraphStep(bs, d, tau)
gs_b = vast_b200.OmcGraphStep(bs, d, tau)
assert gs_a.pg is None or gs_a.pg is not gs_b.pg
fin = []
for it in range(6):
    for gs in (gs_a, gs_b):
        la, _, _ = gs(c[sl].cuda(), t[sl].cuda())
        fin.append(la)                       # device tensors only: no .item() inside the loop
vals = torch.stack([x.detach().clone() for x in fin[-2:]] + [fin[0].detach().clone()]).cpu()
ok &= bool(((vals - o["loss"]).abs() < 1e-3 * abs(o["loss"])).all())
if rank == 0:
    print("two graphed steps, no host sync ok:", bool(ok), flush=True)

# negative-row exchange == all_gather_with_grad(x)[idx], values and gradients (NCCL all_to_all).
# Rank-dependent data and per-rank request lists: wrong-owner routing or swapped send / recv splits cannot pass.
gr = torch.Generator().manual_seed(100 + rank)
xg = (torch.randn(bs, 7, 16, generator=gr) + rank).cuda()
idx = torch.randint(0, n, (bs,), generator=gr).cuda()
wgt = torch.randn(bs, 7, 16, generator=gr).cuda()
xa = xg.clone().requires_grad_()
ref = vast_b200.all_gather_with_grad(xa)[idx]
(ref * wgt).sum().backward()
xb = xg.clone().requires_grad_()
got = vast_b200.exchange_rows(xb, idx)
(got * wgt).sum().backward()
ok &= bool(torch.equal(got, ref)) and bool(torch.allclose(xb.grad, xa.grad, atol=1e-5))
# the same exchange over peer memory, fused into the 3-way concat (what forward_ret uses): values and gradients equal
# the reference formulation all_gather_with_grad(cond)[neg] -> cat(cond, cond_neg, cond)
from vast_b200 import contrastive
S2, H2, L2 = 5, 16, 6
cond = (torch.randn(bs, S2, H2, generator=gr) + rank).cuda()
ids_l = torch.randint(0, 1000, (bs, L2), generator=gr).cuda()
msk_l = torch.ones(bs, L2, dtype=torch.int64).cuda()
ids_a, msk_a = vast_b200.concat_all_gather(ids_l), vast_b200.concat_all_gather(msk_l)
neg_t = torch.randint(0, n, (bs,), generator=gr).cuda()
neg_c = torch.randint(0, n, (bs,), generator=gr).cuda()
w3 = torch.randn(3 * bs, S2, H2, generator=gr).cuda()
for it in range(3):                                  # three rounds: both alternating buffers re-used
    ca = cond.clone().requires_grad_()
    ref3 = torch.cat((ca, vast_b200.all_gather_with_grad(ca)[neg_c], ca))
    (ref3 * w3).sum().backward()
    cb = cond.clone().requires_grad_()
    got = contrastive.gather_negatives_peer(cb, ids_l, msk_l, ids_a, msk_a, neg_t, neg_c)
    if got is None:
        if rank == 0:
            print("peer row exchange: unavailable", flush=True)
        break
    i1, a1, c3 = got
    (c3 * w3).sum().backward()
    ok &= bool(torch.equal(c3, ref3)) and bool(torch.allclose(cb.grad, ca.grad, atol=1e-5))
    ok &= bool(torch.equal(i1, torch.cat((ids_l, ids_l, ids_a[neg_t]))))
if rank == 0:
    print("peer row exchange ok:", bool(ok), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("nccl_check", "OK" if flag.item() == 1 else "FAILED", "world", world, "loss", loss.item(), flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
