"""Achieved HBM bandwidth of the memory-bound kernels of the path at the cfg3 per-rank shapes:
algorithmic bytes (compulsory reads + writes) / CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
Inputs are larger than the 126 MB L2 or rotated so that no call finds its inputs cached.
    python scripts/hbm_bench.py > profiles/<tag>_hbm_kernels.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        flush.zero_()                       # evict L2 between timed calls; the launches below queue up behind it,
        flush.zero_()                       # so host launch latency is not inside the timed region
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


out = []
bs, S, H, L, N = 512, 583, 768, 70, 4096          # cfg3 rank at W = 8
# a1-a4: pool + concat (vision cls over n frames, audio token mean, subtitle cls), bf16
vis = torch.randn(bs, 1, 257, 1408, device=dev).bfloat16()
aud = torch.randn(bs, 1, 256, 768, device=dev).bfloat16()
sub = torch.randn(bs, 70, 768, device=dev).bfloat16()
by = 2 * bs * (1 * 1408 + 1 * 256 * 768 + 768 + 2944)
t = timeit(lambda: ops.pool_concat(vis, aud, sub))
out.append(dict(kernel="pool_concat", shape="bs 512: vision [1,257,1408] cls + audio [1,256,768] mean + subtitle cls, bf16", bytes=by, us=t * 1e6))
# l2norm of [4096, 1024] fp32 (whole global batch on one rank) -> f32 + bf16 slot
x = torch.randn(4096, 1024, device=dev)
slot = torch.empty(4096, 2048, dtype=torch.bfloat16, device=dev)
by = 4096 * 1024 * (4 + 4 + 2)
t = timeit(lambda: ops.l2norm(x, out16=slot[:, :1024]))
out.append(dict(kernel="l2norm", shape="[4096,1024] f32 -> f32 + bf16 all-gather slot", bytes=by, us=t * 1e6))
# pack_pair [4096,1024] fp32 x2 -> bf16 [4096, 2048]
a, b = torch.randn(4096, 1024, device=dev), torch.randn(4096, 1024, device=dev)
by = 4096 * 1024 * (4 + 4 + 2 + 2)
t = timeit(lambda: ops.pack_pair(a, b, out=slot))
out.append(dict(kernel="pack_pair", shape="2 x [4096,1024] f32 -> [4096,2048] bf16", bytes=by, us=t * 1e6))
# the same two at 4x the rows (launch / ramp overheads amortised: the asymptotic rate of the kernel)
x4 = torch.randn(16384, 1024, device=dev)
slot4 = torch.empty(16384, 2048, dtype=torch.bfloat16, device=dev)
t = timeit(lambda: ops.l2norm(x4, out16=slot4[:, :1024]))
out.append(dict(kernel="l2norm (x4 rows)", shape="[16384,1024] f32 -> f32 + bf16 slot", bytes=16384 * 1024 * 10, us=t * 1e6))
a4, b4 = torch.randn(16384, 1024, device=dev), torch.randn(16384, 1024, device=dev)
t = timeit(lambda: ops.pack_pair(a4, b4, out=slot4))
out.append(dict(kernel="pack_pair (x4 rows)", shape="2 x [16384,1024] f32 -> [16384,2048] bf16", bytes=16384 * 1024 * 12, us=t * 1e6))
del x4, slot4, a4, b4
# a10: negative gather + 3-way concat, cfg3 rank: cond [bs,S,H] bf16 local, [N,S,H] gathered
cond_all = torch.randn(N, S, H, device=dev).bfloat16()
cond_loc = cond_all[:bs].clone()
ids_all = torch.randint(0, 30522, (N, L), device=dev)
mask_all = torch.ones(N, L, dtype=torch.int64, device=dev)
neg_t = torch.randint(0, N, (bs,), device=dev)
neg_c = torch.randint(0, N, (bs,), device=dev)
by = (2 + 3) * bs * S * H * 2 + (2 + 3) * bs * L * 8 * 2
t = timeit(lambda: ops.gather_rows_concat3(ids_all[:bs], mask_all[:bs], ids_all, mask_all, cond_loc, cond_all, neg_t, neg_c))
out.append(dict(kernel="gather_rows_concat3", shape="bs 512, S 583, H 768 bf16, L 70 (2 reads + 3 writes of a [bs,S,H] block)", bytes=by, us=t * 1e6))
for o in out:
    o["achieved_gbs"] = round(o["bytes"] / (o["us"] * 1e-6) / 1e9, 1)
    o["peak_gbs"] = peak
    o["frac"] = round(o["achieved_gbs"] / peak, 3)
    o["us"] = round(o["us"], 2)
print(json.dumps(out, indent=1))
