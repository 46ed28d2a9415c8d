"""Per-rank cost of column-sharded cfg5 retrieval on W GPUs, emulated on one: cold shard scan vs the two-phase form
(phase A: a 1/W slice of the rows against the rank's own columns -> per-row bounds; phase B: all rows, warm lists).
    python scripts/cols_bench.py [W] [mode]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
n, d, k = 100_000, 512, 16
sim = ops.SIM_BF16 if mode == "bf16" else ops.SIM_FP32X2
kl = k if mode == "bf16" else 32
g = torch.Generator().manual_seed(4321)
t = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda()
v = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda()
q = ops.sim_pack_operand(t, sim, True)
per = n // W
kop = ops.sim_pack_operand(v[:per], sim, False)
kall = ops.sim_pack_operand(v, sim, False)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# bounds as every rank would contribute them: rank r's row slice against rank r's columns
bounds = torch.cat([ops.sim_topk(q[r * per:(r + 1) * per], ops.sim_pack_operand(v[r * per:(r + 1) * per], sim, False), kl,
                                 col_offset=r * per, want_bounds=True)[1] for r in range(W)])
out = {
    "W": W, "mode": mode,
    "single_gpu_ms": timeit(lambda: ops.sim_topk(q, kall, kl)),
    "rows_shard_ms": timeit(lambda: ops.sim_topk(q[:per], kall, kl)),
    "cols_cold_ms": timeit(lambda: ops.sim_topk(q, kop, kl)),
    "cols_phaseA_ms": timeit(lambda: ops.sim_topk(q[:per], kop, kl, want_bounds=True)),
    "cols_phaseB_warm_ms": timeit(lambda: ops.sim_topk(q, kop, kl, bounds_in=bounds)),
}
warm = ops.sim_topk(q, kop, kl, bounds_in=bounds)
out["warm_list_occupancy"] = float((warm != 0).float().mean())
print(json.dumps(out))
