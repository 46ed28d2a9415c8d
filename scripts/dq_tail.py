"""Fixed cost + exposed epilogue of the fused dQ kernel: one 256 x 256 tile per CTA pair (bs = 1024 rows, VAST_OMC_DQ=256,2,1),
time against K = n_total.  The intercept minus the plain-store kernel's (scripts/gemm_slope.py: ~6 us) is what the
gradient-assembly epilogue of the LAST tile costs when nothing hides it."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VAST_OMC_DQ"] = "256,2,1"
import torch
from vast_b200 import ops

D, bs = 1024, 1024
temp = torch.full((1,), 0.07, device="cuda")
for n in (1024, 2048, 4096, 8192, 16384):
    g = torch.Generator().manual_seed(1)
    t = torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, D, generator=g), dim=-1)
    pack = ops.pack_pair(t.cuda(), c.cuda())
    buf = None
    for _ in range(3):
        buf = ops.omc_step(pack, bs, 0, temp, seed=1, offset=0, buffers=buf)
    torch.cuda.synchronize()
    ops.kernel_timing(True)
    for _ in range(20):
        buf = ops.omc_step(pack, bs, 0, temp, seed=1, offset=0, buffers=buf)
    torch.cuda.synchronize()
    recs = ops.kernel_timing_read()
    ops.kernel_timing(False)
    agg = {}
    for nm, ms in recs:
        agg.setdefault(nm, []).append(ms * 1e3)
    print(json.dumps({"n_total": n, "k_blocks": n // 64, **{k: round(sorted(v)[len(v) // 2], 2) for k, v in agg.items()}}), flush=True)
