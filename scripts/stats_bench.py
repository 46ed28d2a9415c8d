"""Light-epilogue mainloop rate at small K: the two-pass OMC statistics GEMM (row max / sum-exp only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops
for n, d in [(16384, 512), (8192, 512), (4096, 1024), (16384, 256)]:
    t = torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, d, device="cuda"), dim=-1)
    pack = ops.pack_pair(t, c)
    buf = None
    for _ in range(3):
        buf = ops.omc_step(pack, n, 0, 0.07, need_sample=True, need_grad=False, two_pass=True, buffers=buf)
    ops.kernel_timing(True)
    for _ in range(5):
        buf = ops.omc_step(pack, n, 0, 0.07, need_sample=True, need_grad=False, two_pass=True, buffers=buf)
    torch.cuda.synchronize()
    recs = ops.kernel_timing_read()
    ops.kernel_timing(False)
    agg = {}
    for nm, ms in recs:
        agg.setdefault(nm, []).append(ms)
    fl = 4.0 * n * n * d
    for nm in ("omc_stats_gemm", "omc_soft_gemm"):
        us = 1e3 * sum(agg[nm]) / len(agg[nm])
        print(f"n={n} d={d} {nm}: {us:8.1f} us  {fl / us / 1e6:7.1f} TF", flush=True)
