"""Fixed cost vs per-k-block cost of the plain-store GEMM at the dQ shape (M = 4096, N = 1024, one tile per CTA pair):
time against K for the whole kernel, its MMA-only and loads-only probes, and the library GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


m, n = 4096, 1024
for k in (512, 1024, 2048, 4096, 8192, 16384):
    a = torch.randn(m, k, device="cuda").half()
    bt = torch.randn(k, n, device="cuda").half()
    row = {"K": k, "library": round(timeit(lambda: torch.matmul(a, bt)), 1)}
    for mode, label in ((0, "pairs"), (1, "pairs_mma_only"), (2, "pairs_loads_only"), (3, "quads"), (4, "quads_mma_only")):
        os.environ["VAST_GEMM_PROBE"] = str(mode)
        row[label] = round(timeit(lambda: ops.gemm_nn(a, bt)), 1)
    os.environ.pop("VAST_GEMM_PROBE", None)
    print(row, flush=True)
