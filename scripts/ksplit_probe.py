"""Cluster split-K (two CTA pairs per output tile, reduced through distributed shared memory) against the split-K +
reduce-kernel form and the library GEMM at small-M dQ shapes (plain-store epilogue)."""
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


for (m, n) in ((2048, 1024), (1024, 1024)):
    for k in (2048, 4096, 8192, 16384):
        a = torch.randn(m, k, device="cuda").half()
        bt = torch.randn(k, n, device="cuda").half()
        ref = torch.matmul(a, bt).float()
        row = {"M": m, "N": n, "K": k, "library": round(timeit(lambda: torch.matmul(a, bt)), 1)}
        os.environ.pop("VAST_GEMM_PROBE", None)
        row["splitk_plus_reduce"] = round(timeit(lambda: ops.gemm_nn(a, bt)), 1)
        os.environ["VAST_GEMM_PROBE"] = "7"
        try:
            out = ops.gemm_nn(a, bt)
            row["cluster_splitk"] = round(timeit(lambda: ops.gemm_nn(a, bt)), 1)
            row["max_rel_err"] = float(((out - ref).abs().max() / ref.abs().max()))
        except RuntimeError as e:
            row["cluster_splitk"] = str(e)[:80]
        os.environ.pop("VAST_GEMM_PROBE", None)
        print(row, flush=True)
