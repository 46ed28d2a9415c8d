"""Mainloop rate (plain-store epilogue) of the tcgen05 GEMM on the contrastive step's shapes vs cuBLAS:
NT (K-major B, the S GEMM) and NN (row-major B read MN-major, the dQ GEMM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (m, n, k) in [(8192, 4096, 1024), (4096, 4096, 1024), (8192, 1024, 4096), (8192, 8192, 8192)]:
    a = torch.randn(m, k, device="cuda").half()
    b = torch.randn(n, k, device="cuda").half()
    bt = b.t().contiguous()
    t_nt = timeit(lambda: ops.gemm_nt(a, b))
    t_nn = timeit(lambda: ops.gemm_nn(a, bt))
    t_ref = timeit(lambda: torch.matmul(a, b.T))
    t_ref2 = timeit(lambda: torch.matmul(a, bt))
    fl = 2.0 * m * n * k
    print(f"{m}x{n}x{k}: NT {t_nt*1e3:7.1f} us {fl/t_nt/1e9:7.1f} TF | NN {t_nn*1e3:7.1f} us {fl/t_nn/1e9:7.1f} TF | "
          f"cuBLAS NT {t_ref*1e3:7.1f} us {fl/t_ref/1e9:7.1f} TF, NN {t_ref2*1e3:7.1f} us {fl/t_ref2/1e9:7.1f} TF", flush=True)
