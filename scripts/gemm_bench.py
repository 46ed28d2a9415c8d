"""Mainloop throughput check: vast_gemm_nt (plain-store epilogue) vs torch.matmul (cuBLAS) on the same shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops


def timeit(fn, iters=10):
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


for (m, n, k) in [(8192, 4096, 1024), (8192, 8192, 1024), (8192, 8192, 8192), (8192, 1024, 4096), (16384, 16384, 512)]:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    t_ours = timeit(lambda: ops.gemm_nt(a, b))
    t_ref = timeit(lambda: torch.matmul(a, b.T))
    fl = 2.0 * m * n * k
    print(f"{m}x{n}x{k}: ours {t_ours*1e3:8.1f} us {fl/t_ours/1e9:7.1f} TF | cuBLAS {t_ref*1e3:8.1f} us {fl/t_ref/1e9:7.1f} TF", flush=True)
