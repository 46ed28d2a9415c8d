"""Where the tcgen05 mainloop's time goes (plain-store epilogue): the whole GEMM, the tensor pipe alone (no operand
loads), the operand ingest alone (no MMAs), each with lone CTA pairs and with clusters of two pairs that share A through
TMA multicast -- next to the library GEMM.  VAST_GEMM_PROBE selects the mode inside vast_gemm_nt / vast_gemm_nn.
    python scripts/gemm_probe.py > gpurun_out/gemm_probe.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops

MODES = {0: "pairs", 1: "pairs, MMA only", 2: "pairs, loads only", 3: "quads (A shared)", 4: "quads, MMA only", 5: "quads, loads only"}


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


for (m, n, k) in [(8192, 8192, 8192), (4096, 1024, 4096), (4096, 4096, 1024)]:
    a = torch.randn(m, k, device="cuda").half()
    b = torch.randn(n, k, device="cuda").half()
    bt = b.t().contiguous()
    fl = 2.0 * m * n * k
    kblocks = (m // 256) * (n // 256) * (k // 64)
    t_ref = timeit(lambda: torch.matmul(a, b.T))
    t_ref2 = timeit(lambda: torch.matmul(a, bt))
    print(f"{m}x{n}x{k}: library NT {t_ref*1e3:8.1f} us {fl/t_ref/1e9:7.1f} TF | NN {t_ref2*1e3:8.1f} us {fl/t_ref2/1e9:7.1f} TF", flush=True)
    for mode, label in MODES.items():
        os.environ["VAST_GEMM_PROBE"] = str(mode)
        t_nt = timeit(lambda: ops.gemm_nt(a, b))
        t_nn = timeit(lambda: ops.gemm_nn(a, bt))
        print(f"   {label:22s} NT {t_nt*1e3:8.1f} us {fl/t_nt/1e9:7.1f} TF | NN {t_nn*1e3:8.1f} us {fl/t_nn/1e9:7.1f} TF"
              f" | us per 256x256x64 k-block per pair (74 / 64 pairs): NT {t_nt*1e3/kblocks*(64 if mode>=3 else 74):.3f} NN {t_nn*1e3/kblocks*(64 if mode>=3 else 74):.3f}", flush=True)
    os.environ.pop("VAST_GEMM_PROBE", None)
