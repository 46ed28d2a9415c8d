"""Fused projection (vast_project_normalize) vs cuBLAS Linear + vast_l2norm at the pretraining shapes."""
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
    return e0.elapsed_time(e1) / iters * 1e3


for bs in (512, 1024, 4096):
    x = torch.randn(bs, 2944, device="cuda").bfloat16()
    lin = torch.nn.Linear(2944, 1024).cuda().bfloat16()
    slot = torch.empty(bs, 2048, dtype=torch.bfloat16, device="cuda")
    w_op = ops._pack(lin.weight.detach(), ops.SIM_BF16, False)
    with torch.no_grad():
        fused = timeit(lambda: ops.project_normalize(x, lin.weight, lin.bias, out16=slot[:, :1024], w_op=w_op))
        unf = timeit(lambda: ops.l2norm(lin(x).float(), out16=slot[:, :1024]))
        ops.kernel_timing(True)
        for _ in range(10):
            ops.gemm_nt(x, w_op)
            ops.project_normalize(x, lin.weight, lin.bias, out16=slot[:, :1024], w_op=w_op)
            ops.l2norm(lin(x).float(), out16=slot[:, :1024])
        torch.cuda.synchronize()
        recs = ops.kernel_timing_read()
        ops.kernel_timing(False)
        dev = {}
        for k, v in recs:
            dev.setdefault(k, []).append(v * 1e3)
        print("   device (event bracket ~3 us included):", {k: round(sorted(v)[len(v) // 2], 1) for k, v in dev.items()})
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        y = lin(x)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            y = lin(x)
        e1.record()
        torch.cuda.synchronize()
        print("   cuBLAS Linear alone, back to back:", round(e0.elapsed_time(e1) / 20 * 1e3, 1), "us")
    print(f"bs {bs}: fused {fused:.1f} us ({2 * bs * 2944 * 1024 / fused / 1e6:.0f} TFLOP/s) | cuBLAS + cast + l2norm {unf:.1f} us", flush=True)
