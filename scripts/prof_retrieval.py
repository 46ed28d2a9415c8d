"""Small driver for profiling the streaming top-k GEMM: python scripts/prof_retrieval.py [n] [d] [k] [iters]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vast_b200
from vast_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nk = int(os.environ.get("NK", n))
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = torch.Generator().manual_seed(0)
t = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda()
v = torch.nn.functional.normalize(torch.randn(nk, d, generator=g), dim=-1).cuda()
q = ops.sim_pack_operand(t, ops.SIM_BF16, True)
kk = ops.sim_pack_operand(v, ops.SIM_BF16, False)
for _ in range(2):
    ops.sim_topk(q, kk, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    keys = ops.sim_topk(q, kk, k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
ops.kernel_timing(True)
ops.sim_topk(q, kk, k)
torch.cuda.synchronize()
kt = [t for nm, t in ops.kernel_timing_read() if nm == "sim_topk_gemm"]
ops.kernel_timing(False)
print(f"n={n} nk={nk} d={d} k={k}: {ms:.3f} ms/iter, {2.0 * n * nk * d / ms / 1e9:.1f} TFLOP/s; gemm kernel {kt[0]:.3f} ms")
