import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops
a = torch.randn(9472, 512, device="cuda").bfloat16()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = torch.randn(n, 512, device="cuda").bfloat16()
q = ops.sim_pack_operand(a.float(), ops.SIM_BF16, True)
kk = ops.sim_pack_operand(b.float(), ops.SIM_BF16, False)
for k in (16,):
    for _ in range(3): ops.sim_topk(q, kk, k)
    ops.kernel_timing(True)
    for _ in range(3): ops.sim_topk(q, kk, k)
    torch.cuda.synchronize()
    print("debug", os.environ.get("VAST_TOPK_DEBUG", "0"), "n", n, "k", k, [f"{t*1e3:.1f}us" for nm, t in ops.kernel_timing_read() if nm == "sim_topk_gemm"])
    ops.kernel_timing(False)
