import os, sys, torch
sys.path.insert(0, "/root/repo")
from vast_b200 import ops
g = torch.Generator().manual_seed(3)
n, d = 100000, 512
t = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda().bfloat16()
v = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda().bfloat16()
for _ in range(2):
    out = ops.sim_topk(t, v, 16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = ops.sim_topk(t, v, 16)
e1.record()
torch.cuda.synchronize()
print(os.environ.get("VAST_TOPK_RING4"), "ms", e0.elapsed_time(e1) / 5, "chk", int(out[1].sum()) if isinstance(out, tuple) else None)
