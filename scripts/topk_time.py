"""Device time of one cfg5 `sim_topk` call (A/B of VAST_TOPK_RING4 / VAST_TOPK_ARES / VAST_TOPK_DEBUG), cold and with
the lists started from the final bounds of a previous call (filters run, almost nothing reaches the list warps)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vast_b200 import ops


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator().manual_seed(3)
n, d = 100000, 512
t = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda().bfloat16()
v = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda().bfloat16()
print("env", {k: os.environ[k] for k in os.environ if k.startswith("VAST_TOPK")})
print("cold   ms", round(timeit(lambda: ops.sim_topk(t, v, 16)), 3))
if not os.environ.get("VAST_TOPK_DEBUG"):
    keys, bounds = ops.sim_topk(t, v, 16, want_bounds=True)
    print("warm   ms", round(timeit(lambda: ops.sim_topk(t, v, 16, bounds_in=bounds)), 3))
