import os, sys, torch
sys.path.insert(0, "/root/repo")
from vast_b200 import ops
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
g = torch.Generator().manual_seed(3)
n, d = 100000, 512
t32 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda()
v32 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda()
t, v = t32.bfloat16(), v32.bfloat16()
for W in (1, 2, 4, 8):
    rows = (n + W - 1) // W
    ts = t[:rows].contiguous()
    ms = timeit(lambda: ops.sim_topk(ts, v, 16))
    vs = v[:rows].contiguous()
    msc = timeit(lambda: ops.sim_topk(t, vs, 16))
    print(f"W={W}: rows shard {rows}x{n}: {ms:.3f} ms (ideal {8.34/W:.3f}); cols shard {n}x{rows} cold: {msc:.3f} ms", flush=True)
print("cast fp32->bf16 100k x 512:", timeit(lambda: v32.bfloat16()))
