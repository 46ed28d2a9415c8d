"""Probe: contrastive step time with and without CUDA-graph replay (cfg3, W=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops
import bench
N, D = 4096, 1024
R = 5
sets = [tuple(x.cuda() for x in bench.synth(N, D, 1234 + s)) for s in range(R)]
temp = torch.full((1,), 0.07, device="cuda")
pack = torch.empty(N, 2 * D, dtype=torch.bfloat16, device="cuda")
state = {"buf": None}

def step(i):
    ft, fc = sets[i % R]
    ops.pack_pair(ft, fc, out=pack)
    state["buf"] = ops.omc_step(pack, N, 0, temp, 0.1, 1e-4, seed=1234, offset=i, need_sample=True, need_grad=True, buffers=state["buf"])

def timeit(fn, K=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3

print("eager  us/step", timeit(step))
s = torch.cuda.Stream()
graphs = []
with torch.cuda.stream(s):
    for i in range(3): step(i)
    torch.cuda.synchronize()
    for i in range(R):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step(i)
        graphs.append(g)
torch.cuda.synchronize()
print("graphs us/step", timeit(lambda i: graphs[i % R].replay()))
print("loss", state["buf"]["loss"].item())
