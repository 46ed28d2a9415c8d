"""Host-side cost of one e2e step (perf_counter around each segment; bs small so the GPU is never the bound)."""
import json
import sys
import time
import torch
import vast_b200

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dim = 1024
dev = torch.device("cuda", 0)
temp = torch.nn.Parameter(torch.tensor(0.07, device=dev))
step = vast_b200.OmcGraphStep(bs, dim, temp, rank=0, world_size=1, dtype=torch.bfloat16)
pin = [torch.randn(bs, dim).bfloat16().pin_memory() for _ in range(2)]
dbuf = [torch.empty(bs, dim, dtype=torch.bfloat16, device=dev) for _ in range(2)]
copy_stream = torch.cuda.Stream()
ev = torch.cuda.Event()
seg = {}


def mark(name, t0):
    t = time.perf_counter()
    seg[name] = seg.get(name, 0.0) + (t - t0)
    return t


def one(call, record):
    t = time.perf_counter()
    torch.cuda.current_stream().wait_event(ev)
    ft = dbuf[0].detach().requires_grad_()
    fc = dbuf[1].detach().requires_grad_()
    temp.grad = None
    if record: t = mark("prep", t)
    loss, a, b = call(fc, ft)
    if record: t = mark("call", t)
    loss.backward()
    if record: t = mark("backward", t)
    with torch.cuda.stream(copy_stream):
        dbuf[0].copy_(pin[0], non_blocking=True)
        dbuf[1].copy_(pin[1], non_blocking=True)
        ev.record(copy_stream)
    if record: t = mark("prefetch", t)
    v = loss.item()
    if record: t = mark("item", t)


res = {}
for name, call in (("graph", step), ("eager", lambda fc, ft: vast_b200.omc_loss_and_negatives(fc, ft, temp, rank=0, world_size=1))):
    ev.record()
    for _ in range(20):
        one(call, False)
    seg.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 300
    for _ in range(n):
        one(call, True)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / n * 1e6
    res[name] = {"total_us": round(tot, 1), **{k: round(v / n * 1e6, 1) for k, v in seg.items()}}
# raw replay + sync latency
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step._graphs[0].replay()
    torch.cuda.synchronize()
res["replay_plus_sync_us"] = round((time.perf_counter() - t0) / 300 * 1e6, 1)
print(json.dumps(res))
