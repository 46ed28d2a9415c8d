"""Host->device copy bandwidth from pinned memory at the e2e step's transfer sizes (the bound of bench.py's e2e at N=1)."""
import json
import torch

dev = torch.device("cuda", 0)
out = {}
for name, nbytes, parts in (("2x8MiB", 8 << 20, 2), ("1x16MiB", 16 << 20, 1), ("1x256MiB", 256 << 20, 1)):
    src = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(parts)]
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(parts)]
    for _ in range(3):
        for s, d in zip(src, dst):
            d.copy_(s, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        for s, d in zip(src, dst):
            d.copy_(s, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name] = {"us": round(ms * 1e3, 1), "GBps": round(nbytes * parts / (ms * 1e-3) / 1e9, 1)}
print(json.dumps(out))
