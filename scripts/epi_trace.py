"""Where the fused dQ kernel's time goes per CTA (developer build: VAST_NVCC_EXTRA=-DVAST_EPI_TRACE python -m vast_b200.build --force):
%globaltimer stamps of the epilogue warps -- roles start, each tile's accumulator ready, each tile's epilogue done, finish()."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vast_b200 import ops
from vast_b200._lib import lib

n, d = 4096, 1024
g = torch.Generator().manual_seed(1)
t = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).cuda()
c = torch.nn.functional.normalize(t.cpu() + 0.8 * torch.randn(n, d, generator=g), dim=-1).cuda()
temp = torch.full((1,), 0.07, device="cuda")
buf = None
for _ in range(5):
    buf = ops.omc_step_local(t, c, temp, seed=1, offset=0, buffers=buf)
torch.cuda.synchronize()
fn = lib().vast_debug_epi_trace
fn.restype = ctypes.c_int
ctas = 128
arr = np.zeros((ctas, 16), dtype=np.uint64)
assert fn(arr.ctypes.data_as(ctypes.c_void_p), ctas, 0) == 0
a = arr.astype(np.int64)
t0 = a[:, 0].min()
rel = (a - t0) / 1e3   # us


def stat(x):
    return f"min {x.min():6.2f}  med {np.median(x):6.2f}  max {x.max():6.2f}"


print("roles start          ", stat(rel[:, 0]))
print("tile 1 acc ready     ", stat(rel[:, 1]))
print("tile 1 epilogue done ", stat(rel[:, 2]), "  duration", stat(rel[:, 2] - rel[:, 1]))
print("tile 2 acc ready     ", stat(rel[:, 3]), "  since tile 1 ready", stat(rel[:, 3] - rel[:, 1]))
print("tile 2 epilogue done ", stat(rel[:, 4]), "  duration", stat(rel[:, 4] - rel[:, 3]))
print("after finish()       ", stat(rel[:, 15]), "  since tile 2 epilogue", stat(rel[:, 15] - rel[:, 4]))
for ch in range(4):
    print(f"tile 2 chunk {ch}: TMEM values in registers", stat(rel[:, 5 + 2 * ch] - rel[:, 3]), "  functor done", stat(rel[:, 6 + 2 * ch] - rel[:, 3]))

# ---- the symmetric S GEMM (tag 1): 4 tiles per CTA pair
arr = np.zeros((ctas, 16), dtype=np.uint64)
assert fn(arr.ctypes.data_as(ctypes.c_void_p), ctas, 1) == 0
a = arr.astype(np.int64)
rel = (a - a[:, 0].min()) / 1e3
print("--- symmetric S GEMM")
print("roles start          ", stat(rel[:, 0]))
for tl in range(4):
    print(f"tile {tl + 1} acc ready     ", stat(rel[:, 1 + 2 * tl]), "  epilogue done", stat(rel[:, 2 + 2 * tl]), "  duration", stat(rel[:, 2 + 2 * tl] - rel[:, 1 + 2 * tl]))
print("after the last item  ", stat(rel[:, 15]))
