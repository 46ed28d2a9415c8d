"""What ncu sees of the library's bf16 GEMM (torch.matmul, 8192^3) next to this repo's contrastive step at cfg3: cluster
size, grid, L2 sector traffic, tensor-pipe activity.  Run under ncu with a --metrics list (see scripts/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops

a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    c = a @ b
torch.cuda.synchronize()
# dQ-shaped library GEMM: [4096, 4096] x [4096, 1024]
p = torch.randn(4096, 4096, device="cuda", dtype=torch.float16)
k = torch.randn(4096, 1024, device="cuda", dtype=torch.float16)
for _ in range(2):
    d = p @ k
torch.cuda.synchronize()
g = torch.Generator().manual_seed(1)
t = torch.nn.functional.normalize(torch.randn(4096, 1024, generator=g), dim=-1)
cnd = torch.nn.functional.normalize(t + 0.8 * torch.randn(4096, 1024, generator=g), dim=-1)
temp = torch.full((1,), 0.07, device="cuda")
buf = None
for _ in range(2):
    buf = ops.omc_step_local(t.cuda(), cnd.cuda(), temp, seed=1, offset=0, buffers=buf)
torch.cuda.synchronize()
print("loss", buf["loss"].item())
