"""Developer diagnostic (GPU box): run the tcgen05 GEMM on tiny inputs and describe any mismatch."""
import sys
import torch
sys.path.insert(0, ".")
from vast_b200 import ops

torch.manual_seed(0)
for (m, n, k) in [(128, 256, 64), (128, 256, 128), (256, 512, 256), (200, 300, 136)]:
    a = (torch.randint(-4, 5, (m, k), device="cuda").float() / 8).bfloat16()
    b = (torch.randint(-4, 5, (n, k), device="cuda").float() / 8).bfloat16()
    c = ops.gemm_nt(a, b)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T
    bad = (c != ref)
    print(f"[{m}x{n}x{k}] mismatches {bad.sum().item()} / {bad.numel()}  maxerr {(c - ref).abs().max().item():.4g}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        print("  bad rows (first)", rows, " bad cols (first)", cols)
        print("  c[0,:8]  ", c[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        # does the result equal a GEMM over only part of K / permuted K?
        for kk in range(16, k + 1, 16):
            part = a[:, :kk].float() @ b[:, :kk].float().T
            if torch.equal(part, c):
                print("  == partial K", kk)
        break
print("diag done")
