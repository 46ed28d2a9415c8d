"""fp32 'exact' retrieval mode at scale: tensor-core shortlist (3-term bf16 split, K' = 6D) + fp64 re-score."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vast_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = torch.Generator().manual_seed(0)
t = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1).cuda()
v = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1).cuda()
for mode in ("bf16", "fp32"):
    for _ in range(2):
        vast_b200.retrieval_topk(t, v, 16, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        vast_b200.retrieval_topk(t, v, 16, mode=mode)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} mode={mode}: {e0.elapsed_time(e1)/3:.2f} ms per full top-16 retrieval", flush=True)
