"""Fused projection (vast_project_normalize: Linear + bias + L2 normalise + bf16 slot, one tcgen05 kernel) vs the
unfused chain (cuBLAS Linear + cast + vast_l2norm) at the pretraining shapes.  Both are captured in a CUDA graph (20
calls) and replayed, so the numbers are DEVICE time per call, not Python / launch latency."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops


def graph_time(fn, reps=20, replays=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e3


def main():
  for bs in (512, 1024, 4096):
      x = torch.randn(bs, 2944, device="cuda").bfloat16()
      lin = torch.nn.Linear(2944, 1024).cuda().bfloat16()
      slot = torch.empty(bs, 2048, dtype=torch.bfloat16, device="cuda")
      w_op = ops._pack(lin.weight.detach(), ops.SIM_BF16, False)
      with torch.no_grad():
          fused = graph_time(lambda: ops.project_normalize(x, lin.weight, lin.bias, out16=slot[:, :1024], w_op=w_op))
          unf = graph_time(lambda: ops.l2norm(lin(x).float(), out16=slot[:, :1024]))
          gemm = graph_time(lambda: lin(x))
      print(f"bs {bs}: fused {fused:.1f} us ({2 * bs * 2944 * 1024 / fused / 1e6:.0f} TFLOP/s) | cuBLAS Linear + cast + l2norm {unf:.1f} us "
            f"(Linear alone {gemm:.1f} us)", flush=True)


if __name__ == "__main__":
    main()
