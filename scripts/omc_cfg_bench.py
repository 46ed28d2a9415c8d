"""Per-kernel times of the fused contrastive step for dQ-GEMM tile configurations (VAST_OMC_DQ="bn,cl,max_ks") at the
per-rank shapes of BASELINE cfg3 on 1 / 2 / 4 / 8 GPUs (ranks emulated on one GPU through row_offset / n_total).
    python scripts/omc_cfg_bench.py [bs ...] > gpurun_out/omc_cfg.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vast_b200 import ops

N, D = 4096, 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator().manual_seed(1)
t = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
c = torch.nn.functional.normalize(t + 0.8 * torch.randn(N, D, generator=g), dim=-1)
pack = ops.pack_pair(t.cuda(), c.cuda())
temp = torch.full((1,), 0.07, device="cuda")
cfgs = sys.argv[2].split(";") if len(sys.argv) > 2 else ["", "256,2,1", "128,2,1", "256,1,1", "128,1,1", "128,1,2"]
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 2048, 1024, 512]
out = []
for bs in sizes:
    ref = None
    for cfg in cfgs:
        if cfg:
            os.environ["VAST_OMC_DQ"] = cfg
        else:
            os.environ.pop("VAST_OMC_DQ", None)
        buf = None
        try:
            for _ in range(3):
                buf = ops.omc_step(pack, bs, N - bs, temp, seed=1, offset=0, buffers=buf)
            torch.cuda.synchronize()
        except RuntimeError as e:
            out.append(dict(bs=bs, cfg=cfg, error=str(e)[:200]))
            continue
        if ref is None:
            ref = {k: buf[k].clone() for k in ("loss", "grad_t", "grad_cond", "neg_idx")}
        same = all(torch.equal(ref[k], buf[k]) for k in ("grad_t", "grad_cond", "neg_idx"))
        close = float((ref["grad_t"] - buf["grad_t"]).norm() / ref["grad_t"].norm())
        ops.kernel_timing(True)
        iters = 10
        for _ in range(iters):
            flush.zero_()
            buf = ops.omc_step(pack, bs, N - bs, temp, seed=1, offset=0, buffers=buf)
        torch.cuda.synchronize()
        recs = ops.kernel_timing_read()
        ops.kernel_timing(False)
        agg = {}
        for nm, ms in recs:
            agg.setdefault(nm, []).append(ms * 1e3)
        kern = {k: round(sorted(v)[len(v) // 2], 2) for k, v in agg.items()}
        # the whole step as a CUDA-graph replay (what overlaps, overlaps)
        gs = torch.cuda.Stream()
        gs.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(gs):
            with torch.cuda.graph(g, stream=gs):
                buf = ops.omc_step(pack, bs, N - bs, temp, seed=1, offset=0, buffers=buf)
        torch.cuda.current_stream().wait_stream(gs)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) / 50 * 1e3
        out.append(dict(bs=bs, cfg=cfg or "default", kernels_us=kern, sum_us=round(sum(kern.values()), 1), graph_us=round(graph_us, 1),
                        bit_equal_to_default=same, grad_rel_diff=close, loss=buf["loss"].item()))
        print(json.dumps(out[-1]), flush=True)
