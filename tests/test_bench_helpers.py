"""bench.py host-side helpers (no GPU): the peak table, the ncu-summary parser behind `roofline.traffic`, the synthetic
inputs of SURVEY 8d, and the reference arm's JSON line shape (the contract keys the driver reads)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_peaks_and_traffic_parsers():
    import bench
    pk = bench.load_peaks()
    assert pk["hbm"] > 1000 and pk["tf_sustained"] > 100 and pk["tf_burst"] >= pk["tf_sustained"]
    for kernel in ("omc_dq_gemm", "sim_topk_gemm"):
        tr = bench.ncu_traffic(kernel)
        assert tr is not None and tr["bytes"] > 1e6 and tr["source"].endswith("_ncu_full.txt")
        assert os.path.exists(os.path.join(ROOT, "profiles", tr["source"]))
    assert bench.ncu_traffic("no_such_kernel") is None


def test_synthetic_inputs_follow_the_survey():
    import torch
    import bench
    t, c = bench.synth(64, 32, 1234)
    assert t.shape == c.shape == (64, 32) and t.dtype == torch.float32
    assert torch.allclose(t.norm(dim=1), torch.ones(64), atol=1e-5) and torch.allclose(c.norm(dim=1), torch.ones(64), atol=1e-5)
    assert ((t * c).sum(dim=1) > 0.5).all()                 # correlated positives (c = t + 0.8 randn)
    t2, _ = bench.synth(64, 32, 1234, rows=slice(16, 32))   # a rank's slice of the same global batch
    assert torch.equal(t2, t[16:32])


@pytest.mark.timeout(300)
def test_cpu_baseline_objects_have_the_contract_keys(monkeypatch):
    """cpu_baseline of both metrics on a tiny stand-in problem (the real sizes are what bench.py times on the box)."""
    import bench
    monkeypatch.setattr(bench, "N_GLOBAL", 128)
    monkeypatch.setattr(bench, "DIM", 64)
    cb = bench.cpu_contrastive(sample_steps=1)
    assert cb["kind"] == "port" and cb["unit"] == "pairs/s" and cb["value"] > 0 and cb["cores"] >= 1 and "sample" in cb
    import torch
    g = torch.Generator().manual_seed(0)
    monkeypatch.setattr(bench, "RET_N", 512)
    cr = bench.cpu_retrieval(torch.randn(512, 64, generator=g), torch.randn(512, 64, generator=g), rows=128, repeats=1)
    assert cr["extrapolated"] is True and cr["unit"] == "queries/s" and cr["value"] > 0
