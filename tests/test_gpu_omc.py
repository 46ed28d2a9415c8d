"""Fused OMC contrastive step (tcgen05 GEMMs with online-LSE / softmax+race epilogues, dQ GEMM)
vs the oracle (oracle/spec.py, pinned to the reference by tests/golden) through the C-ABI.

Tolerance (BASELINE.json north_star): loss and gradients within 1e-3 relative of the reference
math evaluated on the SAME bf16-rounded features (bf16-in / fp32-accumulate mode)."""
import numpy as np
import pytest
import torch

from oracle import spec

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def bf16_round(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).bfloat16().float().numpy()


def run_step(ft_all, fc_all, bs, rank, temp, **kw):
    from vast_b200 import ops
    pack = ops.pack_pair(torch.from_numpy(ft_all).cuda(), torch.from_numpy(fc_all).cuda())
    out = ops.omc_step(pack, bs, rank * bs, temp, **kw)
    torch.cuda.synchronize()
    return out


def oracle(ft_all, fc_all, bs, rank, temp):
    ft_r, fc_r = bf16_round(ft_all), bf16_round(fc_all)
    sl = slice(rank * bs, (rank + 1) * bs)
    return spec.omc_loss(fc_r[sl], ft_r[sl], ft_r, fc_r, temp, rank=rank)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def close(a, b):
    """norm-wise relative error below RTOL (absolute 1e-6 floor for degenerate all-zero cases)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= RTOL * np.linalg.norm(b) + 1e-6


def rows_close(got, want, temp):
    """Per-ROW gradient check (a norm-wise bound over the whole matrix lets single rows be far off).  A gradient row is
    (1 / (2 bs tau)) (sum_j p_ij k_j - (1 - eps) k_t - (eps / N) sum_j k_j): a difference of O(1)-norm terms, so its
    absolute error scales with g0 = 1 / (2 bs tau) whatever the row's own norm (a well-classified row's gradient nearly
    cancels).  Bound: |err_row| <= RTOL * |want_row| + 5e-4 * g0 -- i.e. every row to 1e-3 of its own norm unless the
    row is smaller than half the cancelling terms, and then to 5e-4 of those."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    g0 = 1.0 / (2.0 * want.shape[0] * temp)
    err = np.linalg.norm(got - want, axis=1)
    lim = RTOL * np.linalg.norm(want, axis=1) + 5e-4 * g0
    bad = np.nonzero(err > lim)[0]
    return bad.size == 0, (bad[:8], (err / lim).max())


def check_against_oracle(out, o, bs, temp=None):
    assert abs(out["loss"].item() - o["loss"]) <= RTOL * abs(o["loss"]) + 1e-6, (out["loss"].item(), o["loss"])
    assert close(out["grad_cond"].cpu().numpy(), o["grad_cond"]), rel(out["grad_cond"].cpu().numpy(), o["grad_cond"])
    assert close(out["grad_t"].cpu().numpy(), o["grad_t"]), rel(out["grad_t"].cpu().numpy(), o["grad_t"])
    if temp is not None:
        for k in ("grad_cond", "grad_t"):
            ok, info = rows_close(out[k].cpu().numpy(), o[k], temp)
            assert ok, (k, info)
    assert abs(out["grad_temp"].item() - o["grad_temp"]) <= RTOL * abs(o["grad_temp"]) + 1e-6
    lse = out["lse"].cpu().numpy()
    np.testing.assert_allclose(lse[0], o["lse_cond2t"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(lse[1], o["lse_t2cond"], rtol=1e-4, atol=1e-4)


def check_negatives(neg, o, expo, rank, bs, floor=1e-4):
    """GPU draw must be the oracle's argmax, or a numerical near-tie of it (fp32 vs fp64 keys)."""
    neg = neg.cpu().numpy()
    exact = 0
    for d, name in enumerate(("sim_cond2t", "sim_t2cond")):
        w = spec.hardneg_weights(o[name], rank, floor)
        key = w / expo[d]
        best = key.argmax(axis=1)
        exact += int((neg[d] == best).sum())
        got = key[np.arange(bs), neg[d]]
        # every mismatching row must be a near-tie of the race: its key within 2e-3 of the winning key (fp32 vs fp64)
        assert np.all(got >= key.max(axis=1) * (1 - 2e-3)), (d, np.nonzero(got < key.max(axis=1) * (1 - 2e-3)))
        assert np.all(neg[d] != rank * bs + np.arange(bs)), "positive sampled as negative"
    assert exact >= int(0.97 * 2 * bs), exact


def _hier_near_tie(z, tcol, expo, units, got, floor, tol=3e-3):
    """Why may the GPU draw `got` differ from the oracle's for this row?  Only through a numerical near-tie at one of
    the sampler's three decisions, each re-evaluated here in fp64 with slack `tol` (the GPU keeps probabilities in fp16
    and sums in fp32): (1) the race between chunks -- the chunk of `got` must be within tol of the winning key;
    (2) the mixture draw softmax-part vs uniform floor part; (3) the inverse CDF inside the chunk / the uniform index
    -- `got` must be a column whose cumulative-weight interval, widened by tol, contains the target."""
    n = z.shape[0]
    p = np.exp(z - (z.max() + np.log(np.exp(z - z.max()).sum())))
    pt = p[tcol]
    p = p.copy()
    p[tcol] = 0.0
    nch = (n + 31) // 32
    pad = np.zeros(nch * 32)
    pad[:n] = p
    cs = pad.reshape(nch, 32).sum(axis=1)
    key = cs / expo[:nch]
    w_a, w_b = 1.0 - pt, floor * (n - 1)
    mix = units[1] * (w_a + w_b)
    can_a = mix < w_a * (1 + tol)
    can_b = mix > w_a * (1 - tol) or cs.max() <= 0
    ok = False
    if can_a:
        c = got // 32
        if key[c] >= key.max() * (1 - tol) and pad[got] > 0:
            seg = pad[c * 32:(c + 1) * 32]
            pre = np.cumsum(seg)
            j = got - c * 32
            target = units[0] * pre[-1]
            lo_edge = pre[j] - seg[j]
            ok |= (lo_edge - tol * pre[-1] <= target <= pre[j] + tol * pre[-1]) or \
                  (j == np.nonzero(seg > 0)[0][-1] and target >= lo_edge - tol * pre[-1])
    if can_b:
        x = units[2] * (n - 1)
        for jj in {int(np.floor(x - 1e-3)), int(np.floor(x)), int(np.floor(x + 1e-3))}:
            jj = min(max(jj, 0), n - 2)
            ok |= got == (jj + 1 if jj >= tcol else jj)
    return bool(ok)


def check_negatives_hier(neg, o, seed, offset, rank, bs, n, floor=1e-4):
    """Production sampler (chunk race + inverse CDF + uniform floor component) vs its oracle restatement fed with the
    same Philox words: index-exact, except rows where a decision of the sampler is a numerical near-tie -- every
    mismatching row is audited (`_hier_near_tie`) and listed, and there may be only a few."""
    neg = neg.cpu().numpy()
    nch = (n + 31) // 32
    mism = []
    for d, name in enumerate(("sim_cond2t", "sim_t2cond")):
        expo = spec.sampler_expo(seed, offset, d, bs, nch, row0=rank * bs)
        units = spec.sampler_tail_units(seed, offset, d, bs, row0=rank * bs)
        want, _, _ = spec.hardneg_hier_sample(o[name], rank, expo, units, floor)
        if n > 1:
            assert np.all(neg[d] != rank * bs + np.arange(bs)), "positive sampled as negative"
            assert np.all((neg[d] >= 0) & (neg[d] < n))
            for i in np.nonzero(neg[d] != want)[0]:
                assert _hier_near_tie(np.asarray(o[name][i], dtype=np.float64), rank * bs + i, expo[i], units[i], int(neg[d][i]), floor), \
                    ("draw differs from the oracle and is not a near-tie", d, int(i), int(neg[d][i]), int(want[i]))
                mism.append((d, int(i)))
    assert len(mism) <= max(2, int(0.03 * 2 * bs)), (len(mism), mism[:10])


def test_cfg1_golden_loss_grads_and_reference_negatives(golden):
    """cfg1 (bs 64, D 512, W 1): loss/grads vs oracle on bf16-rounded inputs and vs the reference's
    fp32 numbers; with the reference's own Exp(1) noise the sampled negatives reproduce."""
    g = golden("omc_w1")
    bs = 64
    noise = torch.from_numpy(np.ascontiguousarray(g["expo"][::-1])).cuda()  # ours: [0]=cond2t, [1]=t2cond
    out = run_step(g["feat_t"], g["feat_cond"], bs, 0, float(g["contra_temp"]), debug_noise=noise, want_lse=True)
    o = oracle(g["feat_t"], g["feat_cond"], bs, 0, float(g["contra_temp"]))
    check_against_oracle(out, o, bs, float(g["contra_temp"]))
    # vs the reference run on un-rounded fp32 features: bf16 input rounding only
    assert abs(out["loss"].item() - float(g["loss_itc"])) < 1e-2 * float(g["loss_itc"])
    assert rel(out["grad_t"].cpu().numpy(), g["grad_t"]) < 2e-2
    assert rel(out["grad_cond"].cpu().numpy(), g["grad_cond"]) < 2e-2
    check_negatives(out["neg_idx"], o, g["expo"][::-1], 0, bs)
    neg = out["neg_idx"].cpu().numpy()
    assert (neg[0] == g["neg_cond2t"]).mean() > 0.9 and (neg[1] == g["neg_t2cond"]).mean() > 0.9


@pytest.mark.parametrize("rank", [0, 1])
def test_w2_rank_offsets_golden(golden, rank):
    g = golden("dist_w2")
    bs = 16
    temp = float(g["contra_temp"])
    noise = torch.from_numpy(np.ascontiguousarray(g[f"r{rank}_expo"][::-1])).cuda()
    out = run_step(g["feat_t_all"], g["feat_cond_all"], bs, rank, temp, debug_noise=noise, want_lse=True)
    o = oracle(g["feat_t_all"], g["feat_cond_all"], bs, rank, temp)
    check_against_oracle(out, o, bs, temp)
    assert abs(out["loss"].item() - float(g[f"r{rank}_loss_itc"])) < 1e-2 * float(g[f"r{rank}_loss_itc"])
    check_negatives(out["neg_idx"], o, g[f"r{rank}_expo"][::-1], rank, bs)


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("bs,world,rank,dim,temp", [(100, 3, 1, 72, 0.07), (1, 1, 0, 8, 0.5), (130, 2, 1, 264, 0.02),
                                                     (256, 4, 3, 512, 0.07)])
def test_ragged_shapes_philox(bs, world, rank, dim, temp, two_pass):
    n = bs * world
    gen = torch.Generator().manual_seed(bs + dim)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, dim, generator=gen), dim=-1)
    seed, offset = 0x1234567887654321, (7 << 32) | 5
    out = run_step(t.numpy(), c.numpy(), bs, rank, temp, seed=seed, offset=offset, want_lse=True, two_pass=two_pass)
    o = oracle(t.numpy(), c.numpy(), bs, rank, temp)
    check_against_oracle(out, o, bs, temp)
    check_negatives_hier(out["neg_idx"], o, seed, offset, rank, bs, n)


@pytest.mark.parametrize("case", ["hard_rows", "small_tau", "untrained"])
def test_single_pass_range_and_fallback(case):
    """The single-pass form keeps softmax numerators relative to the positive pair in fp16.  Rows whose best
    negative beats the positive by more than 16 ln2 nats overflow that range: the on-device fallback must
    redo the step in two passes and still match the oracle.  'untrained' (positives indistinguishable from
    negatives) must stay within range and match as well."""
    n, dim = 192, 128
    gen = torch.Generator().manual_seed(5)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    if case == "untrained":
        c = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
        temp = 0.07
    else:
        c = torch.nn.functional.normalize(t + 0.5 * torch.randn(n, dim, generator=gen), dim=-1)
        temp = 0.07
        if case == "hard_rows":   # mislabeled pairs: positive anti-correlated, a perfect negative elsewhere
            c[3] = -t[3]
            c[77] = t[3]
            c[150] = -t[150]
        else:
            temp = 0.004          # 1 / tau = 250: ordinary negatives already exceed the range
    seed, offset = 17, 9
    out = run_step(t.numpy(), c.numpy(), n, 0, temp, seed=seed, offset=offset, want_lse=True)
    ref = run_step(t.numpy(), c.numpy(), n, 0, temp, seed=seed, offset=offset, want_lse=True, two_pass=True)
    o = oracle(t.numpy(), c.numpy(), n, 0, temp)
    check_against_oracle(out, o, n, temp)
    check_against_oracle(ref, o, n, temp)
    check_negatives_hier(out["neg_idx"], o, seed, offset, 0, n, n)
    if case != "untrained":       # the fallback IS the two-pass form: identical bits
        assert torch.equal(out["grad_t"], ref["grad_t"]) and torch.equal(out["neg_idx"], ref["neg_idx"])


def test_cfg3_shape_full_size():
    """BASELINE cfg3 at W=1: N = 4096, D = 1024 (oracle in fp64 numpy, a few seconds)."""
    n, dim, temp = 4096, 1024, 0.07
    gen = torch.Generator().manual_seed(1234)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, dim, generator=gen), dim=-1)
    seed, offset = 99, 3
    out = run_step(t.numpy(), c.numpy(), n, 0, temp, seed=seed, offset=offset, want_lse=True)
    o = oracle(t.numpy(), c.numpy(), n, 0, temp)
    check_against_oracle(out, o, n, temp)
    check_negatives_hier(out["neg_idx"], o, seed, offset, 0, n, n)
    # determinism: same seed/offset -> identical outputs, different offset -> different draws
    out2 = run_step(t.numpy(), c.numpy(), n, 0, temp, seed=seed, offset=offset)
    assert torch.equal(out["neg_idx"], out2["neg_idx"]) and torch.equal(out["grad_t"], out2["grad_t"])
    assert out["loss"].item() == out2["loss"].item()
    out3 = run_step(t.numpy(), c.numpy(), n, 0, temp, seed=seed, offset=offset + 1)
    assert (out3["neg_idx"] != out["neg_idx"]).float().mean().item() > 0.3


def test_sampler_distribution_chi2():
    """Distributional parity with torch.multinomial: empirical frequencies over many Philox offsets
    follow w / sum(w) (chi-square, df = N - 2 per row)."""
    from vast_b200 import ops
    bs = n = 8
    gen = torch.Generator().manual_seed(7)
    t = torch.nn.functional.normalize(torch.randn(n, 16, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 1.0 * torch.randn(n, 16, generator=gen), dim=-1)
    temp = 0.3
    pack = ops.pack_pair(t.cuda(), c.cuda())
    draws = 3000
    counts = np.zeros((2, bs, n))
    res = []
    for k in range(draws):
        res.append(ops.omc_step(pack, bs, 0, temp, seed=42, offset=k, need_grad=False)["neg_idx"])
    neg = torch.stack(res).cpu().numpy()  # [draws, 2, bs]
    for d in range(2):
        for b in range(bs):
            counts[d, b] = np.bincount(neg[:, d, b], minlength=n)
    o = oracle(t.numpy(), c.numpy(), bs, 0, temp)
    for d, name in enumerate(("sim_cond2t", "sim_t2cond")):
        w = spec.hardneg_weights(o[name], 0)
        p = w / w.sum(axis=1, keepdims=True)
        exp = p * draws
        for b in range(bs):
            assert counts[d, b, b] == 0
            m = exp[b] > 0
            chi2 = ((counts[d, b][m] - exp[b][m]) ** 2 / exp[b][m]).sum()
            assert chi2 < 35.0, (d, b, chi2)  # df = 6: P(chi2 > 35) ~ 4e-6


def test_loss_only_and_errors():
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(8)
    t = torch.nn.functional.normalize(torch.randn(40, 64, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + torch.randn(40, 64, generator=gen), dim=-1)
    pack = ops.pack_pair(t.cuda(), c.cuda())
    out = ops.omc_step(pack, 40, 0, 0.07, need_sample=False, need_grad=False)
    o = oracle(t.numpy(), c.numpy(), 40, 0, 0.07)
    assert abs(out["loss"].item() - o["loss"]) <= RTOL * abs(o["loss"])
    with pytest.raises(RuntimeError):
        ops.omc_step(pack, 40, 8, 0.07)  # local rows outside [0, n_total)
    with pytest.raises(RuntimeError):
        ops.omc_step(pack, 40, 0, -1.0)


def test_step_counter_advances_the_philox_offset():
    """A device-side step counter makes the step replayable (CUDA graphs) with fresh noise: call i with counter c
    draws exactly what a call with offset + c draws, and the library increments the counter."""
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(21)
    n, d = 160, 64
    t = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, d, generator=gen), dim=-1)
    pack = ops.pack_pair(t.cuda(), c.cuda())
    ctr = torch.zeros(1, dtype=torch.int64, device="cuda")
    a0 = ops.omc_step(pack, n, 0, 0.07, seed=5, offset=40, step_counter=ctr)["neg_idx"].clone()
    a1 = ops.omc_step(pack, n, 0, 0.07, seed=5, offset=40, step_counter=ctr)["neg_idx"].clone()
    assert ctr.item() == 2
    b0 = ops.omc_step(pack, n, 0, 0.07, seed=5, offset=40)["neg_idx"]
    b1 = ops.omc_step(pack, n, 0, 0.07, seed=5, offset=41)["neg_idx"]
    assert torch.equal(a0, b0) and torch.equal(a1, b1) and not torch.equal(a0, a1)
    # and the same through a captured graph
    g = torch.cuda.CUDAGraph()
    buf = ops.omc_step(pack, n, 0, 0.07, seed=5, offset=40, step_counter=ctr)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            buf = ops.omc_step(pack, n, 0, 0.07, seed=5, offset=40, step_counter=ctr, buffers=buf)
    torch.cuda.current_stream().wait_stream(s)
    ctr.zero_()
    g.replay()
    r0 = buf["neg_idx"].clone()
    g.replay()
    r1 = buf["neg_idx"].clone()
    torch.cuda.synchronize()
    assert torch.equal(r0, b0) and torch.equal(r1, b1) and ctr.item() == 2


@pytest.mark.parametrize("bs,world,rank,dim,noise", [(4096, 1, 0, 1024, False), (2048, 2, 1, 1024, False), (1024, 1, 0, 512, False),
                                                      (300, 1, 0, 520, False), (512, 4, 2, 256, False), (64, 1, 0, 512, True),
                                                      (640, 1, 0, 384, True)])
def test_fused_row_stats_equal_separate_kernel(bs, world, rank, dim, noise):
    """The row statistics + hard-negative draw run inside the dQ GEMM's epilogue by default; the stand-alone kernel
    (VAST_OMC_SEPARATE_ROW_STATS) must give the same bits: gradients, negatives, lse (the loss and d tau sum <q, sum K>
    in a different order: 1e-6)."""
    n = bs * world
    gen = torch.Generator().manual_seed(3 * bs + dim)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, dim, generator=gen), dim=-1)
    kw = dict(seed=77, offset=5, want_lse=True)
    if noise:
        kw["debug_noise"] = torch.empty(2, bs, n).exponential_(generator=gen).cuda()
    a = run_step(t.numpy(), c.numpy(), bs, rank, 0.07, separate_row_stats=False, **kw)
    b = run_step(t.numpy(), c.numpy(), bs, rank, 0.07, separate_row_stats=True, **kw)
    for k in ("grad_t", "grad_cond", "neg_idx", "lse"):
        assert torch.equal(a[k], b[k]), k
    assert abs(a["loss"].item() - b["loss"].item()) <= 2e-6 * abs(b["loss"].item())
    assert abs(a["grad_temp"].item() - b["grad_temp"].item()) <= 1e-5 * abs(b["grad_temp"].item()) + 1e-7
    o = oracle(t.numpy(), c.numpy(), bs, rank, 0.07)
    check_against_oracle(a, o, bs, 0.07)


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("n,dim,dtype", [(4096, 1024, torch.float32), (100, 72, torch.float32), (1, 8, torch.float32),
                                         (130, 264, torch.bfloat16), (333, 136, torch.float16), (64, 128, torch.bfloat16),
                                         (65, 1032, torch.float32)])
def test_single_rank_step_from_features(n, dim, dtype, two_pass):
    """vast_omc_step_local (one rank: packing fused into the step's first kernel) vs the oracle, its `pack` output
    bit-identical to vast_pack_pair, negatives vs the sampler's restatement."""
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(n + dim)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, dim, generator=gen), dim=-1)
    td, cd = t.to(dtype).cuda(), c.to(dtype).cuda()
    seed, offset = 4242, 11
    temp = 0.5 if n == 1 else 0.07   # n = 1: loss and gradients are exactly 0, only fp32 rounding times 1 / tau^2 remains
    out = ops.omc_step_local(td, cd, temp, seed=seed, offset=offset, want_lse=True, two_pass=two_pass)
    torch.cuda.synchronize()
    assert torch.equal(out["pack"], ops.pack_pair(td, cd))
    tn, cn = td.float().cpu().numpy(), cd.float().cpu().numpy()
    o = oracle(tn, cn, n, 0, temp)
    check_against_oracle(out, o, n, temp)
    check_negatives_hier(out["neg_idx"], o, seed, offset, 0, n, n)
    # against the two-kernel form: same operands, z_t summed in another order -> agreement far inside the tolerance
    ref = ops.omc_step(ops.pack_pair(td, cd), n, 0, temp, seed=seed, offset=offset, want_lse=True, two_pass=two_pass)
    assert rel(out["grad_t"].cpu().numpy(), ref["grad_t"].cpu().numpy()) < 1e-4
    assert abs(out["loss"].item() - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item()) + 1e-6
    assert (out["neg_idx"] == ref["neg_idx"]).float().mean().item() >= 0.97 or n < 8
    # loss-only / no sampling, and reuse of the returned buffers
    lo = ops.omc_step_local(td, cd, temp, need_sample=False, need_grad=False)
    assert abs(lo["loss"].item() - o["loss"]) <= RTOL * abs(o["loss"]) + 1e-6
    again = ops.omc_step_local(td, cd, temp, seed=seed, offset=offset, want_lse=True, two_pass=two_pass, buffers=out)
    torch.cuda.synchronize()
    assert again["loss"] is out["loss"]


def test_single_rank_step_rejects_unaligned():
    from vast_b200 import ops
    x = torch.randn(16, 12).cuda()
    with pytest.raises(RuntimeError):
        ops.omc_step_local(x, x, 0.07)     # D % 8 != 0


def test_assume_in_range_skips_fallback_and_poisons_on_violation(monkeypatch):
    """VAST_OMC_ASSUME_IN_RANGE: without the gated fallback launches an in-range step gives the same bits as the
    default; a step whose numerators leave the fp16 range reports NaN for the loss and d tau (never wrong numbers),
    and the workspace is still left clean for the next step."""
    from vast_b200 import ops
    n, dim = 192, 128
    gen = torch.Generator().manual_seed(5)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.5 * torch.randn(n, dim, generator=gen), dim=-1)
    ref = run_step(t.numpy(), c.numpy(), n, 0, 0.07, seed=3, offset=1)
    monkeypatch.setenv("VAST_OMC_ASSUME_IN_RANGE", "1")
    out = run_step(t.numpy(), c.numpy(), n, 0, 0.07, seed=3, offset=1)
    assert torch.equal(out["grad_t"], ref["grad_t"]) and torch.equal(out["neg_idx"], ref["neg_idx"])
    assert out["loss"].item() == ref["loss"].item()
    bad = c.clone()
    bad[3] = -t[3]
    bad[77] = t[3]                       # a perfect negative against an anti-correlated positive: 2 / 0.07 = 28 nats
    pack = ops.pack_pair(t.cuda(), bad.cuda())
    o1 = ops.omc_step(pack, n, 0, 0.07, seed=3, offset=1)
    assert torch.isnan(o1["loss"]).item() and torch.isnan(o1["grad_temp"]).item()
    # the poisoned step left the flag block clean: re-using its buffers for an in-range step gives the right numbers
    o2 = ops.omc_step(ops.pack_pair(t.cuda(), c.cuda()), n, 0, 0.07, seed=3, offset=1, buffers=o1)
    assert o2["loss"].item() == ref["loss"].item() and torch.equal(o2["grad_t"], ref["grad_t"])
    monkeypatch.delenv("VAST_OMC_ASSUME_IN_RANGE")
    o3 = ops.omc_step(pack, n, 0, 0.07, seed=3, offset=1)          # default: the fallback handles it
    assert torch.isfinite(o3["loss"]).item()


def test_workspace_cap_refused_cleanly_and_row_chunks_equal_whole(monkeypatch):
    """The Pt workspace is O(bs * n_total).  Beyond the cap (VAST_OMC_MAX_WS_GB) the C-ABI refuses with
    VAST_ERR_UNSUPPORTED and a message; the Python op runs the step on row chunks instead -- same loss, gradients, lse and
    the SAME negatives (Philox words are keyed by the global row) as the whole step."""
    from vast_b200 import ops
    n, dim = 1024, 128
    gen = torch.Generator().manual_seed(12)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    c = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, dim, generator=gen), dim=-1)
    pack = ops.pack_pair(t.cuda(), c.cuda())
    whole = ops.omc_step(pack, n, 0, 0.07, seed=9, offset=4, want_lse=True)
    need = ops.lib().vast_omc_workspace_bytes(n, n, dim, 1, 1)
    monkeypatch.setenv("VAST_OMC_MAX_WS_GB", str(need / 2 ** 30 / 3))       # a third of what the whole step needs
    with pytest.raises(RuntimeError, match="VAST_OMC_MAX_WS_GB"):            # the C-ABI itself refuses (status -2), no crash
        ops.omc_step(pack, n, 0, 0.07, seed=9, offset=4, want_lse=True, buffers=whole)
    parts = ops.omc_step(pack, n, 0, 0.07, seed=9, offset=4, want_lse=True)
    assert parts["_ws"][0] is None                                          # went through the chunked path
    assert torch.equal(parts["neg_idx"], whole["neg_idx"])
    assert abs(parts["loss"].item() - whole["loss"].item()) < 1e-6 * abs(whole["loss"].item())
    for k in ("grad_t", "grad_cond", "lse"):
        assert torch.allclose(parts[k], whole[k], rtol=1e-5, atol=1e-9), k
    assert abs(parts["grad_temp"].item() - whole["grad_temp"].item()) < 1e-5 * abs(whole["grad_temp"].item())
    loc = ops.omc_step_local(t.cuda(), c.cuda(), 0.07, seed=9, offset=4)    # the single-rank entry chunks as well
    assert torch.equal(loc["neg_idx"], whole["neg_idx"]) and torch.allclose(loc["grad_t"], whole["grad_t"], rtol=1e-5, atol=1e-9)
    o = oracle(t.numpy(), c.numpy(), n, 0, 0.07)
    check_against_oracle(parts, o, n, 0.07)


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("case", ["in_range", "hard_rows", "wide_spread", "small_tau"])
def test_single_rank_symmetric_form_and_its_fallback(case, fused, monkeypatch):
    """One rank: the symmetric form (one S GEMM problem for both directions) -- as two kernels (default) or, opt-in, as
    ONE persistent kernel (omc_fused_gemm_kernel, VAST_OMC_FUSED=1) whose dQ items wait on per-row-block / per-column-
    tile completion counters.  Range violations -- a negative that beats its positive by more than the fp16 range
    ('hard_rows', 'small_tau'), or target logits spread over more than 9 log2 units ('wide_spread': half the pairs
    untrained) -- raise the flag and the gated launches redo the step in the two-pass, two-problem form; either way the
    result matches the oracle, and the workspace is left clean for the next step."""
    from vast_b200 import ops
    monkeypatch.setenv("VAST_OMC_FUSED", fused)
    n, dim, temp = 2048, 1024, 0.07     # whole-K 256-wide pair tiles: the shape class of the headline (symmetric form in force)
    gen = torch.Generator().manual_seed(41)
    t0 = torch.randn(n, dim, generator=gen)
    t = torch.nn.functional.normalize(t0, dim=-1)
    c = torch.nn.functional.normalize(t0 + 0.5 * torch.randn(n, dim, generator=gen), dim=-1)   # trained pairs: cosine ~0.89
    if case == "hard_rows":
        c[3] = -t[3]
        c[77] = t[3]
    elif case == "wide_spread":
        c[n // 2:] = torch.nn.functional.normalize(torch.randn(n - n // 2, dim, generator=gen), dim=-1)
    elif case == "small_tau":
        temp = 0.004
    seed, offset = 23, 2
    out = ops.omc_step_local(t.cuda(), c.cuda(), temp, seed=seed, offset=offset, want_lse=True)
    o = oracle(t.numpy(), c.numpy(), n, 0, temp)
    check_against_oracle(out, o, n, temp)
    check_negatives_hier(out["neg_idx"], o, seed, offset, 0, n, n)
    if case != "in_range":     # the fallback IS the two-pass form
        ref = ops.omc_step_local(t.cuda(), c.cuda(), temp, seed=seed, offset=offset, want_lse=True, two_pass=True)
        assert torch.equal(out["grad_t"], ref["grad_t"]) and torch.equal(out["neg_idx"], ref["neg_idx"])
    # re-use of the (self-cleaned) workspace: a clean step after a fallback step gives the clean step's numbers
    good_c = torch.nn.functional.normalize(t0 + 0.5 * torch.randn(n, dim, generator=torch.Generator().manual_seed(5)), dim=-1)
    a = ops.omc_step_local(t.cuda(), good_c.cuda(), 0.07, seed=seed, offset=offset, want_lse=True, buffers=out)
    b = ops.omc_step_local(t.cuda(), good_c.cuda(), 0.07, seed=seed, offset=offset)
    torch.cuda.synchronize()
    assert a["loss"].item() == b["loss"].item() and torch.equal(a["grad_t"], b["grad_t"]) and torch.equal(a["neg_idx"], b["neg_idx"])


@pytest.mark.gpu
@pytest.mark.parametrize("bs", [512, 1024])
def test_cluster_split_k_dq_equals_default(bs, monkeypatch):
    """VAST_OMC_KSPLIT=1: the dQ GEMM of a small per-rank batch as clusters of two CTA pairs per output tile, the second
    pair's accumulator handed over through distributed shared memory.  Same step, same negatives, gradients equal to
    the default tiling up to the fp32 summation order."""
    from vast_b200 import ops
    n, dim = 4096, 1024
    g = torch.Generator().manual_seed(11)
    t = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=-1)
    c = torch.nn.functional.normalize(t + 0.7 * torch.randn(n, dim, generator=g), dim=-1)
    pack = ops.pack_pair(t.cuda(), c.cuda())
    temp = torch.full((1,), 0.07, device="cuda")
    monkeypatch.delenv("VAST_OMC_KSPLIT", raising=False)
    ref = ops.omc_step(pack, bs, n - 2 * bs, temp, seed=5, offset=0)
    ref = {k: ref[k].clone() for k in ("loss", "grad_t", "grad_cond", "grad_temp", "neg_idx")}
    monkeypatch.setenv("VAST_OMC_KSPLIT", "1")
    out = ops.omc_step(pack, bs, n - 2 * bs, temp, seed=5, offset=0)
    torch.cuda.synchronize()
    assert torch.equal(out["neg_idx"], ref["neg_idx"])
    assert abs(out["loss"].item() - ref["loss"].item()) <= 1e-6 * abs(ref["loss"].item())
    for k in ("grad_t", "grad_cond"):
        err = (out[k] - ref[k]).norm(dim=1) / ref[k].norm(dim=1).clamp_min(1e-30)
        assert err.max().item() < 1e-5, (k, err.max().item())
    assert abs(out["grad_temp"].item() - ref["grad_temp"].item()) <= 1e-5 * abs(ref["grad_temp"].item()) + 1e-9
