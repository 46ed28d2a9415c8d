"""world_size-2 gloo (CPU) tests of the data-parallel helpers (vast_b200/distributed.py) against the
golden outputs of the reference's utils/distributed.py, and of the host-side shard logic."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vast_b200 import distributed as D
    out = {}
    x = torch.arange(rank * 10, rank * 10 + 6, dtype=torch.float32).reshape(3, 2)
    out["concat_all_gather"] = D.concat_all_gather(x).numpy()
    ragged = torch.arange((rank + 2) * 3, dtype=torch.float32).reshape(rank + 2, 3) + 100 * rank
    out["ddp_allgather"] = D.ddp_allgather(ragged).numpy()
    out["all_gather_list"] = np.array([len(v) for v in D.all_gather_list(list(range(rank + 1)))])
    xg = x.clone().requires_grad_()
    yg = D.all_gather_with_grad(xg)
    (yg * torch.arange(yg.numel(), dtype=torch.float32).reshape(yg.shape)).sum().backward()
    out["agwg_out"] = yg.detach().numpy()
    out["agwg_grad"] = xg.grad.numpy()
    out["bcast"] = D.any_broadcast({"r": rank}, 1)["r"]
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_collectives_match_reference_w2():
    g = np.load(os.path.join(GOLD, "dist_w2.npz"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    for r in range(2):
        for k in ("concat_all_gather", "ddp_allgather", "all_gather_list", "agwg_out", "agwg_grad"):
            assert np.array_equal(res[r][k], g[f"r{r}_{k}"]), (r, k)
        assert res[r]["bcast"] == 1


def test_single_process_identities():
    from vast_b200 import distributed as D
    x = torch.randn(4, 3)
    assert torch.equal(D.concat_all_gather(x), x) and torch.equal(D.ddp_allgather(x), x)
    assert D.all_gather_with_grad(x) is x and D.all_gather_list(5) == [5] and D.any_broadcast("a", 0) == "a"


def test_shard_bounds_cover_columns():
    from vast_b200.retrieval import _shard_bounds
    for n in (1, 7, 100000, 12501):
        for w in (1, 2, 4, 8):
            spans = [_shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_metric_helpers_host_logic():
    from vast_b200.retrieval import _backward_pairs, _first_index, _format
    assert _first_index(["a", "b", "a"]) == {"a": 0, "b": 1}
    rows, cols = _backward_pairs(["v0", "v1"], ["v0", "v0", "v1"])
    assert rows == [0, 1, 2] and cols == [0, 0, 1]
    assert _format("forward", 0.365, 0.635, 0.75) == {"forward_r1": 36.5, "forward_recall": "36.5/63.5/75.0",
                                                       "forward_ravg": 58.3}


def _exchange_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vast_b200 import distributed as D
    bs = 5
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(bs, 3, 4, generator=g)
    idx = torch.randint(0, world * bs, (bs,), generator=g)
    wgt = torch.randn(bs, 3, 4, generator=g)
    # reference path (model/vast.py:422 + indexing): gather everything with gradient, then index
    xa = x.clone().requires_grad_()
    ref = D.all_gather_with_grad(xa)[idx]
    (ref * wgt).sum().backward()
    # exchange path: only the requested rows travel
    xb = x.clone().requires_grad_()
    got = D.exchange_rows(xb, idx)
    (got * wgt).sum().backward()
    q.put((rank, bool(torch.equal(got, ref)), bool(torch.allclose(xb.grad, xa.grad, atol=1e-6)), float(xa.grad.abs().sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_exchange_rows_equals_gather_with_grad_then_index(world):
    """negative-row exchange == all_gather_with_grad(condition_feats)[neg_idx], values and gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, 29750 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] and r[2] for r in res), res
    assert any(r[3] > 0 for r in res)


def test_exchange_rows_single_process():
    from vast_b200 import distributed as D
    x = torch.randn(6, 2, requires_grad=True)
    idx = torch.tensor([5, 0, 0, 3, 2, 5])
    y = D.exchange_rows(x, idx)
    assert torch.equal(y, x[idx])
    y.sum().backward()
    assert torch.equal(x.grad[:, 0], torch.tensor([2., 0., 1., 1., 0., 2.]))


def _ids_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vast_b200 import distributed as D
    cases = {
        "str": [f"vidéo{rank}_{i}" for i in range(rank + 2)],              # ragged, non-ASCII
        "int": [1000 * rank + i for i in range(3 - rank)] if rank < 3 else [],
        "empty_on_some": ["only0"] if rank == 0 else [],
        "mixed": [rank, f"s{rank}"],                                        # falls back to the pickled gather
        "tuple": [(rank, 1)],
    }
    out = {}
    for k, v in cases.items():
        got = D.all_gather_ids(v)
        want = [j for i in D.all_gather_list(v) for j in i]                 # the reference's expression
        out[k] = (got == want, [type(x).__name__ for x in got] == [type(x).__name__ for x in want], len(got))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_all_gather_ids_equals_pickled_gather(world):
    """ids gathered as sized tensor all-gathers == `[j for i in all_gather_list(ids) for j in i]`
    (evaluation_mm.py:208-209) for string / integer / empty / mixed id lists."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ids_worker, args=(r, world, 29760 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, out in res:
        for k, (same, same_types, n) in out.items():
            assert same and same_types, (rank, k)
    assert res[0][1]["str"][2] == sum(r + 2 for r in range(world))


def test_all_gather_ids_single_process():
    from vast_b200 import distributed as D
    assert D.all_gather_ids(["a", "b"]) == ["a", "b"] and D.all_gather_ids([]) == [] and D.all_gather_ids((3, 4)) == [3, 4]
