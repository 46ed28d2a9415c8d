"""Fused pack + all-gather over NVLink 5 / NVSwitch peer memory (replaces `vast_pack_pair` + NCCL all-gather on the
training path, model/vast.py:395,404 / utils/distributed.py:50-66).

The gathered `[N, 2D]` bf16 buffer lives in symmetric memory (`torch.distributed._symmetric_memory`: the same
allocation mapped into every rank of the node, plus an NVSwitch multicast mapping where the driver offers one).
`vast_pack_pair_push` converts this rank's rows and stores every 16-byte vector straight into the buffer of ALL ranks
(one `multimem.st` to the multicast address -- more than four ranks -- or one peer store per rank); a
symmetric-memory barrier on the same stream publishes it.  Two buffers alternate so that a rank that is one step ahead
never overwrites rows a slower rank is still reading.  (An arrival-flag protocol without the barrier exists as an opt-in
experiment, see PackedGather.)"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from ._lib import check, dtype_code, lib, ptr, require_cuda, stream_ptr


class PackedGather:
    def __init__(self, bs: int, dim: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.bs, self.dim = bs, dim
        self.bufs, self.hdls = [], []
        for _ in range(2):
            b = symm.empty(self.world * bs, 2 * dim, dtype=torch.bfloat16, device=device)
            self.bufs.append(b)
            self.hdls.append(symm.rendezvous(b, self.group))
        # multimem.st pays ~20 us for the end-of-kernel flush of its replicated stores whatever the size, peer stores
        # pay (W - 1) x the bytes: measured on B200 + NVSwitch, peer stores win at W = 2 (109 vs 120 us/step) and
        # W = 4 (94.4 vs 96.9), multicast is what was measured at W = 8.  VAST_PEER_MULTICAST=0/1 overrides.
        env = os.environ.get("VAST_PEER_MULTICAST")
        self.use_multicast = (env == "1") if env in ("0", "1") else self.world > 4
        self.turn = 0
        # Opt-in experiment (VAST_PEER_SIGNAL=1): arrival flags instead of a barrier after the push
        # (vast_pack_pair_push_signal / vast_wait_arrivals) -- per buffer a uint32[world] array in symmetric memory that
        # the peers' push kernels write, plus local epochs; the waiting kernel starts while the push kernel's stores
        # still drain.  Correct (multi_gpu_check, bench parity) but measured SLOWER at 2 GPUs: 113.0 vs 106.9 us/step --
        # system-scope fences inside the push kernel (one per block, 296 blocks) cost more than the end-of-grid flush
        # plus the barrier kernel they replace.
        self.use_flags = os.environ.get("VAST_PEER_SIGNAL", "0") == "1"
        self.flags, self.fhdls, self.state = [], [], []
        if self.use_flags:
            for _ in range(2):
                f = symm.empty(32, dtype=torch.int32, device=device)
                f.zero_()
                self.flags.append(f)
                self.fhdls.append(symm.rendezvous(f, self.group))
                self.state.append(torch.zeros(4, dtype=torch.int32, device=device))
            torch.cuda.synchronize(device)
            for h in self.fhdls:
                h.barrier(channel=0)          # every rank's flags are zero before anyone signals
            torch.cuda.synchronize(device)

    def _dst(self, i):
        h = self.hdls[i]
        mc = int(getattr(h, "multicast_ptr", 0) or 0) if self.use_multicast else 0
        peers = (ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])
        return mc, peers

    @property
    def mode(self) -> str:
        mc, _ = self._dst(0)
        return "multimem.st to the NVSwitch multicast address" if mc else "peer stores to every rank"

    @torch.no_grad()
    def gather(self, feat_t: torch.Tensor, feat_cond: torch.Tensor, slot: int | None = None) -> torch.Tensor:
        """[N, 2D] bf16 = rows of (feat_t | feat_cond) of all ranks in rank order (valid on the current stream)."""
        require_cuda(feat_t, feat_cond)
        assert feat_t.shape == feat_cond.shape == (self.bs, self.dim) and feat_t.dtype == feat_cond.dtype
        i = self.turn if slot is None else slot
        self.turn = i ^ 1
        ft, fc = feat_t.contiguous(), feat_cond.contiguous()
        mc, peers = self._dst(i)
        if self.use_flags:
            fl = (ctypes.c_void_p * self.world)(*[int(p) for p in self.fhdls[i].buffer_ptrs])
            check(lib().vast_pack_pair_push_signal(ptr(ft), ptr(fc), dtype_code(ft.dtype), self.bs, self.dim, self.dim,
                                                   self.rank * self.bs, ctypes.c_void_p(mc) if mc else None, peers,
                                                   self.world, fl, self.rank, ptr(self.state[i]), stream_ptr()),
                  "pack_pair_push_signal")
            check(lib().vast_wait_arrivals(ptr(self.flags[i]), self.world, ptr(self.state[i]), stream_ptr()), "wait_arrivals")
            return self.bufs[i]
        check(lib().vast_pack_pair_push(ptr(ft), ptr(fc), dtype_code(ft.dtype), self.bs, self.dim, self.dim,
                                        self.rank * self.bs, ctypes.c_void_p(mc) if mc else None, peers, self.world,
                                        stream_ptr()), "pack_pair_push")
        self.hdls[i].barrier(channel=0)
        return self.bufs[i]


class PeerRows:
    """Per-rank [bs, ...] row blocks in symmetric memory for the negative-row exchange: the forward block
    (this rank's condition tokens, read by whoever sampled one of its rows) and the backward block (the gradients of the
    rows this rank fetched, pulled by their owners).  Two buffers of each alternate, one barrier per publish: a rank that
    is a step ahead never overwrites a block a slower rank is still reading."""

    def __init__(self, bs: int, row_shape, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.bs, self.row_shape, self.dtype = bs, tuple(row_shape), dtype
        self.bufs, self.hdls = {"fwd": [], "bwd": []}, {"fwd": [], "bwd": []}
        for kind in ("fwd", "bwd"):
            for _ in range(2):
                b = symm.empty((bs,) + self.row_shape, dtype=dtype, device=device)
                self.bufs[kind].append(b)
                self.hdls[kind].append(symm.rendezvous(b, self.group))
        self.turn = {"fwd": 0, "bwd": 0}

    @torch.no_grad()
    def publish(self, x: torch.Tensor, kind: str = "fwd"):
        """copy x [bs, ...] into this rank's block, barrier; returns every rank's block address (rank order)."""
        i = self.turn[kind]
        self.turn[kind] = i ^ 1
        self.bufs[kind][i].copy_(x)
        self.hdls[kind][i].barrier(channel=0)
        return [int(p) for p in self.hdls[kind][i].buffer_ptrs]


_cache: dict = {}
_rows_cache: dict = {}
_available: dict = {}


def _single_node(group=None) -> bool:
    """The symmetric-memory path needs every rank of the group on this node (peer / multicast mappings)."""
    w = dist.get_world_size(group)
    lw = os.environ.get("LOCAL_WORLD_SIZE")
    if lw is not None:
        return int(lw) == w
    return torch.cuda.device_count() >= w


def peer_rows(bs: int, row_shape, dtype, device) -> PeerRows | None:
    """Cached PeerRows for this block shape, or None (symmetric memory unavailable / VAST_PEER_GATHER=0): the caller then
    uses the all_to_all exchange.  Collective on first use per shape; the answer is the same on every rank."""
    if not _enabled():
        return None
    key = (bs, tuple(row_shape), dtype, torch.device(device).index)
    if key not in _rows_cache:
        _rows_cache[key] = _agree(lambda: PeerRows(bs, row_shape, dtype, torch.device(device)), torch.device(device),
                                  "peer-memory row exchange", "the all_to_all row exchange")
    return _rows_cache[key]


def _agree(factory, dev, what: str, fallback: str):
    """run a collective constructor; every rank gets the object or every rank gets None (all-reduce MIN of success)."""
    obj, err = None, None
    if _single_node():
        try:
            obj = factory()
        except Exception as e:  # symmetric memory not supported on this system / build
            err = e
    else:
        err = RuntimeError("ranks span more than one node")
    ok = torch.tensor([1 if obj is not None else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        if dist.get_rank() == 0:
            import warnings
            warnings.warn(f"vast_b200: {what} unavailable ({type(err).__name__ if err else 'on another rank'}: {err}); using {fallback}")
        return None
    return obj


def _build_collectively(bs: int, dim: int, device) -> PackedGather | None:
    """Construct a PackedGather or return None -- the SAME answer on every rank.  Construction is collective
    (symmetric allocation + rendezvous); if it fails on some ranks only (out of memory, a peer without P2P or multicast)
    and those fell back to NCCL while the others waited in the symmetric-memory barrier, the job would hang: so the
    outcome is agreed on with an all-reduce(MIN) before anyone uses it."""
    dev = torch.device(device)
    pg, err = None, None
    if _single_node():
        try:
            pg = PackedGather(bs, dim, dev)
        except Exception as e:  # symmetric memory not supported on this system / build
            err = e
    else:
        err = RuntimeError("ranks span more than one node")
    ok = torch.tensor([1 if pg is not None else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        if dist.get_rank() == 0:
            import warnings
            warnings.warn(f"vast_b200: peer-memory all-gather unavailable ({type(err).__name__ if err else 'on another rank'}: "
                          f"{err}); using vast_pack_pair + NCCL all_gather_into_tensor")
        return None
    return pg


def _enabled() -> bool:
    return os.environ.get("VAST_PEER_GATHER", "1") == "1" and dist.is_initialized() and dist.get_world_size() > 1


def packed_gather(bs: int, dim: int, device) -> PackedGather | None:
    """The EAGER path's PackedGather for this shape (cached: consecutive eager gathers alternate its two buffers), or
    None when symmetric memory is unavailable / disabled (VAST_PEER_GATHER=0): the caller then uses vast_pack_pair +
    NCCL all_gather_into_tensor.  Collective on first use per shape."""
    if not _enabled():
        return None
    key = (bs, dim, torch.device(device).index)
    if key not in _cache:
        _cache[key] = _build_collectively(bs, dim, device)
    return _cache[key]


def private_gather(bs: int, dim: int, device) -> PackedGather | None:
    """A PackedGather of its OWN for a captured step (OmcGraphStep).  The write-after-read argument of the two
    alternating buffers -- a rank that is one gather ahead never overwrites rows a slower rank still reads -- holds only
    if consecutive gathers on a pair of buffers alternate; two graphed steps (or a graphed step and the eager path)
    sharing one pair could push into the same buffer back to back.  Collective."""
    if not _enabled():
        return None
    return _build_collectively(bs, dim, device)
