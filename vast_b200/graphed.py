"""The fused contrastive step (model/vast.py:395-440 + backward) as ONE CUDA-graph launch per call.

`omc_loss_and_negatives` enqueues about a dozen kernels per call through ctypes; on a B200 the kernels take
~150 us at the headline size while Python + launch overhead takes longer than that, so a training loop that reads
its loss every step is host-bound.  `OmcGraphStep` captures the same launches (pack, the all-gather push when
world_size > 1, `vast_omc_step`) once per input slot and replays them: the host cost of a step is one graph launch.

    step = vast_b200.OmcGraphStep(bs, dim, model.contra_temp)         # collective when world_size > 1
    loss, neg_text, neg_cond = step(feat_cond, feat_t)                # same returns as omc_loss_and_negatives
    loss.backward()

Lifetime rule (the usual one for graphed callables): the step owns two sets of static outputs that alternate, so the
tensors returned by call i are overwritten by call i+2.  `backward()` of a loss whose gradients were overwritten
raises instead of returning stale values.  Hard negatives draw fresh Philox noise on every replay through the
device-side step counter of `vast_omc_step`."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from ._lib import require_cuda
from .distributed import _rank, _world


class _GraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_cond, feat_t, contra_temp, owner, slot):
        owner._graphs[slot].replay()
        owner._gen[slot] += 1
        out = owner._out[slot]
        ctx.owner, ctx.slot, ctx.gen = owner, slot, owner._gen[slot]
        ctx.temp_shape = contra_temp.shape if isinstance(contra_temp, torch.Tensor) else None
        ctx.dtypes = (feat_cond.dtype, feat_t.dtype)
        loss = out["loss"].reshape(())
        if owner.need_negatives:
            neg_text, neg_cond = out["neg_idx"][0], out["neg_idx"][1]
            ctx.mark_non_differentiable(neg_text, neg_cond)
            return loss, neg_text, neg_cond
        return loss, None, None

    @staticmethod
    def backward(ctx, g, _a, _b):
        owner, s = ctx.owner, ctx.slot
        if owner._gen[s] != ctx.gen:
            raise RuntimeError("OmcGraphStep: the gradients of this loss were overwritten by a later call "
                               "(outputs of call i live until call i+2); call backward() earlier")
        out = owner._out[s]
        need = ctx.needs_input_grad
        want_temp = ctx.temp_shape is not None and need[2]
        src = [out[k] for k, w in (("grad_cond", need[0]), ("grad_t", need[1]), ("grad_temp", want_temp)) if w]
        scaled = iter(torch._foreach_mul(src, g) if src else ())       # one launch for all three products
        d_cond = next(scaled).to(ctx.dtypes[0]) if need[0] else None
        d_t = next(scaled).to(ctx.dtypes[1]) if need[1] else None
        d_temp = next(scaled).reshape(ctx.temp_shape) if want_temp else None
        return d_cond, d_t, d_temp, None, None


class OmcGraphStep:
    """Captured fused OMC step for fixed shapes.  Construction is collective when world_size > 1 (it runs the
    step twice eagerly and captures it); every rank must then call the step the same number of times."""

    def __init__(self, bs: int, dim: int, contra_temp, *, rank: int | None = None, world_size: int | None = None,
                 dtype: torch.dtype = torch.float32, device=None, label_smoothing: float = 0.1,
                 weight_floor: float = 1e-4, need_negatives: bool = True, seed: int | None = None):
        self.rank = _rank() if rank is None else rank
        self.world = _world() if world_size is None else world_size
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.bs, self.dim, self.device = bs, dim, dev
        self.need_negatives = need_negatives
        self.ls, self.wf = float(label_smoothing), float(weight_floor)
        self.seed = torch.initial_seed() if seed is None else int(seed)
        if isinstance(contra_temp, torch.Tensor):
            require_cuda(contra_temp)
            if contra_temp.dtype != torch.float32 or contra_temp.numel() != 1:
                raise RuntimeError("OmcGraphStep: contra_temp must be a float or a 1-element fp32 CUDA tensor "
                                   "(its storage is read in place by the captured kernels)")
            self._temp_arg = contra_temp
            self._temp = contra_temp.detach().reshape(1)          # a view: the graph reads the live parameter
        else:
            self._temp_arg = float(contra_temp)
            self._temp = torch.full((1,), float(contra_temp), device=dev)
        self._temp_ptr = self._temp.data_ptr()
        # static inputs: one [2, bs, dim] block (feat_t, feat_cond) per slot, so a loader can fill a slot with ONE copy
        self._in = [torch.zeros(2, bs, dim, dtype=dtype, device=dev) for _ in range(2)]
        self.feat_t_in = [b[0] for b in self._in]
        self.feat_cond_in = [b[1] for b in self._in]
        self._ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        self._out = [None, None]
        self._gen = [0, 0]
        self._turn = 0
        self.pg = None
        if self.world > 1 and dim % 8 == 0:
            # its own pair of symmetric buffers: sharing the eager path's (or another step's) would let two gathers in a
            # row target the same buffer while a slower rank still reads it
            from .peer import private_gather
            self.pg = private_gather(bs, dim, dev)
        if self.pg is None:
            self._pack_local = [torch.empty(bs, 2 * dim, dtype=torch.bfloat16, device=dev) for _ in range(2)]
            if self.world > 1:
                self._pack_all = [torch.empty(self.world * bs, 2 * dim, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        self._graphs = [None, None]
        self._capture()

    # the launches of one step on the current stream, static buffers only
    def _enqueue(self, s: int):
        ft, fc = self.feat_t_in[s], self.feat_cond_in[s]
        if self.world == 1 and ops.local_step_ok(ft):
            self._out[s] = ops.omc_step_local(ft, fc, self._temp, self.ls, self.wf, seed=self.seed, offset=0,
                                              need_sample=self.need_negatives, need_grad=True, buffers=self._out[s],
                                              step_counter=self._ctr)
            return
        if self.pg is not None:
            pack = self.pg.gather(ft, fc, slot=s)
        else:
            pack = ops.pack_pair(ft, fc, out=self._pack_local[s])
            if self.world > 1:
                dist.all_gather_into_tensor(self._pack_all[s], pack)
                pack = self._pack_all[s]
        self._out[s] = ops.omc_step(pack, self.bs, self.rank * self.bs, self._temp, self.ls, self.wf, seed=self.seed,
                                    offset=0, need_sample=self.need_negatives, need_grad=True, buffers=self._out[s],
                                    step_counter=self._ctr)

    def _capture(self):
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for s in range(2):            # eager warm-up: allocates outputs + workspace, loads the kernels
                self._enqueue(s)
            for s in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self._enqueue(s)
                self._graphs[s] = g
        cur.wait_stream(side)
        self._turn = 0
        if self.pg is not None:
            self.pg.turn = 0

    def inputs(self, slot: int | None = None):
        """(feat_cond, feat_t) static input buffers of the slot the NEXT call will use (or of `slot`): a loader
        may copy straight into them and pass them to the call, which then skips its device-to-device copy."""
        s = self._turn if slot is None else slot
        return self.feat_cond_in[s], self.feat_t_in[s]

    def input_block(self, slot: int | None = None) -> torch.Tensor:
        """The same buffers as one contiguous [2, bs, dim] tensor: [0] = feat_t, [1] = feat_cond."""
        return self._in[self._turn if slot is None else slot]

    @property
    def next_slot(self) -> int:
        return self._turn

    def __call__(self, feat_cond: torch.Tensor, feat_t: torch.Tensor):
        """(loss, neg_idx_cond2t, neg_idx_t2cond) like `omc_loss_and_negatives`; differentiable in feat_cond,
        feat_t and the temperature given at construction."""
        s = self._turn
        self._turn = s ^ 1
        if self.pg is not None:
            self.pg.turn = s ^ 1
        if self._temp.data_ptr() != self._temp_ptr:  # pragma: no cover
            raise RuntimeError("OmcGraphStep: the temperature tensor moved; build a new step")
        for src, dst in ((feat_cond, self.feat_cond_in[s]), (feat_t, self.feat_t_in[s])):
            if src.shape != dst.shape:
                raise RuntimeError(f"OmcGraphStep: expected features of shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src.detach(), non_blocking=True)
        return _GraphFn.apply(feat_cond, feat_t, self._temp_arg, self, s)
