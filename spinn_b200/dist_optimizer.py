"""Data-parallel version of the training loop: same `Optimizer` (reference common.py:154-314),
but every rank evaluates its slice of the interior points (distributed.shard_interior) and the
parameter gradients and the loss travel in ONE flat buffer and ONE all-reduce per closure
call (the parameters' .grad are views into it) before the (replicated) optimizer update.  Launch with torchrun, one process per GPU:

    python -m torch.distributed.run --nproc-per-node 8 -m spinn_b200.problems.poisson2d_sine --gpu ...

`App.run` picks this class automatically when WORLD_SIZE > 1.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import distributed
from .common import Optimizer


class DistributedOptimizer(Optimizer):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.rank, self.world, _ = distributed.init()
        distributed.shard_interior(self.pde, self.rank, self.world)
        self.boundary_sharded = distributed.shard_boundary(self.pde, self.rank, self.world)
        self._flat = None

    def _local_loss(self):
        """This rank's share of the global loss, such that a SUM over ranks is the unsharded
        loss: interior means are divided by the world size (equal shards), interior sums are
        kept (`interior_reduction = "sum"` on the problem class, e.g. the cavity); the boundary
        term is a plain local sum when the boundary points are sharded too (problem classes that
        declare `boundary_reduction = "sum"`), otherwise every rank evaluates all of it and it is
        divided by the world size."""
        from .common import PDE
        pde, w = self.pde, float(self.world)
        if type(pde).loss is not PDE.loss:          # custom total loss: treat as a mean
            return pde.loss(self.nn) / w
        li = pde.interior_loss(self.nn)
        lb = pde.boundary_loss(self.nn)
        if not self.boundary_sharded:
            lb = lb / w
        if getattr(pde, "interior_reduction", "mean") != "sum":
            li = li / w
        return li + lb

    def bind_gradients(self):
        """Make every parameter's .grad a view into ONE flat buffer [grads | loss]: autograd
        accumulates into the views in place, so the buffer IS the message of the step's single
        all-reduce (no pack / unpack copies)."""
        params = [p for p in self.nn.parameters() if p.requires_grad]
        bound = self._flat is not None and all(
            p.grad is not None and p.grad.untyped_storage().data_ptr() == self._flat.untyped_storage().data_ptr()
            for p in params)
        if bound:
            return self._flat
        n = sum(p.numel() for p in params)
        self._flat = torch.zeros(n + 1, dtype=params[0].dtype, device=params[0].device)
        off = 0
        for p in params:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        return self._flat

    def reduce_step(self, loss):
        """loss -> last slot; ONE collective for gradients + loss; returns the global loss (a view)."""
        flat = self._flat
        flat[-1:].copy_(loss.detach().reshape(1))
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return flat[-1]

    def closure(self):
        self.bind_gradients().zero_()               # instead of opt.zero_grad(): keeps the views
        loss = self._local_loss()
        loss.backward(retain_graph=True)
        loss = self.reduce_step(loss).clone()
        self.loss.append(loss.item())
        return loss

    def _report(self, i, loss):
        if self.rank == 0:
            return super()._report(i, loss)
        # other ranks follow the same control flow (tolerance stop) without printing
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            return super()._report(i, loss)
