"""Data-parallel version of the training loop: same `Optimizer` (reference common.py:154-314),
but every rank evaluates its slice of the interior points (distributed.shard_interior) and the
parameter gradients (+ the loss, for reporting) are combined with ONE all-reduce per closure
call before the (replicated) optimizer update.  Launch with torchrun, one process per GPU:

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
        self._buf = None

    def _local_loss(self):
        """This rank's share of the global loss, such that a SUM over ranks is the unsharded
        loss: interior means are divided by the world size (equal shards), interior sums are
        kept (`interior_reduction = "sum"` on the problem class, e.g. the cavity), and the
        boundary term -- evaluated in full by every rank -- is divided by the world size."""
        from .common import PDE
        pde, w = self.pde, float(self.world)
        if type(pde).loss is not PDE.loss:          # custom total loss: treat as a mean
            return pde.loss(self.nn) / w
        li = pde.interior_loss(self.nn)
        lb = pde.boundary_loss(self.nn)
        if getattr(pde, "interior_reduction", "mean") == "sum":
            return li + lb / w
        return (li + lb) / w

    def closure(self):
        self.opt.zero_grad()
        loss = self._local_loss()
        loss.backward(retain_graph=True)
        params = list(self.nn.parameters())
        self._buf = distributed.all_reduce_gradients(params, average=False, buf=self._buf)
        if self.world > 1:
            loss = loss.detach().clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
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
