"""CUDA-graphed training step ("fused train step & sync removal").

The small BASELINE configurations (ode3, poisson2d_sine, irregular domain, cavity) are
launch- and sync-bound, not math-bound: one `opt.step(closure)` is ~100 tiny kernels plus a
`loss.item()` round trip (reference common.py:242-248).  `GraphedOptimizer` runs the same
loop as `common.Optimizer` but captures ONE step -- loss, backward (incl. the fused SPINN
backward kernels, enqueued through the C ABI on the capturing stream) and the Adam update
(`capturable=True`) -- into a CUDA graph and replays it; per-step losses go to a device
buffer and are read back only when the loop reports (every `n_skip`).

Requirements on the problem class (met by spinn_b200.problems.*): no host round trip inside
`loss()` (boundary targets are cached on the device) and random sub-sampling through a static
index buffer (`enable_static_sampling` / `graph_pre_step`).  Select with `--graph`.
"""
from __future__ import annotations

import time

import torch

from . import distributed
from .dist_optimizer import DistributedOptimizer


class GraphedOptimizer(DistributedOptimizer):
    """Single- or multi-GPU (torchrun): with WORLD_SIZE > 1 the interior points are sharded and the
    packed gradient all-reduce (NCCL) is captured inside the graph as well."""

    WARMUP_STEPS = 3          # eager steps before capture (they are real training steps)
    FUSED_ADAM = True

    def _body(self):
        loss = self._local_loss()
        # gradients for the network parameters only: the reference also accumulates (never-read,
        # never-zeroed) .grad on the sample tensors (pde2d_base.py:91); skipping that saves kernels
        # and keeps those buffers from growing without bound over long graphed runs
        loss.backward(retain_graph=True, inputs=self._params)
        # gradients accumulate in place into views of one flat buffer whose last slot takes the loss:
        # ONE all-reduce per step (captured inside the graph when WORLD_SIZE > 1)
        return self.reduce_step(loss)

    def _pre_step(self):
        hook = getattr(self.pde, "graph_pre_step", None)
        if hook is not None:
            hook()

    def solve(self):
        if self.opt_class is not torch.optim.Adam:
            raise NotImplementedError("--graph supports the Adam optimizer only")
        params = self._params = list(self.nn.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("--graph needs --gpu")
        if self.opt is None:
            # one multi-tensor kernel for the whole update (same arithmetic as the default
            # implementation; SURVEY.md 8f-1 "fused Adam")
            self.opt = torch.optim.Adam(params, lr=self.lr, capturable=True, fused=self.FUSED_ADAM)
        if hasattr(self.pde, "enable_static_sampling"):
            self.pde.enable_static_sampling()
        n_train = self.n_train
        loss_buf = torch.zeros(max(n_train, 1), dtype=torch.float32, device=dev)
        done_before = len(self.loss)
        t0 = time.perf_counter()

        # ---- eager warm-up on a side stream (PyTorch's capture recipe) ----
        # A later solve() on the same optimiser (FD-SPINN: one per time level) replays the step
        # captured by the first one: same parameters, same Adam state, same static buffers.
        graph, static_loss = getattr(self, "_captured", (None, None))
        n_eager = min(self.WARMUP_STEPS, n_train) if graph is None else 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(n_eager):
                self._pre_step()
                self.bind_gradients().zero_()
                loss = self._body()
                self.opt.step()
                loss_buf[i].copy_(loss.detach())
        torch.cuda.current_stream().wait_stream(side)

        if graph is None and n_train > n_eager:
            self._pre_step_capture = True
            flat = self.bind_gradients()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                flat.zero_()
                static_loss = self._body()
                self.opt.step()
            self._captured = (graph, static_loss)

        err, stop, i = 1.0, False, 0
        for i in range(1, n_train + 1):
            if i > n_eager:
                self._pre_step()
                graph.replay()
                loss_buf[i - 1].copy_(static_loss.detach())
            if err < self.tol:
                stop = True
            if i % self.n_skip == 0 or i == n_train or stop:
                self.loss = self.loss[:done_before] + loss_buf[:i].tolist()      # one sync per report
                err = self._report(i, loss_buf[i - 1])
            if stop:
                break
        self.loss = self.loss[:done_before] + loss_buf[:i].tolist()
        if hasattr(self.pde, "disable_static_sampling"):
            self.pde.disable_static_sampling()
        self.time_taken = time.perf_counter() - t0
        print(f"Done. Took {self.time_taken:.3f} seconds.")
        if self.plot:
            self.plotter.show()
