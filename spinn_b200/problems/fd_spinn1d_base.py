"""Hybrid finite-difference / SPINN time stepping in 1-D (mirrors
/root/reference/code/fd_spinn1d_base.py: `FDPlotter1D` :13-73, `FDSPINN1D` :76-139,
`AppFD1D` :142-190).

Implicit Euler in time, SPINN1D in space: every time level trains the network on the residual
`pde(x, u, ux, uxx, u0, dt)` where `u0` is the network's own output at the previous level, so one
run is `T/dt` consecutive `solver.solve()` calls on the same network, the same optimiser state and
the 1-D fused jets op.  With `--graph` the captured training step is re-used across the time levels
(`u0` lives in one buffer that is overwritten in place).
"""
import os

import numpy as np
import torch

from ..common import add_flags, device
from ..spinn1d import App1D, Plotter1D
from .ode_base import BasicODE


class FDPlotter1D(Plotter1D):
    """Errors are taken against the exact solution at the END of the current time level, i.e. at
    `t + dt` (fd_spinn1d_base.py:17-31); one model/result file per saved level (:62-73)."""

    def show(self):
        return None

    def _exact_at_next_level(self, xn):
        pde = self.pde
        t = pde.t
        pde.t = t + pde.dt
        try:
            return pde.exact(xn)
        finally:
            pde.t = t

    def get_error(self, xn=None, pn=None):
        if not self.pde.has_exact():
            return 0.0, 0.0, 0.0
        if xn is None and pn is None:
            xn, pn = self.get_plot_data()
        diff = self._exact_at_next_level(xn) - pn
        return np.mean(np.abs(diff)), np.sqrt(np.mean(diff ** 2)), np.max(np.abs(diff))

    def save(self, dirname):
        it, t = self.pde.iter, self.pde.t
        torch.save(self.nn.state_dict(), os.path.join(dirname, "model_%04d.pt" % it))
        x, y = self.get_plot_data()
        np.savez(os.path.join(dirname, "results_%04d.npz" % it), x=x, y=y,
                 y_exact=self.pde.exact(x), t=t, iter=it)


class FDSPINN1D(BasicODE):
    FD_FLAGS = (
        (("--dt",), "dt", 1e-3, float, "Time step for finite difference scheme."),
        (("--duration", "-T"), "T", 1.0, float, "Duration of simulation"),
        (("--t-skip",), "t_skip", 10, int, "Iterations"),
    )

    @classmethod
    def from_args(cls, args):
        return cls(args.nodes, args.samples, args.sample_frac, args.dt, args.T, args.t_skip)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        super().setup_argparse(parser, **kw)
        add_flags(parser, cls.FD_FLAGS, kw)

    #: homogeneous Dirichlet penalty used by all three reference drivers: w * sum(u(boundary)^2)
    boundary_penalty = 100
    #: number of plot/error points as a function of the sample count
    n_plot = staticmethod(lambda ns: min(2 * ns, 500))

    def __init__(self, n, ns, sample_frac=1.0, dt=1e-3, T=1.0, t_skip=10):
        super().__init__(n, ns, sample_frac)          # geometry on `domain` (ode_base.BasicODE)
        self.iter, self.t = 0, 0.0
        self.dt, self.T, self.t_skip = dt, T, t_skip
        self.u0 = self.initial(self.xs)

    def boundary_loss(self, nn):
        return self.boundary_penalty * (nn(self.boundary()) ** 2).sum()

    def plot_points(self):
        lo, hi = self.domain
        return np.linspace(lo, hi, self.n_plot(self.ns))

    def initial(self, x):
        return torch.zeros_like(x).to(device())

    def pde(self, x, u, ux, uxx, u0, dt):
        raise NotImplementedError()

    def on_shard(self):
        """Data-parallel runs slice `xs` (distributed.shard_interior): restart the previous-level
        field on the local slice."""
        self.u0 = self.initial(self.xs)

    def sample_arrays(self, *arrays):
        return self.pick(*arrays)                # the same draw for the points and for u0

    def interior(self):
        return self.xs

    def interior_loss(self, nn):
        xs, u0 = self.sample_arrays(self.interior(), self.u0)
        u = nn(xs)
        u, ux, uxx = self._compute_derivatives(u, xs)
        res = self.pde(xs, u, ux, uxx, u0, self.dt)
        return (res ** 2).mean()

    # -- CUDA-graph support (spinn_b200/graphed.py); not in the reference ------------------
    def enable_static_sampling(self):
        super().enable_static_sampling()
        if not getattr(self, "_u0_static", False):
            # one buffer for the previous-level field, detached from whatever produced it
            self.u0 = self.u0.detach().clone()
            self._u0_static = True

    def advance(self, nn):
        """End of a time level (fd_spinn1d_base.py:178-182): t += dt and u0 <- nn(xs).  The field is
        overwritten in place so that a captured training step keeps reading the right buffer."""
        self.t += self.dt
        self.iter += 1
        with torch.no_grad():
            new = nn(self.xs)
            if self.u0.requires_grad or self.u0.shape != new.shape:
                self.u0 = new.clone()
            else:
                self.u0.copy_(new)


class AppFD1D(App1D):
    def run(self, args=None, **kw):
        args = self.setup_argparse(**kw).parse_args(args)
        if args.gpu and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")
        else:
            device("cuda" if args.gpu else "cpu")
        activation = self._get_activation(args)
        self.pde = self.pde_cls.from_args(args)
        self.nn = self.nn_cls.from_args(self.pde, activation, args).to(device())
        self.plotter = self.plotter_cls.from_args(self.pde, self.nn, args)
        self.solver = self.optimizer.from_args(self.pde, self.nn, self.plotter, args)
        self.solver.use_loss_for_tolerance()
        self.iterate()

    def n_levels(self):
        return round(self.pde.T / self.pde.dt + 0.49)           # fd_spinn1d_base.py:164

    def iterate(self):
        out_dir = self.solver.out_dir
        if out_dir is None:
            print("No output directory set.  Not saving.")
        else:
            if not os.path.exists(out_dir):
                os.makedirs(out_dir)
                print("Saving output to", out_dir)
            self.plotter.save(out_dir)
        for level in range(self.n_levels()):
            self.solver.solve()
            self.pde.advance(self.nn)
            print("Current time: %f" % self.pde.t)
            if (level + 1) % self.pde.t_skip == 0 and out_dir is not None:
                self.plotter.save(out_dir)
