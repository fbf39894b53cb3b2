"""1-D SPINN on the fused B200 path -- drop-in for the reference's ``spinn1d`` module.

Mirrors /root/reference/code/spinn1d.py: ``Plotter1D`` :11-85, ``Shift`` :88-112,
``SPINN1D`` :115-171, ``gaussian`` :176-177, ``SoftPlus`` :180-189, ``App1D`` :214-235.
Parameter names: ``layer1.center``, ``layer1.h``, ``layer2.weight``, ``layer2.bias``.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .common import NO_LIVE_PLOT, App, Plotter, add_flags, tensor


def gaussian(x):
    return torch.exp(-0.5 * x * x)


class SoftPlus:
    def __init__(self):
        self._sp = nn.Softplus()
        self.k = tensor([1.0 + 2.0 * np.log(2.0)])
        self.fac = self._sp(tensor([1.0]))

    def __call__(self, x):
        sp = self._sp
        return sp(self.k - sp(2.0 * x) - sp(-2.0 * x)) / self.fac


def kernel_kind(activation):
    if activation is gaussian or getattr(activation, "__name__", None) == "gaussian":
        return ops.GAUSSIAN
    if type(activation).__name__ == "SoftPlus":
        return ops.SOFTPLUS
    raise NotImplementedError(
        f"activation {activation!r} is not supported by the fused B200 path (gaussian, SoftPlus only)")


class Shift(nn.Module):
    """Centres and widths in 1-D (reference spinn1d.py:88-112): h=(xmax-xmin)/n."""

    def __init__(self, points, fixed_points, fixed_h=False):
        super().__init__()
        pts = np.asarray(points, dtype=float).ravel()
        fix = np.asarray(fixed_points, dtype=float).ravel()
        self.n = n = pts.size + fix.size
        allx = np.hstack((pts, fix)) if fix.size else pts
        dx = (allx.max() - allx.min()) / n
        self.center = nn.Parameter(tensor(pts))
        self.fixed = tensor(fix)
        self.fixed_h = fixed_h
        self.h = nn.Parameter(tensor(dx) if fixed_h else dx * torch.ones(n, device=self.fixed.device))

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self.fixed = fn(self.fixed)
        return self

    def centers(self):
        return torch.cat((self.center, self.fixed))

    def forward(self, x):
        return (x - self.centers()) / self.h


class SPINN1D(nn.Module):
    FLAGS = (
        (("--fixed-h",), "fixed_h", False, None, "Use fixed width nodes."),
        (("--pu",), "pu", False, None, "Use a partition of unity."),
    )

    @classmethod
    def from_args(cls, pde, activation, args):
        return cls(pde, activation, fixed_h=args.fixed_h, use_pu=args.pu)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, {})

    def __init__(self, pde, activation, fixed_h=False, use_pu=False):
        super().__init__()
        self.fixed_h, self.use_pu = fixed_h, use_pu
        self.activation = activation
        self.kind = kernel_kind(activation)
        if pde.n_vars() + int(use_pu) > ops.MAX_K:
            raise NotImplementedError(f"n_vars={pde.n_vars()} (+1 with --pu) > {ops.MAX_K} not supported")
        self.layer1 = Shift(pde.nodes(), pde.fixed_nodes(), fixed_h=fixed_h)
        self.layer2 = nn.Linear(self.layer1.n, pde.n_vars(), bias=not use_pu)
        self.layer2.weight.data.fill_(0.0)
        if self.layer2.bias is not None:
            self.layer2.bias.data.fill_(0.0)

    def jets(self, x, level=2):
        l1, l2 = self.layer1, self.layer2
        if not self.use_pu:
            return ops.fused_jets(1, self.kind, level, x.detach().contiguous(), None,
                                  l1.center, None, l1.fixed, None, l1.h, l2.weight, l2.bias)
        # partition of unity (reference spinn1d.py:151-155): extra all-ones output row, then the
        # quotient rule on the jets (see spinn2d.SPINN2D.jets)
        W = torch.cat((l2.weight, torch.ones_like(l2.weight[:1])), 0)
        J = ops.fused_jets(1, self.kind, level, x.detach().contiguous(), None,
                           l1.center, None, l1.fixed, None, l1.h, W, None)
        return ops.quotient_jets_1d([j[:, :-1] for j in J], [j[:, -1:] for j in J])

    def forward(self, x):
        x = x.reshape(-1)
        if not (torch.is_grad_enabled() and x.requires_grad):
            (u,) = self.jets(x, level=0)
            return u.squeeze()
        u, ux, uxx = self.jets(x, level=2)
        return ops.attach_tower_1d(x, u, ux, uxx).squeeze()

    def centers(self):
        return self.layer1.centers()

    def widths(self):
        return self.layer1.h

    def weights(self):
        return self.layer2.weight

    def show(self):
        print("Basis centers: ", self.centers())
        print("Mesh widths: ", self.widths())
        print("Nodal weights: ", self.weights())
        print("Output bias: ", self.layer2.bias)


class Plotter1D(Plotter):
    def get_error(self, xn=None, pn=None):
        if not self.pde.has_exact():
            return 0.0, 0.0, 0.0
        if xn is None and pn is None:
            xn, pn = self.get_plot_data()
        diff = self.pde.exact(xn) - pn
        return np.mean(np.abs(diff)), np.sqrt(np.mean(diff ** 2)), np.max(np.abs(diff))

    def save(self, dirname):
        torch.save(self.nn.state_dict(), os.path.join(dirname, "model.pt"))
        x, y = self.get_plot_data()
        np.savez(os.path.join(dirname, "results.npz"), x=x, y=y, y_exact=self.pde.exact(x))

    def get_plot_data(self):
        x = self.pde.plot_points()
        return x, self.nn(tensor(x)).detach().cpu().numpy()

    # live plotting is not provided (the reference's matplotlib code, spinn1d.py:44-83, is visualisation)
    def plot(self):
        raise NotImplementedError(NO_LIVE_PLOT)

    plot_solution = plot_weights = plot


class App1D(App):
    ACTIVATION_FLAGS = (
        (("--activation", "-a"), "activation", "gaussian", ("gaussian", "softplus", "tanh", "kernel"),
         "Select the activation function for particles."),
        (("--kernel-size",), "kernel_size", 5, int, "Activation kernel size (in place of a Gaussian)."),
    )

    def _setup_activation_options(self, p, **kw):
        add_flags(p, self.ACTIVATION_FLAGS, kw)

    def _get_activation(self, args):
        if args.activation == "gaussian":
            return gaussian
        if args.activation == "softplus":
            return SoftPlus()
        raise NotImplementedError(
            f"--activation {args.activation}: outside the fused B200 path (gaussian, softplus only)")
