"""2-D SPINN on the fused B200 path -- drop-in for the reference's ``spinn2d`` module.

Mirrors /root/reference/code/spinn2d.py: ``Plotter2D`` :11-94, ``Shift2D`` :97-135,
``SPINN2D`` :138-188, ``gaussian`` :191-192, ``SoftPlus`` :195-205, ``App2D`` :252-275
(same constructor arguments, parameter names/shapes -- ``layer1.x, layer1.y,
layer1.h, layer2.weight[K,M], layer2.bias[K]`` -- and CLI flags), but
``SPINN2D.forward`` evaluates u AND its spatial derivatives in one fused CUDA
kernel (spinn_b200.ops) instead of building (N, M) tensors for autograd.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .common import NO_LIVE_PLOT, App, Plotter, add_flags, tensor


def gaussian(x, y):
    """exp(-(x^2+y^2)/2) (reference spinn2d.py:191-192).  Passed as ``activation``
    it selects the fused Gaussian kernels; callable for completeness."""
    return torch.exp(-0.5 * (x * x + y * y))


class SoftPlus:
    """Softplus 'hat' kernel (reference spinn2d.py:195-205); an instance passed as
    ``activation`` selects the fused softplus-hat kernels."""

    def __init__(self):
        self._sp = nn.Softplus()
        self.k = tensor([1.0 + 4.0 * np.log(2.0)])
        self.fac = self._sp(tensor([1.0]))

    def __call__(self, x, y):
        sp = self._sp
        return sp(self.k - sp(2.0 * x) - sp(-2.0 * x) - sp(2.0 * y) - sp(-2.0 * y)) / self.fac


def kernel_kind(activation, gaussian_fn=gaussian, softplus_cls=SoftPlus):
    """Map the ``activation`` object chosen by ``App._get_activation`` to a fused
    kernel id.  Anything else (learned MLP kernels, reference spinn2d.py:208-249)
    is not on the fused path and is rejected loudly -- there is no fallback."""
    if activation is gaussian_fn or getattr(activation, "__name__", None) == "gaussian":
        return ops.GAUSSIAN
    if isinstance(activation, softplus_cls) or type(activation).__name__ == "SoftPlus":
        return ops.SOFTPLUS
    raise NotImplementedError(
        f"activation {activation!r} is not supported by the fused B200 path "
        "(supported: gaussian, SoftPlus); use the stock reference network for learned kernels")


class Shift2D(nn.Module):
    """Node centres and widths (reference spinn2d.py:97-135): free centres are
    parameters, fixed centres plain tensors, h = ptp(x)/sqrt(n) (or one shared
    scalar with ``fixed_h``)."""

    def __init__(self, points, fixed_points, fixed_h=False):
        super().__init__()
        px, py = np.asarray(points[0], dtype=float), np.asarray(points[1], dtype=float)
        fx, fy = np.asarray(fixed_points[0], dtype=float), np.asarray(fixed_points[1], dtype=float)
        self.n = n = px.size + fx.size
        span = np.ptp(np.hstack((px, fx))) if fx.size else np.ptp(px)
        self.dx = dx = span / np.sqrt(n)
        self.x = nn.Parameter(tensor(px))
        self.y = nn.Parameter(tensor(py))
        self.xf = tensor(fx)
        self.yf = tensor(fy)
        self.fixed_h = fixed_h
        self.h = nn.Parameter(tensor(dx) if fixed_h else dx * torch.ones(n, device=self.xf.device))

    def _apply(self, fn, *a, **k):
        # unlike the reference, moving the module also moves the fixed centres
        super()._apply(fn, *a, **k)
        self.xf, self.yf = fn(self.xf), fn(self.yf)
        return self

    def centers(self):
        return torch.cat((self.x, self.xf)), torch.cat((self.y, self.yf))

    def widths(self):
        if self.fixed_h:
            return self.h * torch.ones(self.n, device=self.h.device)
        return self.h

    def forward(self, x, y):
        xc, yc = self.centers()
        fac = 1.0 / self.h
        return (x - xc) * fac, (y - yc) * fac


class SPINN2D(nn.Module):
    FLAGS = (
        (("--fixed-h",), "fixed_h", False, None, "Use fixed width nodes."),
        (("--pu",), "pu", False, None, "Use a partition of unity."),
        # not in the reference: compact-support evaluation (Gaussian kernel, n_vars=1)
        (("--cutoff",), "cutoff", 0.0, float,
         "Skip node blocks farther than CUTOFF widths from a group of points (0 = dense)."),
    )
    #: refresh the Morton visiting order of the nodes every this many fused calls (it only
    #: affects how much is skipped, never the result)
    perm_refresh = 64
    #: also produce u_xy so that ``ag.grad(u_x, (x, y))`` is complete (the reference
    #: computes it and drops it, pde2d_base.py:52-62); costs 2 of 16 FP32 ops per pair.  Problem
    #: classes whose derivative recipe provably drops the mixed term say so with the class
    #: attribute ``needs_mixed_derivative = False`` (spinn_b200.problems.*); anything else --
    #: e.g. the reference's own classes through dropin/ -- gets the complete tower.
    mixed_derivative = True

    @classmethod
    def from_args(cls, pde, activation, args):
        return cls(pde, activation, fixed_h=args.fixed_h, use_pu=args.pu,
                   cutoff=getattr(args, "cutoff", 0.0))

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, {})

    def __init__(self, pde, activation, fixed_h=False, use_pu=False, cutoff=0.0):
        super().__init__()
        self.cutoff = float(cutoff) if cutoff else None
        self._perm, self._perm_age = None, 0
        self.activation = activation
        self.kind = kernel_kind(activation)
        self.use_pu = use_pu
        # u_xy is skipped only if the problem class says so AND still uses the packaged derivative
        # recipe the statement was made for: a subclass that overrides `_compute_derivatives`
        # (e.g. to use u_xy) gets the complete tower again instead of a silently zero mixed term
        recipe = getattr(type(pde), "_compute_derivatives", None)
        vouched = getattr(recipe, "drops_mixed_derivative", False)
        self.mixed_derivative = not (getattr(pde, "needs_mixed_derivative", True) is False and vouched)
        self.layer1 = Shift2D(pde.nodes(), pde.fixed_nodes(), fixed_h=fixed_h)
        self.layer2 = nn.Linear(self.layer1.n, pde.n_vars(), bias=not use_pu)
        self.layer2.weight.data.fill_(0.0)
        if self.layer2.bias is not None:
            self.layer2.bias.data.fill_(0.0)
        if pde.n_vars() + int(use_pu) > ops.MAX_K:
            raise NotImplementedError(f"n_vars={pde.n_vars()} (+1 with --pu) > {ops.MAX_K} not supported")

    # -- fused evaluation ------------------------------------------------------
    def jets(self, x, y, level=2):
        """Raw fused call: tuple of (N, K) tensors u, ux, uy, uxx, uyy[, uxy]."""
        l1, l2 = self.layer1, self.layer2
        if not self.use_pu:
            cutoff = self.cutoff if ops.sparse_supported(2, self.kind, level, l2.weight.shape[0]) else None
            return ops.fused_jets(2, self.kind, level, x.detach().contiguous(), y.detach().contiguous(),
                                  l1.x, l1.y, l1.xf, l1.yf, l1.h, l2.weight, l2.bias,
                                  cutoff, self.node_order() if cutoff else None)
        # partition of unity (reference spinn2d.py:174-178): u = sum_j W_j phi_j / sum_j phi_j.
        # One fused call with an extra all-ones output row gives T_a = sum W phi_a and
        # S_a = sum phi_a; the quotient rule on the jets is O(N) differentiable torch code, so
        # autograd carries the upstream gradients back to the same fused backward.
        W = torch.cat((l2.weight, torch.ones_like(l2.weight[:1])), 0)
        J = ops.fused_jets(2, self.kind, level, x.detach().contiguous(), y.detach().contiguous(),
                           l1.x, l1.y, l1.xf, l1.yf, l1.h, W, None)
        return ops.quotient_jets_2d([j[:, :-1] for j in J], [j[:, -1:] for j in J])

    def node_order(self):
        """Cached Morton visiting order of the nodes for the compact-support kernels."""
        if self._perm is None or self._perm_age >= self.perm_refresh or self._perm.device != self.layer1.h.device:
            xc, yc = self.layer1.centers()
            self._perm, self._perm_age = ops.morton_order(xc, yc), 0
        self._perm_age += 1
        return self._perm

    def forward(self, x, y):
        x = x.reshape(-1)
        y = y.reshape(-1)
        spatial = torch.is_grad_enabled() and (x.requires_grad or y.requires_grad)
        if not spatial:
            (u,) = self.jets(x, y, level=0)
            return u.squeeze()
        if self.mixed_derivative:
            u, ux, uy, uxx, uyy, uxy = self.jets(x, y, level=3)
        else:
            u, ux, uy, uxx, uyy = self.jets(x, y, level=2)
            uxy = None
        return ops.attach_tower_2d(x, y, u, ux, uy, uxx, uyy, uxy).squeeze()

    def centers(self):
        return self.layer1.centers()

    def widths(self):
        return self.layer1.widths()

    def weights(self):
        return self.layer2.weight


class Plotter2D(Plotter):
    """Error norms and result files (reference spinn2d.py:11-43); no live plotting."""

    def get_error(self, xn=None, yn=None, pn=None):
        if not self.pde.has_exact():
            return 0.0, 0.0, 0.0
        if xn is None and pn is None:
            xn, yn, pn = self.get_plot_data()
        diff = self.pde.exact(xn, yn) - pn
        return np.mean(np.abs(diff)), np.sqrt(np.mean(diff ** 2)), np.max(np.abs(diff))

    def save(self, dirname):
        torch.save(self.nn.state_dict(), os.path.join(dirname, "model.pt"))
        x, y, u = self.get_plot_data()
        fields = dict(x=x, y=y, u=u)
        exact = self.pde.exact(x, y) if self.pde.has_exact() else None
        if exact is not None:                    # no pickled `None` entry for problems without an exact solution
            fields["u_exact"] = exact
        np.savez(os.path.join(dirname, "results.npz"), **fields)

    def field_on_grid(self, x, t):
        """u evaluated on a NumPy grid (level-0 call of the fused forward), shaped like the grid."""
        u = self.nn(tensor(x.ravel()), tensor(t.ravel())).detach().cpu().numpy()
        return u.reshape(x.shape)

    def get_plot_data(self):
        x, y = self.pde.plot_points()
        pn = self.nn(tensor(x.ravel()), tensor(y.ravel())).detach().cpu().numpy()
        pn.shape = x.shape
        return x, y, pn

    # live plotting is not provided (the reference's mayavi code, spinn2d.py:45-94, is visualisation)
    def plot(self):
        raise NotImplementedError(NO_LIVE_PLOT)

    plot_solution = plot_weights = plot


class App2D(App):
    ACTIVATION_FLAGS = (
        (("--activation", "-a"), "activation", "gaussian", ("gaussian", "softplus", "kernel", "nnkernel"),
         "Select the activation function for particles."),
        (("--kernel-size",), "kernel_size", 10, int, "Activation kernel size (in place of a Gaussian)."),
    )

    def _setup_activation_options(self, p, **kw):
        add_flags(p, self.ACTIVATION_FLAGS, kw)

    def _get_activation(self, args):
        if args.activation == "gaussian":
            return gaussian
        if args.activation == "softplus":
            return SoftPlus()
        raise NotImplementedError(
            f"--activation {args.activation}: learned kernels are outside the fused B200 path")
