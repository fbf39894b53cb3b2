"""1-D (viscous) Burgers equation as a space-time problem on IBVP2D (mirrors
/root/reference/code/burgers1d.py): u_t + u u_x - nu u_xx = 0, residual uses u, u_x, u_t, u_xx.

    python -m spinn_b200.problems.burgers1d --gpu
"""
import os

import numpy as np
import torch

from ..common import add_flags, tensor
from ..spinn2d import App2D, Plotter2D, SPINN2D
from .ibvp2d_base import IBVP2D


def initial_profile(ic, x, t, a=0.0):
    """Boundary/initial data (reference burgers1d.py:57-74)."""
    if ic == "jump":
        z = x - a * t + 1.1
        return np.heaviside(z, 0.5) - 2 * np.heaviside(z - 1.1, 0.5)
    if ic == "gaussian":
        z = (x - a * t + 0.3) / 0.15
        return np.exp(-0.5 * z ** 2)
    shift, freq = (0.75, 1.0) if ic == "sin" else (0.5, 2.0)
    z = x - a * t + shift
    return (np.heaviside(z, 0.5) - np.heaviside(z - 1.0, 0.5)) * np.sin(z * np.pi * freq)


class MyPlotter(Plotter2D):
    """results.npz of the reference's burgers plotter (burgers1d.py:46-60): the field on a fixed
    500 x 11 grid of (x, t) in [-1,1] x [0,1] (keys x, t, u) next to the plot-grid data (xp, tp, up);
    there is no exact solution, so no u_exact entry."""

    def save(self, dirname):
        torch.save(self.nn.state_dict(), os.path.join(dirname, "model.pt"))
        x, t = np.mgrid[-1:1:500j, 0:1:11j]
        xp, tp, up = self.get_plot_data()
        np.savez(os.path.join(dirname, "results.npz"), x=x, t=t, u=self.field_on_grid(x, t), xp=xp, tp=tp, up=up)


class Burgers1D(IBVP2D):
    #: the boundary penalty is a plain sum over boundary points -> they can be sharded across ranks
    boundary_reduction = "sum"

    EXTRA_FLAGS = (
        (("--viscosity",), "viscosity", 0.0, float, "Viscosity in viscous Burgers equation."),
        (("--ic",), "ic", "gaussian", ("jump", "gaussian", "sin", "sin2"),
         "Initial profiles for linear advection equation."),
    )

    @classmethod
    def from_args(cls, args):
        return cls(args.nodes, args.samples, args.b_nodes, args.b_samples, args.xL, args.xR, args.T,
                   args.ic, args.viscosity, args.sample_frac)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        super().setup_argparse(parser, **kw)
        add_flags(parser, cls.EXTRA_FLAGS, kw)

    def __init__(self, n, ns, nb=None, nbs=None, xL=0.0, xR=1.0, T=1.0, ic="gaussian", viscosity=0.0,
                 sample_frac=1.0):
        super().__init__(n, ns, nb, nbs, xL, xR, T, sample_frac)
        self.ic, self.viscosity = ic, viscosity

    def pde(self, x, t, u, ux, ut, uxx, utt):
        return ut + u * ux - self.viscosity * uxx

    def has_exact(self):
        return False

    def boundary_loss(self, nn):
        xb, tb = self.boundary()
        ub = getattr(self, "_ub", None)
        if ub is None or ub.device != xb.device:
            ub = self._ub = tensor(initial_profile(self.ic, xb.detach().cpu().numpy(), tb.detach().cpu().numpy()))
        bc = nn(xb, tb) - ub
        return (bc ** 2).sum()


DEFAULTS = dict(nodes=200, samples=1000, xL=-1.0, xR=1.0, lr=1e-3)


def make_app():
    return App2D(pde_cls=Burgers1D, nn_cls=SPINN2D, plotter_cls=MyPlotter)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
