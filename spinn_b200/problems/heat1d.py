"""1-D heat equation as a space-time problem on IBVP2D (mirrors /root/reference/code/heat1d.py):
u_t = c^2 u_xx, exact solution b1 exp(-a^2 t) sin(pi x / L).

    python -m spinn_b200.problems.heat1d --gpu
"""
import os

import numpy as np
import torch

from ..common import tensor
from ..spinn2d import App2D, Plotter2D, SPINN2D
from .ibvp2d_base import IBVP2D


class HeatPlotter(Plotter2D):
    """results.npz of the reference's HeatPlotter (heat1d.py:13-29): the field and the exact solution
    on a fixed 100 x 21 grid of (x, t) in [0,1] x [0,0.2] (keys x, t, u, u_exact) next to the plot-grid
    data (xp, yp, up, uex_p); model.pt as usual."""

    def save(self, dirname):
        torch.save(self.nn.state_dict(), os.path.join(dirname, "model.pt"))
        xp, yp, up = self.get_plot_data()
        x, t = np.mgrid[0:1:100j, 0:0.2:21j]
        np.savez(os.path.join(dirname, "results.npz"), x=x, t=t, u=self.field_on_grid(x, t),
                 u_exact=self.pde.exact(x, t), xp=xp, yp=yp, up=up, uex_p=self.pde.exact(xp, yp))


class Heat1D(IBVP2D):
    #: the boundary penalty is a plain sum over boundary points -> they can be sharded across ranks
    boundary_reduction = "sum"

    C = 1.0
    B1 = 2.0

    def pde(self, x, t, u, ux, ut, uxx, utt):
        return ut - self.C * self.C * uxx

    def has_exact(self):
        return True

    def exact(self, x, t):
        L = self.xR - self.xL
        a2 = (np.pi * self.C / L) ** 2
        return self.B1 * np.exp(-a2 * t) * np.sin(np.pi * x / L)

    def boundary_loss(self, nn):
        xb, tb = self.boundary()
        ub = getattr(self, "_ub", None)          # constant target, built once (reference: every step)
        if ub is None or ub.device != xb.device:
            ub = self._ub = tensor(self.exact(xb.detach().cpu().numpy(), tb.detach().cpu().numpy()))
        bc = nn(xb, tb) - ub
        return (bc ** 2).sum()


DEFAULTS = dict(nodes=50, samples=200, lr=1e-2)


def make_app():
    return App2D(pde_cls=Heat1D, nn_cls=SPINN2D, plotter_cls=HeatPlotter)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
