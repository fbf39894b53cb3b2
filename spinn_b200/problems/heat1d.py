"""1-D heat equation as a space-time problem on IBVP2D (mirrors /root/reference/code/heat1d.py):
u_t = c^2 u_xx, exact solution b1 exp(-a^2 t) sin(pi x / L).

    python -m spinn_b200.problems.heat1d --gpu
"""
import numpy as np

from ..common import tensor
from ..spinn2d import App2D, Plotter2D, SPINN2D
from .ibvp2d_base import IBVP2D


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
    return App2D(pde_cls=Heat1D, nn_cls=SPINN2D, plotter_cls=Plotter2D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
