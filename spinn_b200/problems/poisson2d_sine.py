"""BASELINE config 2: Poisson on the unit square with a sine source, Dirichlet walls
(mirrors /root/reference/code/poisson2d_sine.py).

    python -m spinn_b200.problems.poisson2d_sine --gpu
"""
import numpy as np
import torch

from ..common import tensor
from ..spinn2d import App2D, Plotter2D, SPINN2D
from .pde2d_base import RegularPDE

PI = np.pi


class Poisson2D(RegularPDE):
    #: the boundary penalty is a plain sum over boundary points -> they can be sharded across ranks
    boundary_reduction = "sum"

    def pde(self, x, y, u, ux, uy, uxx, uyy):
        f = 20 * PI ** 2 * torch.sin(2 * PI * x) * torch.sin(4 * PI * y)
        return 0.1 * (uxx + uyy + f)

    def has_exact(self):
        return True

    def exact(self, x, y):
        return np.sin(2 * PI * x) * np.sin(4 * PI * y)

    def boundary_loss(self, nn):
        xb, yb = self.boundary()
        # the reference recomputes the (constant) target through NumPy every step
        # (poisson2d_sine.py:27-31: D2H, exact(), H2D); same values, computed once here
        ub = getattr(self, "_ub", None)
        if ub is None or ub.device != xb.device:
            xbn, ybn = (t.detach().cpu().numpy() for t in (xb, yb))
            ub = self._ub = tensor(self.exact(xbn, ybn))
        bc = nn(xb, yb) - ub
        return (bc ** 2).sum()


DEFAULTS = dict(nodes=100, samples=400, n_train=25000, lr=1e-3, tol=1e-3)


def make_app():
    return App2D(pde_cls=Poisson2D, nn_cls=SPINN2D, plotter_cls=Plotter2D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
