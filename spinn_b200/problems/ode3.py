"""BASELINE config 1: u'' + f = 0 on (0,1), u(0)=u(1)=0 (mirrors /root/reference/code/ode3.py).

    python -m spinn_b200.problems.ode3 --gpu
"""
import numpy as np
import torch

from ..common import tensor
from ..spinn1d import App1D, Plotter1D, SPINN1D
from .ode_base import BasicODE

K_WIDTH = 0.01


class ODESimple(BasicODE):
    def pde(self, x, u, ux, uxx):
        K = K_WIDTH
        f = -(K * (-6 * x + 4 / 3.) + x * (2 * x - 2. / 3) ** 2) * torch.exp(-(x - 1 / 3.) ** 2 / K) / K ** 2
        return uxx + f

    def has_exact(self):
        return True

    def exact(self, x):
        K = K_WIDTH
        return x * (np.exp(-((x - (1.0 / 3.0)) ** 2) / K) - np.exp(-4.0 / (9.0 * K)))

    def boundary_loss(self, nn):
        ub = getattr(self, "_ub", None)        # constant target, built once (reference: every step)
        if ub is None or ub.device != self.xb.device:
            ub = self._ub = tensor(self.exact(self.xbn))
        bc = nn(self.boundary()) - ub
        return 10 * (bc ** 2).sum()

    def plot_points(self):
        return np.linspace(0.0, 1.0, 50)


DEFAULTS = dict(nodes=3, samples=60, n_train=20000, lr=1e-3, tol=1e-3)


def make_app():
    return App1D(pde_cls=ODESimple, nn_cls=SPINN1D, plotter_cls=Plotter1D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
