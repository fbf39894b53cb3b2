"""Linear advection of a Gaussian pulse on (-1, 1) by FD-SPINN time stepping (mirrors
/root/reference/code/advection1d_fd_spinn.py).

    python -m spinn_b200.problems.advection1d_fd_spinn --gpu [--graph]
"""
import numpy as np
import torch

from ..common import tensor
from ..spinn1d import SPINN1D
from .fd_spinn1d_base import AppFD1D, FDPlotter1D, FDSPINN1D

SPEED = 0.5
PULSE_CENTRE, PULSE_WIDTH = -0.3, 0.15


class Advection1D(FDSPINN1D):
    def __init__(self, n, ns, sample_frac=1.0, dt=1e-3, T=1.0, t_skip=10):
        super().__init__(n, ns, sample_frac, dt=dt, T=T, t_skip=t_skip)
        # the unit-interval geometry of the base class is replaced by (-1, 1)
        self.xn = np.asarray([-1.0 + 2 * i / (n + 1) for i in range(1, n + 1)])
        self.xs = tensor([-1.0 + 2 * i / (ns + 1) for i in range(1, ns + 1)], requires_grad=True)
        self.xbn = np.array([-1.0, 1.0])
        self.xb = tensor(self.xbn)
        self.u0 = self.initial(self.xs)

    def initial(self, x):
        z = (x - PULSE_CENTRE) / PULSE_WIDTH
        return torch.exp(-0.5 * z ** 2)

    def pde(self, x, u, ux, uxx, u0, dt):
        return -dt * SPEED * ux - u + u0

    def has_exact(self):
        return True

    def exact(self, x):
        z = (x - SPEED * self.t - PULSE_CENTRE) / PULSE_WIDTH
        return np.exp(-0.5 * z ** 2)

    def boundary_loss(self, nn):
        bc = nn(self.boundary()) - 0.0
        return 100 * (bc ** 2).sum()

    def plot_points(self):
        return np.linspace(-1.0, 1.0, min(2 * self.ns, 500))


DEFAULTS = dict(nodes=20, samples=500, dt=0.0025, tol=1e-6, lr=1e-4, n_train=5000)


def make_app():
    return AppFD1D(pde_cls=Advection1D, nn_cls=SPINN1D, plotter_cls=FDPlotter1D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
