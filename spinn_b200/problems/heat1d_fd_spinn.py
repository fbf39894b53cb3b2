"""1-D heat equation by FD-SPINN time stepping (mirrors /root/reference/code/heat1d_fd_spinn.py).

    python -m spinn_b200.problems.heat1d_fd_spinn --gpu [--graph]
"""
import numpy as np
import torch

from ..spinn1d import SPINN1D
from .fd_spinn1d_base import AppFD1D, FDPlotter1D, FDSPINN1D

AMPLITUDE = 2.0
SPEED = 1.0


class Heat1D(FDSPINN1D):
    def initial(self, x):
        return AMPLITUDE * torch.sin(np.pi * x)

    def pde(self, x, u, ux, uxx, u0, dt):
        # the reference reads self.dt here, not the argument (heat1d_fd_spinn.py:19)
        return SPEED * SPEED * self.dt * uxx - u + u0

    def has_exact(self):
        return True

    def exact(self, x):
        return AMPLITUDE * np.exp(-np.pi * np.pi * SPEED * SPEED * self.t) * np.sin(np.pi * x)

    def boundary_loss(self, nn):
        bc = nn(self.boundary()) - 0.0
        return 100 * (bc ** 2).sum()

    def plot_points(self):
        return np.linspace(0.0, 1.0, 50)


DEFAULTS = dict(nodes=10, dt=0.01, samples=100, tol=2e-5, T=0.1, lr=1e-3)


def make_app():
    return AppFD1D(pde_cls=Heat1D, nn_cls=SPINN1D, plotter_cls=FDPlotter1D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
