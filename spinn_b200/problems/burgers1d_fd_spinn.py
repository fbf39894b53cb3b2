"""Inviscid 1-D Burgers equation u_t + u u_x = 0, u(0)=u(1)=0, by FD-SPINN time stepping
(problem of /root/reference/code/burgers1d_fd_spinn.py; no exact solution).

    python -m spinn_b200.problems.burgers1d_fd_spinn --gpu [--graph]
"""
import numpy as np
import torch

from ..spinn1d import SPINN1D
from .fd_spinn1d_base import AppFD1D, FDPlotter1D, FDSPINN1D


class Burgers1D(FDSPINN1D):
    def initial(self, x):
        return torch.sin(2.0 * np.pi * x)

    def has_exact(self):
        return False

    def pde(self, x, u, ux, uxx, u0, dt):
        return -dt * u * ux - u + u0


DEFAULTS = dict(nodes=40, samples=400, dt=0.01, tol=5e-6, lr=1e-4, n_train=5000, n_skip=500)


def make_app():
    return AppFD1D(pde_cls=Burgers1D, nn_cls=SPINN1D, plotter_cls=FDPlotter1D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
