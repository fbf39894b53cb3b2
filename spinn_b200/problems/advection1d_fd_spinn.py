"""Linear advection u_t + a u_x = 0 of a Gaussian pulse on (-1, 1) by FD-SPINN time stepping
(problem of /root/reference/code/advection1d_fd_spinn.py).

    python -m spinn_b200.problems.advection1d_fd_spinn --gpu [--graph]
"""
import numpy as np
import torch

from ..spinn1d import SPINN1D
from .fd_spinn1d_base import AppFD1D, FDPlotter1D, FDSPINN1D

SPEED = 0.5
PULSE_CENTRE, PULSE_WIDTH = -0.3, 0.15


def _pulse(z, lib):
    return lib.exp(-0.5 * z ** 2)


class Advection1D(FDSPINN1D):
    domain = (-1.0, 1.0)

    def initial(self, x):
        return _pulse((x - PULSE_CENTRE) / PULSE_WIDTH, torch)

    def exact(self, x):
        return _pulse((x - SPEED * self.t - PULSE_CENTRE) / PULSE_WIDTH, np)

    def has_exact(self):
        return True

    def pde(self, x, u, ux, uxx, u0, dt):
        return -dt * SPEED * ux - u + u0


DEFAULTS = dict(nodes=20, samples=500, dt=0.0025, tol=1e-6, lr=1e-4, n_train=5000)


def make_app():
    return AppFD1D(pde_cls=Advection1D, nn_cls=SPINN1D, plotter_cls=FDPlotter1D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
