"""1-D heat equation u_t = c^2 u_xx, u(0)=u(1)=0, by FD-SPINN time stepping
(problem of /root/reference/code/heat1d_fd_spinn.py).

    python -m spinn_b200.problems.heat1d_fd_spinn --gpu [--graph]
"""
import numpy as np
import torch

from ..spinn1d import SPINN1D
from .fd_spinn1d_base import AppFD1D, FDPlotter1D, FDSPINN1D

AMPLITUDE, DIFFUSIVITY = 2.0, 1.0      # u(x, t) = A exp(-pi^2 c^2 t) sin(pi x)


class Heat1D(FDSPINN1D):
    n_plot = staticmethod(lambda ns: 50)

    def initial(self, x):
        return AMPLITUDE * torch.sin(np.pi * x)

    def exact(self, x):
        decay = np.exp(-np.pi * np.pi * DIFFUSIVITY * DIFFUSIVITY * self.t)
        return AMPLITUDE * decay * np.sin(np.pi * x)

    def has_exact(self):
        return True

    def pde(self, x, u, ux, uxx, u0, dt):
        # implicit Euler residual; the step comes from the problem object, as in the reference
        return DIFFUSIVITY * DIFFUSIVITY * self.dt * uxx - u + u0


DEFAULTS = dict(nodes=10, dt=0.01, samples=100, tol=2e-5, T=0.1, lr=1e-3)


def make_app():
    return AppFD1D(pde_cls=Heat1D, nn_cls=SPINN1D, plotter_cls=FDPlotter1D)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
