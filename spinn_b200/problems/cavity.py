"""BASELINE config 4: lid-driven cavity, steady Navier-Stokes, three outputs (u, v, p)
(mirrors /root/reference/code/cavity.py:12-211).

    python -m spinn_b200.problems.cavity --gpu
"""
import os

import numpy as np
import torch

from ..common import tensor
from ..spinn2d import App2D, Plotter2D, SPINN2D
from . import derivs
from .pde2d_base import RegularPDE, _grid, _side


class CavityPDE(RegularPDE):
    NU = 0.01
    interior_reduction = "sum"     # interior_loss is a SUM over points (cavity.py:96-100)

    def __init__(self, n_nodes, ns, nb=None, nbs=None, sample_frac=1.0):
        self.sample_frac = sample_frac
        # free nodes: end-point style grid with a full cell of margin (cavity.py:16-22)
        self.n = n = _side(n_nodes)
        m = 1.0 / (n + 1)
        self.i_nodes = _grid(np.mgrid[m:1.0 - m:n * 1j])
        # fixed nodes (cavity.py:24-36)
        self.nb = nb = n if nb is None else nb
        if nb == 0:
            self.f_nodes = ([], [])
        else:
            e = np.linspace(1.0 / nb, 1.0 - 1.0 / nb, nb)
            one = np.ones_like(e)
            self.f_nodes = (np.hstack((e, one, e, 0.0 * one)), np.hstack((one * 0.0, e, one, e)))
        # interior samples (cavity.py:38-46)
        self.ns = ns = _side(ns)
        m = 1.0 / ns
        sx, sy = _grid(np.mgrid[m:1.0 - m:ns * 1j])
        self.p_samples = (tensor(sx, requires_grad=True), tensor(sy, requires_grad=True))
        # wall samples, order top, bottom, left, right (cavity.py:48-65)
        self.nbs = nbs = ns if nbs is None else nbs
        w = np.linspace(m, 1.0 - m, nbs)
        o = np.ones_like(w)
        leaf = lambda a: tensor(a, requires_grad=True)
        self.left = (leaf(0.0 * o), leaf(w))
        self.right = (leaf(o), leaf(w))
        self.bottom = (leaf(w), leaf(o) * 0.0)
        self.top = (leaf(w), leaf(o))
        walls = (self.top, self.bottom, self.left, self.right)
        self.b_samples = (torch.cat([s[0] for s in walls]), torch.cat([s[1] for s in walls]))
        self.n_interior = len(sx)
        self.rng_interior = np.arange(self.n_interior)
        self.sample_size = int(self.sample_frac * self.n_interior)

    def enable_static_sampling(self):
        """CUDA-graph mode: besides the static index buffer, cut the persistent autograd graph
        of b_samples (a cat() of leaf tensors built once, cavity.py:62-65 -- the reason the
        reference needs retain_graph=True): its AccumulateGrad nodes live on the legacy stream
        and cannot be replayed inside a capture.  Same values, fresh leaves."""
        super().enable_static_sampling()
        self.b_samples = tuple(t.detach().clone().requires_grad_(True) for t in self.b_samples)

    def _compute_gradient(self, u, xs, ys):
        return derivs.grad_xy(u, xs, ys)

    def n_vars(self):
        return 3

    def has_exact(self):
        return False

    def interior_loss(self, nn):
        xs, ys = self.interior()
        U = nn(xs, ys)
        u, ux, uy, uxx, uyy = self._compute_derivatives(U[:, 0], xs, ys)
        v, vx, vy, vxx, vyy = self._compute_derivatives(U[:, 1], xs, ys)
        px, py = self._compute_gradient(U[:, 2], xs, ys)
        nu = self.NU
        ce = ((ux + vy) ** 2).sum()
        mex = ((u * ux + v * uy + px - nu * (uxx + uyy)) ** 2).sum()
        mey = ((u * vx + v * vy + py - nu * (vxx + vyy)) ** 2).sum()
        return ce + (mex + mey)

    def boundary_loss(self, nn):
        bc_weight = self.ns * 20
        xb, yb = self.boundary()
        Ub = nn(xb, yb)
        ub, vb, pb = Ub[:, 0], Ub[:, 1], Ub[:, 2]
        pbx, pby = self._compute_gradient(pb, xb, yb)
        nt = len(self.top[0])
        loss = ((ub[:nt] - 1.0) ** 2).sum()
        loss = loss + (ub[nt:] ** 2).sum() * bc_weight
        loss = loss + (vb ** 2).sum() * bc_weight
        loss = loss + (pby[:2 * nt] ** 2).sum()
        loss = loss + (pbx[2 * nt:] ** 2).sum()
        return loss

    def plot_points(self):
        n = min(self.ns * 2, 200)
        x, y = np.mgrid[0:1:n * 1j, 0:1:n * 1j]
        return x, y


class CavityPlotter(Plotter2D):
    def get_plot_data(self):
        x, y = self.pde.plot_points()
        pn = self.nn(tensor(x.ravel()), tensor(y.ravel())).detach().cpu().numpy()
        pn.shape = x.shape + (3,)
        return x, y, pn

    def centerline(self):
        """u along x=0.5 and v along y=0.5 (cavity.py:195-201; used for the Ghia comparison)."""
        lc = tensor(np.linspace(0, 1, 100))
        mid = tensor(0.5 * np.ones(100))
        data = self.nn(torch.cat((lc, mid)), torch.cat((mid, lc))).detach().cpu().numpy()
        return lc.cpu().numpy(), data[:, 0][100:], data[:, 1][:100]

    def save(self, dirname):
        torch.save(self.nn.state_dict(), os.path.join(dirname, "model.pt"))
        x, y, u = self.get_plot_data()
        xc, uc, vc = self.centerline()
        np.savez(os.path.join(dirname, "results.npz"), x=x, y=y, u=u, xc=xc, uc=uc, vc=vc)


DEFAULTS = dict(nodes=100, samples=2500, sample_frac=0.1, lr=5e-4, n_train=5000)


def make_app():
    return App2D(pde_cls=CavityPDE, nn_cls=SPINN2D, plotter_cls=CavityPlotter)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
