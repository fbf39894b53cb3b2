"""BASELINE config 3: Poisson on an irregular domain given as point files
(mirrors /root/reference/code/poisson2d_irreg_dom.py:12-213).

The point sets are read from ``.dat`` files (first line = count, then ``x y z``
rows, reference :74-89) or, by default, from the packaged ``mesh_data.npz`` which
holds the same four arrays (tests/golden/make_golden.py wrote it from the
reference's code/mesh_data/*.dat).

    python -m spinn_b200.problems.poisson2d_irreg_dom --gpu
"""
import os

import numpy as np

from ..common import PDE, add_flags, tensor
from ..spinn2d import App2D, Plotter2D, SPINN2D
from . import derivs
from .sampling import SubSampling

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mesh_data.npz")
_KEYS = {"f_nodes_int": "interior_nodes", "f_nodes_bdy": "boundary_nodes",
         "f_samples_int": "interior_samples", "f_samples_bdy": "boundary_samples"}


def read_points(spec):
    """``npz:<key>`` -> packaged array; anything else -> reference-format text file."""
    if spec.startswith("npz:"):
        pts = np.load(_DATA)[spec[4:]]
        return pts[:, 0].copy(), pts[:, 1].copy()
    with open(spec, "r") as fh:
        n = int(fh.readline().split()[0])
        rows = [fh.readline().split()[:2] for _ in range(n)]
    pts = np.asarray(rows, dtype=float)
    return pts[:, 0], pts[:, 1]


class Poisson2D(SubSampling, PDE):
    #: `_compute_derivatives` keeps u_xx = d(u_x)/dx and u_yy = d(u_y)/dy only (as the reference's
    #: recipe): SPINN2D may skip the mixed derivative u_xy
    needs_mixed_derivative = False

    FLAGS = (
        (("--f_nodes_int", "-ni"), "f_nodes_int", "npz:interior_nodes", str, "File containing interior nodes."),
        (("--f_nodes_bdy", "-nb"), "f_nodes_bdy", "npz:boundary_nodes", str, "File containing boundary nodes."),
        (("--f_samples_int", "-si"), "f_samples_int", "npz:interior_samples", str, "File containing interior samples."),
        (("--f_samples_bdy", "-sb"), "f_samples_bdy", "npz:boundary_samples", str, "File containing boundary samples."),
        (("--sample-frac", "-f"), "sample_frac", 1.0, float, "Fraction of interior nodes used for sampling."),
    )

    @classmethod
    def from_args(cls, args):
        return cls(args.f_nodes_int, args.f_nodes_bdy, args.f_samples_int, args.f_samples_bdy,
                   args.sample_frac)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, kw)

    def _compute_derivatives(self, u, xs, ys):
        return derivs.jets_2d(u, xs, ys)

    _compute_derivatives.drops_mixed_derivative = True      # the recipe this class's flag vouches for

    def __init__(self, f_nodes_int, f_nodes_bdy, f_samples_int, f_samples_bdy, sample_frac=1.0):
        self.f_nodes_int, self.f_nodes_bdy = f_nodes_int, f_nodes_bdy
        self.f_samples_int, self.f_samples_bdy = f_samples_int, f_samples_bdy
        self.sample_frac = sample_frac
        self.interior_nodes = read_points(f_nodes_int)          # free
        self.boundary_nodes = read_points(f_nodes_bdy)          # fixed
        xi, yi = read_points(f_samples_int)
        self.interior_samples = (tensor(xi, requires_grad=True), tensor(yi, requires_grad=True))
        self.n_interior = len(xi)
        self.rng_interior = np.arange(self.n_interior)
        self.sample_size = int(self.sample_frac * self.n_interior)
        xb, yb = read_points(f_samples_bdy)
        self.boundary_samples = (tensor(xb, requires_grad=True), tensor(yb, requires_grad=True))

    def nodes(self):
        return self.interior_nodes

    def fixed_nodes(self):
        return self.boundary_nodes

    def interior(self):
        return self.pick(*self.interior_samples)

    def _sampling_device(self):
        return self.interior_samples[0].device

    def boundary(self):
        return self.boundary_samples

    def plot_points(self):
        xi, yi = read_points(self.f_samples_int)
        return tensor(xi), tensor(yi)            # tensors, not numpy (reference :142-144)

    def pde(self, x, y, u, ux, uy, uxx, uyy):
        return uxx + uyy + 1.0

    def has_exact(self):
        return False

    def _get_residue(self, nn):
        xs, ys = self.interior()
        u = nn(xs, ys)
        u, ux, uy, uxx, uyy = self._compute_derivatives(u, xs, ys)
        return self.pde(xs, ys, u, ux, uy, uxx, uyy)

    def interior_loss(self, nn):
        res = self._get_residue(nn)
        return (res ** 2).mean()

    def boundary_loss(self, nn):
        xb, yb = self.boundary()
        bc = nn(xb, yb) - 0.0
        return (bc ** 2).sum()


class PointCloud(Plotter2D):
    def get_plot_data(self):
        x, y = self.pde.plot_points()
        pn = self.nn(x, y).detach().cpu().numpy()
        pn.shape = x.shape
        return x.detach().cpu().numpy(), y.detach().cpu().numpy(), pn


DEFAULTS = dict(lr=1e-3)


def make_app():
    return App2D(pde_cls=Poisson2D, nn_cls=SPINN2D, plotter_cls=PointCloud)


if __name__ == "__main__":
    make_app().run(**DEFAULTS)
