"""Regular-grid 2-D problem base (mirrors /root/reference/code/pde2d_base.py:10-141)."""
import numpy as np

from ..common import PDE, add_flags, tensor
from . import derivs
from .sampling import SubSampling


def _cell_centres(n, half):
    """n points from `half` to 1-`half` inclusive -- np.mgrid, bit-for-bit what the
    reference uses (pde2d_base.py:71-72, 89-92)."""
    return np.mgrid[half:1.0 - half:n * 1j]


def _grid(v):
    x, y = np.meshgrid(v, v, indexing="ij")
    return x.ravel(), y.ravel()


def _side(count):
    return round(np.sqrt(count) + 0.49)


class RegularPDE(SubSampling, PDE):
    #: `_compute_derivatives` keeps u_xx = d(u_x)/dx and u_yy = d(u_y)/dy only (as the reference's
    #: recipe): SPINN2D may skip the mixed derivative u_xy
    needs_mixed_derivative = False

    FLAGS = (
        (("--nodes", "-n"), "nodes", 25, int, "Number of nodes to use."),
        (("--samples", "-s"), "samples", 100, int, "Number of sample points to use."),
        (("--b-nodes",), "b_nodes", None, int, "Number of boundary nodes to use per edge"),
        (("--b-samples",), "b_samples", None, int, "Number of boundary samples to use per edge."),
        (("--sample-frac", "-f"), "sample_frac", 1.0, float, "Fraction of interior nodes used for sampling."),
    )

    @classmethod
    def from_args(cls, args):
        return cls(args.nodes, args.samples, args.b_nodes, args.b_samples, args.sample_frac)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, kw)

    def _compute_derivatives(self, u, xs, ys):
        return derivs.jets_2d(u, xs, ys)

    _compute_derivatives.drops_mixed_derivative = True      # the recipe this class's flag vouches for

    def __init__(self, n_nodes, ns, nb=None, nbs=None, sample_frac=1.0):
        self.sample_frac = sample_frac
        # free nodes: n x n cell-centred grid (pde2d_base.py:66-74)
        self.n = n = _side(n_nodes)
        self.i_nodes = _grid(_cell_centres(n, 0.5 / (n + 1)))
        # fixed nodes: nb per edge, order bottom, right, top, left (pde2d_base.py:76-84)
        self.nb = nb = n if nb is None else nb
        e = np.linspace(0.5 / nb, 1.0 - 0.5 / nb, nb)
        one, zero = np.ones_like(e), np.zeros_like(e)
        self.f_nodes = (np.hstack((e, one, e, zero)), np.hstack((zero, e, one, e)))
        # interior samples (pde2d_base.py:86-97)
        self.ns = ns = _side(ns)
        half = 0.5 / ns
        sx, sy = _grid(_cell_centres(ns, half))
        self.p_samples = (tensor(sx, requires_grad=True), tensor(sy, requires_grad=True))
        self.n_interior = len(sx)
        self.rng_interior = np.arange(self.n_interior)
        self.sample_size = int(self.sample_frac * self.n_interior)
        # boundary samples: ring of an nbs x nbs end-point grid outside [half, 1-half] (pde2d_base.py:99-107)
        self.nbs = nbs = ns if nbs is None else nbs
        bx, by = np.mgrid[0.0:1.0:nbs * 1j, 0.0:1.0:nbs * 1j]
        lo, hi = half, 1.0 - half
        ring = ((bx < lo) | (bx > hi)) | ((by < lo) | (by > hi))
        self.b_samples = (tensor(bx[ring], requires_grad=True), tensor(by[ring], requires_grad=True))

    def nodes(self):
        return self.i_nodes

    def fixed_nodes(self):
        return self.f_nodes

    def interior(self):
        return self.pick(*self.p_samples)

    def _sampling_device(self):
        return self.p_samples[0].device

    def boundary(self):
        return self.b_samples

    def plot_points(self):
        n = self.ns * 2
        x, y = np.mgrid[0:1:n * 1j, 0:1:n * 1j]
        return x, y

    def _get_residue(self, nn):
        xs, ys = self.interior()
        u = nn(xs, ys)
        u, ux, uy, uxx, uyy = self._compute_derivatives(u, xs, ys)
        return self.pde(xs, ys, u, ux, uy, uxx, uyy)

    def interior_loss(self, nn):
        res = self._get_residue(nn)
        return (res ** 2).mean()
