"""Space-time (x, t) problem base (mirrors /root/reference/code/ibvp2d_base.py:12-196)."""
import numpy as np

from ..common import PDE, add_flags, tensor
from . import derivs
from .sampling import SubSampling


class IBVP2D(SubSampling, PDE):
    #: `_compute_derivatives` keeps u_xx = d(u_x)/dx and u_yy = d(u_y)/dy only (as the reference's
    #: recipe): SPINN2D may skip the mixed derivative u_xy
    needs_mixed_derivative = False

    FLAGS = (
        (("--nodes", "-n"), "nodes", 25, int, "Number of nodes to use."),
        (("--samples", "-s"), "samples", 100, int, "Number of sample points to use."),
        (("--b-nodes",), "b_nodes", None, int, "Number of boundary nodes to use per edge."),
        (("--b-samples",), "b_samples", None, int, "Number of boundary samples to use per edge."),
        (("--xL",), "xL", 0.0, float, "Left end of domain."),
        (("--xR",), "xR", 1.0, float, "Right end of domain."),
        (("--duration",), "T", 1.0, float, "Duration of simulation."),
        (("--sample-frac", "-f"), "sample_frac", 1.0, float, "Fraction of interior nodes used for sampling."),
    )

    @classmethod
    def from_args(cls, args):
        return cls(args.nodes, args.samples, args.b_nodes, args.b_samples,
                   args.xL, args.xR, args.T, args.sample_frac)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, kw)

    def _compute_derivatives(self, u, xs, ts):
        return derivs.jets_2d(u, xs, ts)

    _compute_derivatives.drops_mixed_derivative = True      # the recipe this class's flag vouches for

    @staticmethod
    def _get_points_split(L, T, n):
        size = np.sqrt(L * T / n)                                     # ibvp2d_base.py:83-87
        return round(L / size + 0.49), round(T / size + 0.49)

    @staticmethod
    def _get_points_mgrid(xL, xR, T, nx, nt, endpoint=False):
        if endpoint:
            xv, tv = np.mgrid[xL:xR:nx * 1j], np.mgrid[0.0:T:nt * 1j]
        else:                                                         # ibvp2d_base.py:95-102
            half = 0.5 * (xR - xL) / (nx + 1)
            xv = np.mgrid[xL + half:xR - half:nx * 1j]
            tv = np.mgrid[T / (nt + 1):T:nt * 1j]
        x, t = np.meshgrid(xv, tv, indexing="ij")
        return x.ravel(), t.ravel()

    @staticmethod
    def _get_boundary_nodes(xL, xR, T, nx, nt):
        x0 = xL + np.arange(nx) * (xR - xL) / (nx - 1)                # t = 0 edge
        tt = np.arange(1, nt + 1) * T / nt                            # both side walls
        xs = np.concatenate((x0, np.full(nt, xL), np.full(nt, xR)))
        ts = np.concatenate((np.zeros(nx), tt, tt))
        return xs, ts

    def __init__(self, n_nodes, ns, nb=None, nbs=None, xL=0.0, xR=1.0, T=1.0, sample_frac=1.0):
        self.xL, self.xR, self.T = xL, xR, T
        self.sample_frac = sample_frac
        L = xR - xL
        self.nx, self.nt = self._get_points_split(L, T, n_nodes)
        self.i_nodes = self._get_points_mgrid(xL, xR, T, self.nx, self.nt)
        nbx, nbt = (self.nx, self.nt) if nb is None else self._get_points_split(L, T, nb + 1)
        self.f_nodes = ([], []) if nb == 0 else self._get_boundary_nodes(xL, xR, T, nbx, nbt)
        self.nsx, self.nst = self._get_points_split(L, T, ns)
        xi, ti = self._get_points_mgrid(xL, xR, T, self.nsx, self.nst)
        self.p_samples = (tensor(xi, requires_grad=True), tensor(ti, requires_grad=True))
        self.n_interior = len(xi)
        self.rng_interior = np.arange(self.n_interior)
        self.sample_size = int(self.sample_frac * self.n_interior)
        nbsx, nbst = (self.nsx, self.nst) if nbs is None else self._get_points_split(L, T, nbs)
        xb, tb = self._get_boundary_nodes(xL, xR, T, nbsx, nbst)
        self.b_samples = (tensor(xb, requires_grad=True), tensor(tb, requires_grad=True))

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
        return np.mgrid[self.xL:self.xR:self.nsx * 1j, 0.0:self.T:self.nst * 1j]

    def _get_residue(self, nn):
        xs, ts = self.interior()
        u = nn(xs, ts)
        u, ux, ut, uxx, utt = self._compute_derivatives(u, xs, ts)
        return self.pde(xs, ts, u, ux, ut, uxx, utt)

    def interior_loss(self, nn):
        res = self._get_residue(nn)
        return (res ** 2).mean()
