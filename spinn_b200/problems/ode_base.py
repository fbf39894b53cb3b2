"""1-D problem base: uniform interior nodes/samples on an interval, Dirichlet end points
(behaviour of /root/reference/code/ode_base.py:8-92; FD-SPINN re-uses it on other intervals)."""
import numpy as np

from ..common import PDE, add_flags, tensor
from . import derivs
from .sampling import SubSampling


def interior_grid(count, lo=0.0, hi=1.0):
    """The `count` interior points of the uniform grid with count+2 points on [lo, hi]
    (ode_base.py:49: i/(n+1) on the unit interval; advection1d_fd_spinn.py:15-20 on (-1, 1))."""
    if (lo, hi) == (0.0, 1.0):
        return np.arange(1, count + 1) / (count + 1)
    return np.asarray([lo + (hi - lo) * i / (count + 1) for i in range(1, count + 1)])


class BasicODE(SubSampling, PDE):
    FLAGS = (
        (("--nodes", "-n"), "nodes", 10, int, "Number of nodes to use."),
        (("--samples", "-s"), "samples", 30, int, "Number of sample points to use."),
        (("--sample-frac", "-f"), "sample_frac", 1.0, float, "Fraction of interior nodes used for sampling."),
    )
    domain = (0.0, 1.0)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, kw)

    @classmethod
    def from_args(cls, args):
        return cls(args.nodes, args.samples, args.sample_frac)

    def __init__(self, n, ns, sample_frac=1.0):
        lo, hi = self.domain
        self.n, self.ns, self.sample_frac = n, ns, sample_frac
        self.xn = interior_grid(n, lo, hi)                            # free node centres
        self.xs = tensor(interior_grid(ns, lo, hi), requires_grad=True)   # collocation points
        self.xbn = np.array([lo, hi])                                 # fixed nodes = boundary points
        self.xb = tensor(self.xbn)
        self.rng_interior = np.arange(ns)
        self.sample_size = int(sample_frac * ns)

    # geometry handed to the network / plotter
    def nodes(self):
        return self.xn

    def fixed_nodes(self):
        return self.xbn

    def boundary(self):
        return self.xb

    def plot_points(self):
        lo, hi = self.domain
        return np.linspace(lo, hi, 2 * self.ns)

    def interior(self):
        # one draw from the unseeded global NumPy stream per call (ode_base.py:64-68)
        return self.pick(self.xs)[0]

    def _sampling_device(self):
        return self.xs.device

    # physics
    def _compute_derivatives(self, u, x):
        return derivs.jets_1d(u, x)

    def _get_residue(self, nn):
        x = self.interior()
        return self.pde(*((x,) + self._compute_derivatives(nn(x), x)))

    def interior_loss(self, nn):
        return (self._get_residue(nn) ** 2).mean()
