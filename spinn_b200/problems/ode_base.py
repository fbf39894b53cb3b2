"""1-D problem base (mirrors /root/reference/code/ode_base.py:8-92)."""
import numpy as np

from ..common import PDE, add_flags, tensor
from . import derivs


class BasicODE(PDE):
    FLAGS = (
        (("--nodes", "-n"), "nodes", 10, int, "Number of nodes to use."),
        (("--samples", "-s"), "samples", 30, int, "Number of sample points to use."),
        (("--sample-frac", "-f"), "sample_frac", 1.0, float, "Fraction of interior nodes used for sampling."),
    )

    @classmethod
    def from_args(cls, args):
        return cls(args.nodes, args.samples, args.sample_frac)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        add_flags(parser, cls.FLAGS, kw)

    def _compute_derivatives(self, u, x):
        return derivs.jets_1d(u, x)

    def __init__(self, n, ns, sample_frac=1.0):
        self.n, self.ns, self.sample_frac = n, ns, sample_frac
        self.xn = np.arange(1, n + 1) / (n + 1)                       # ode_base.py:49
        self.xs = tensor(np.arange(1, ns + 1) / (ns + 1), requires_grad=True)
        self.xbn = np.array([0.0, 1.0])
        self.xb = tensor(self.xbn)
        self.rng_interior = np.arange(self.ns)
        self.sample_size = int(self.sample_frac * self.ns)

    def nodes(self):
        return self.xn

    def fixed_nodes(self):
        return self.xbn

    def interior(self):
        if abs(self.sample_frac - 1.0) < 1e-3:
            return self.xs
        idx = np.random.choice(self.rng_interior, size=self.sample_size, replace=False)
        return self.xs[idx]

    def boundary(self):
        return self.xb

    def plot_points(self):
        return np.linspace(0.0, 1.0, 2 * self.ns)

    def _get_residue(self, nn):
        xs = self.interior()
        u = nn(xs)
        u, ux, uxx = self._compute_derivatives(u, xs)
        return self.pde(xs, u, ux, uxx)

    def interior_loss(self, nn):
        res = self._get_residue(nn)
        return (res ** 2).mean()
