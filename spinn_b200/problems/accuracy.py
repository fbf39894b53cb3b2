"""External accuracy references the reference code base uses for the two configs without an
exact solution, with the same error recipes:
  irregular-domain Poisson vs a FEM field   /root/reference/code/poisson2d_irreg_dom.py:179-192
      (code/fem/poisson_irregular000000.vtu: 2074 points + nodal values, stored here as arrays)
  lid-driven cavity vs Ghia et al. Re=100   /root/reference/code/misc/cavity_error.py:12-31
      (code/data/ghia_re_100.txt; centre-line sampling as CavityPlotter.save, cavity.py:195-201)
"""
import os

import numpy as np

from ..common import tensor

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "accuracy_refs.npz")


def load_refs():
    d = np.load(_DATA)
    return {k: d[k] for k in d.files}


def fem_errors(nn, refs=None):
    """L1, L2, Linf of nn against the FEM solution at the FEM mesh points."""
    refs = refs or load_refs()
    u = nn(tensor(refs["fem_x"]), tensor(refs["fem_y"])).detach().cpu().numpy()
    du = u - refs["fem_u"]
    return float(np.mean(np.abs(du))), float(np.sqrt(np.mean(du ** 2))), float(np.max(np.abs(du)))


def ghia_errors(nn, refs=None):
    """Linf of u along x=0.5 and of v along y=0.5 against the Ghia table (linear interpolation)."""
    refs = refs or load_refs()
    lc = np.linspace(0, 1, 100)
    mid = 0.5 * np.ones(100)
    data = nn(tensor(np.concatenate((lc, mid))), tensor(np.concatenate((mid, lc)))).detach().cpu().numpy()
    vc, uc = data[:, 1][:100], data[:, 0][100:]
    ud = np.abs(np.interp(refs["ghia_y"], lc, uc) - refs["ghia_u"])
    vd = np.abs(np.interp(refs["ghia_x"], lc, vc) - refs["ghia_v"])
    return float(np.max(ud)), float(np.max(vd))
