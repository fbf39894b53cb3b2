"""Helper of tests/test_host_surface.py: execute the reference's problem scripts BYTE-FOR-BYTE
UNCHANGED as `__main__` (their own defaults, their own App/PDE/Plotter subclasses) for a few
training steps, either on the stock reference modules (`ref`) or with spinn_b200/dropin's
common / spinn1d / spinn2d shadowing the reference's (`ours`, oracle-backed op on CPU).

    python tests/run_reference_scripts.py ref|ours <repo root> <reference code dir> <out.npz>

Build container only (needs /root/reference)."""
import contextlib
import io
import os
import runpy
import sys
import types

import numpy as np
import torch

# script -> extra CLI arguments (FD-SPINN: two time levels)
SCRIPTS = {
    "ode1": [], "ode2": [], "ode4": [], "ode1_var": [], "ode3_var": [],
    "poisson2d_pulse": [], "poisson2d_square_slit": [], "advection1d": [],
    "burgers1d_fd_spinn": ["--dt", "0.01", "-T", "0.02"],
    "advection1d_fd_spinn": ["--dt", "0.0025", "-T", "0.005"],
    "allen_cahn_fd": ["--dt", "0.005", "-T", "0.01"],
}


def main():
    side, root, ref, outp = sys.argv[1:5]
    for n in ("matplotlib", "matplotlib.pyplot", "mayavi", "mayavi.mlab"):      # not installed here
        sys.modules[n] = types.ModuleType(n)
    sys.path.insert(0, ref)
    if side == "ours":
        sys.path.insert(0, os.path.join(root, "spinn_b200", "dropin"))
        sys.path.insert(0, root)
        sys.path.insert(0, os.path.join(root, "tests"))
        from helpers import OracleBackend
        from spinn_b200 import ops
        ops.set_backend(OracleBackend())
    os.chdir(ref)
    out = {}
    for name, extra in SCRIPTS.items():
        np.random.seed(0)
        torch.manual_seed(0)
        sys.argv = [name + ".py", "--n-train", "10", "--n-skip", "5", "--tol", "-1"] + extra
        with contextlib.redirect_stdout(io.StringIO()):
            g = runpy.run_path(os.path.join(ref, name + ".py"), run_name="__main__")
        if side == "ours":
            import common
            assert common.__file__.startswith(root), common.__file__
        out[name] = np.asarray(g["app"].solver.loss)
    np.savez(outp, **out)


if __name__ == "__main__":
    main()
