"""Import the UNMODIFIED reference (nn4pde/SPINN `code/`) from baseline/_ref/code.

The reference is a directory of flat Python modules without packaging metadata (no setup.py /
pyproject.toml), so `pip install --target baseline/_ref /root/reference` has nothing to build;
`__graft_entry__.build()` copies `/root/reference/code` to `baseline/_ref/code` instead (git-ignored,
travels to the GPU box with the gpurun snapshot).  Users: `bench.py --impl reference` /
`cpu_baseline` (times the reference itself) and the `-m gpu` test that executes the reference's
problem scripts unchanged on top of spinn_b200/dropin.  The product never imports this.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CODE = os.path.join(HERE, "_ref", "code")
UPSTREAM = "/root/reference/code"

# modules the reference imports at top level that are not installed in this image (plotting only)
_STUBS = ("matplotlib", "matplotlib.pyplot", "mayavi", "mayavi.mlab")
# flat module names that exist both in the reference and in spinn_b200/dropin
_FLAT = ("common", "spinn1d", "spinn2d", "ode_base", "pde2d_base", "ibvp2d_base", "poisson2d_sine",
         "poisson2d_irreg_dom", "cavity", "ode3", "heat1d", "burgers1d", "fd_spinn1d_base")


def install():
    """Copy the reference's code/ directory next to this file (build container only)."""
    if not os.path.isdir(UPSTREAM):
        return os.path.isdir(REF_CODE)
    os.makedirs(os.path.dirname(REF_CODE), exist_ok=True)
    shutil.copytree(UPSTREAM, REF_CODE, dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


def available():
    return os.path.isfile(os.path.join(REF_CODE, "spinn2d.py"))


def stub_plotting_modules():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mayavi"].mlab = sys.modules["mayavi.mlab"]


def load(*names):
    """Import reference modules by flat name from baseline/_ref/code (the stock ones: nothing of
    spinn_b200 is on the path ahead of them).  Returns a namespace with one attribute per module."""
    if not available():
        raise FileNotFoundError(f"{REF_CODE} missing: run __graft_entry__.build() where /root/reference exists")
    stub_plotting_modules()
    for n in _FLAT:                      # a drop-in twin imported earlier must not shadow the reference
        m = sys.modules.get(n)
        if m is not None and not getattr(m, "__file__", "").startswith(REF_CODE):
            del sys.modules[n]
    if REF_CODE in sys.path:
        sys.path.remove(REF_CODE)
    sys.path.insert(0, REF_CODE)
    ns = types.SimpleNamespace()
    for n in names:
        mod = importlib.import_module(n)
        assert mod.__file__.startswith(REF_CODE), (n, mod.__file__)
        setattr(ns, n, mod)
    return ns
