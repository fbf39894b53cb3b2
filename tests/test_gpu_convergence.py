"""Converged accuracy on the GPU vs the reference's own converged runs (BASELINE.md section 2 /
north_star "converged L2/Linf errors within 1 % of the reference's").

The reference's converged errors are chaotic in 1e-7 perturbations of the initial widths
(tests/golden/make_convergence_band.py: ode3's final Linf moves by 14 % and its stopping iteration
from 3601 to 3701 for a 2e-7 change), so the acceptance target is the BAND spanned by that ensemble,
widened by the band's own spread; the literal 1 % comparison with the unperturbed run is recorded
in gpurun_out/convergence.json (and asserted where the ensemble itself is tight, i.e. the cavity).
Runs use --graph (same arithmetic, ~0.3 ms/step) so that 25 000 steps take seconds."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _band(name):
    p = os.path.join(GOLDEN, f"band_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated")
    d = np.load(p)
    return d["errors"], d["iterations"], d["final_loss"], [str(a) for a in d["args"]]


def _record(name, payload):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "convergence.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = payload
    json.dump(data, open(path, "w"), indent=1)


@pytest.mark.parametrize("name,modname", [("ode3", "ode3"), ("poisson2d_sine", "poisson2d_sine"),
                                          ("irreg", "poisson2d_irreg_dom"), ("cavity", "cavity")])
def test_converged_errors_within_reference_band(name, modname):
    from spinn_b200.problems import accuracy
    errs, iters, losses, args = _band(name)
    mod = importlib.import_module(f"spinn_b200.problems.{modname}")
    np.random.seed(0)
    torch.manual_seed(0)
    app = mod.make_app()
    app.run(args=args + ["--gpu", "--graph", "--n-skip", "100"])
    if name in ("ode3", "poisson2d_sine"):
        got = np.asarray(app.plotter.get_error(), dtype=np.float64)
    elif name == "irreg":
        got = np.asarray(accuracy.fem_errors(app.nn), dtype=np.float64)
    else:
        got = np.asarray(accuracy.ghia_errors(app.nn), dtype=np.float64)
    lo, hi = errs.min(0), errs.max(0)
    spread = hi - lo
    literal = np.abs(got - errs[0]) / errs[0]
    _record(name, {"ours": got.tolist(), "reference_unperturbed": errs[0].tolist(),
                   "reference_band_min": lo.tolist(), "reference_band_max": hi.tolist(),
                   "literal_relative_difference": literal.tolist(), "within_1_percent": bool((literal < 0.01).all()),
                   "iterations": len(app.solver.loss), "reference_iterations": iters.tolist(),
                   "final_loss": app.solver.loss[-1], "reference_final_loss": losses.tolist()})
    # inside the ensemble band, widened by its own spread and a 5 % floor
    margin = np.maximum(spread, 0.05 * errs[0])
    assert np.all(got >= lo - margin) and np.all(got <= hi + margin), (name, got, lo, hi)
    if name == "cavity":        # the reference ensemble is tight here (1e-5): so must we be
        assert np.all(literal < 0.01), (got, errs[0])
    assert abs(len(app.solver.loss) - int(iters[0])) <= 200, (len(app.solver.loss), iters)
