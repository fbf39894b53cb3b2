#!/usr/bin/env python
"""Full training runs of a chaotic config (irreg: irregular-domain Poisson, 2500 steps, errors vs
FEM; sine: poisson2d_sine, 25000 steps, errors vs exact) with the initial widths scaled by
(1 + k*1e-7): final errors for (a) the fused path with --graph, (b) the fused path, eager loop,
(c) the reference's dense eager algorithm (oracle/torch_port.py) on the same GPU.  Shows how
chaotic the converged error is and whether the implementations draw from the same distribution.

    python tests/perf/converged_ensemble.py [irreg|sine] [graph] [eager] [dense] [k ...]"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import torch_port  # noqa: E402
from spinn_b200 import common  # noqa: E402
from spinn_b200.problems import accuracy, poisson2d_irreg_dom as irr, poisson2d_sine as sine  # noqa: E402
from spinn_b200.spinn2d import Plotter2D  # noqa: E402
from spinn_b200.spinn2d import App2D  # noqa: E402


CONFIG = {"irreg": (irr, irr.Poisson2D, irr.PointCloud, ["--lr", "1e-3"]),
          "sine": (sine, sine.Poisson2D, Plotter2D,
                   ["--nodes", "100", "--samples", "400", "--n-train", "25000", "--lr", "1e-3", "--tol", "1e-3"])}
WHICH = ["irreg"]


def run(kind, k):
    np.random.seed(0)
    torch.manual_seed(0)
    mod, pde_cls, plot_cls, base = CONFIG[WHICH[0]]
    if kind == "dense":
        app = App2D(pde_cls=pde_cls, nn_cls=torch_port.DenseSPINN2D, plotter_cls=plot_cls)
    else:
        app = mod.make_app()
    extra = ["--graph"] if kind == "graph" else []
    p = app.setup_argparse().parse_args(base + ["--gpu", "--n-skip", "100000"] + extra)
    common.device("cuda")
    pde = app.pde_cls.from_args(p)
    nn = app.nn_cls.from_args(pde, app._get_activation(p), p).to("cuda")
    with torch.no_grad():
        nn.layer1.h.mul_(1.0 + k * 1e-7)
    plotter = app.plotter_cls.from_args(pde, nn, p)
    solver = app.optimizer.from_args(pde, nn, plotter, p)
    with contextlib.redirect_stdout(io.StringIO()):
        solver.solve()
    return solver.loss[-1], (accuracy.fem_errors(nn) if WHICH[0] == "irreg" else plotter.get_error())


if __name__ == "__main__":
    WHICH[0] = ([a for a in sys.argv[1:] if a in CONFIG] or ["irreg"])[0]
    kinds = [a for a in sys.argv[1:] if a in ("graph", "eager", "dense")] or ["graph", "eager", "dense"]
    ks = [int(a) for a in sys.argv[1:] if a.lstrip("-").isdigit()] or list(range(-3, 4))
    summary = {}
    for kind in kinds:
        rows = []
        for k in ks:
            loss, e = run(kind, k)
            rows.append((loss,) + tuple(e))
            print(kind, k, f"loss {loss:.4e}  L1 {e[0]:.3e} L2 {e[1]:.3e} Linf {e[2]:.3e}", flush=True)
        a = np.asarray(rows)
        summary[kind] = {"n": len(ks), "median": np.median(a, 0).tolist(), "q25": np.quantile(a, 0.25, 0).tolist(),
                         "q75": np.quantile(a, 0.75, 0).tolist(), "min": a.min(0).tolist(), "max": a.max(0).tolist()}
    import json
    print("SUMMARY", WHICH[0], "(columns: final loss, L1, L2, Linf)", json.dumps(summary))
