#!/usr/bin/env python
"""Train-step time (ms per opt.step(closure)) of the four BASELINE configs on one GPU:
the fused B200 path vs the reference's dense eager algorithm on the SAME GPU
(oracle/torch_port.py:DenseSPINN*, driven by the same problem classes and Optimizer).

    python tests/perf/train_step_bench.py [--steps 200]

(lives under tests/ because it times the oracle's dense modules as the comparison baseline)
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import torch_port  # noqa: E402
from spinn_b200 import common  # noqa: E402
from spinn_b200.problems import cavity, ode3, poisson2d_irreg_dom, poisson2d_sine  # noqa: E402
from spinn_b200.spinn1d import App1D, Plotter1D  # noqa: E402
from spinn_b200.spinn2d import App2D, Plotter2D  # noqa: E402

CONFIGS = {
    "ode3": (ode3, ["--nodes", "3", "--samples", "60", "--lr", "1e-3", "--tol", "1e-30"]),
    "poisson2d_sine": (poisson2d_sine, ["--nodes", "100", "--samples", "400", "--lr", "1e-3", "--tol", "1e-30"]),
    "poisson2d_irreg_dom": (poisson2d_irreg_dom, ["--lr", "1e-3"]),
    "cavity": (cavity, ["--nodes", "100", "--samples", "2500", "--sample-frac", "0.1", "--lr", "5e-4"]),
}


def run(app, args, steps, extra=()):
    np.random.seed(0)
    torch.manual_seed(0)
    a = args + ["--gpu", "--n-train", str(steps), "--n-skip", str(10 ** 9)] + list(extra)
    with contextlib.redirect_stdout(io.StringIO()):
        app.run(args=a)                      # includes warm-up effects; time a second solve below
        torch.cuda.synchronize()
        app.solver.n_train = steps
        t0 = time.perf_counter()
        app.solver.solve()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return 1e3 * dt / steps, app.solver.loss[-1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    out = {}
    for name, (mod, args) in CONFIGS.items():
        fused = mod.make_app()
        ms_f, loss_f = run(fused, args, a.steps)
        # graph mode: the step captured by the first solve() is replayed by the timed one
        ms_g, loss_g = run(mod.make_app(), args, 10 * a.steps, extra=["--graph"])
        is1d = name == "ode3"
        pde_cls, plot_cls = fused.pde_cls, fused.plotter_cls
        dense_cls = torch_port.DenseSPINN1D if is1d else torch_port.DenseSPINN2D
        dense = (App1D if is1d else App2D)(pde_cls=pde_cls, nn_cls=dense_cls, plotter_cls=plot_cls)
        ms_d, loss_d = run(dense, args, a.steps)
        out[name] = {"fused_ms_per_step": ms_f, "fused_graphed_ms_per_step": ms_g,
                     "dense_eager_same_gpu_ms_per_step": ms_d,
                     "speedup_eager": ms_d / ms_f, "speedup_graphed": ms_d / ms_g,
                     "loss_fused": loss_f, "loss_graphed": loss_g, "loss_dense": loss_d}
        print(name, json.dumps(out[name]), flush=True)
    out["heat1d_fd_spinn"] = fd_run(a.steps)
    print("heat1d_fd_spinn", json.dumps(out["heat1d_fd_spinn"]), flush=True)
    print(json.dumps(out))


def fd_run(steps):
    """FD-SPINN outer loop (10 time levels x `steps` training steps, no tolerance stop): wall time
    of the whole App.run per training step, fused eager / fused --graph / dense eager."""
    from spinn_b200.problems import heat1d_fd_spinn as fd
    from spinn_b200.problems.fd_spinn1d_base import AppFD1D
    args = ["--nodes", "10", "--samples", "100", "--dt", "0.01", "-T", "0.1", "--lr", "1e-3", "--tol", "1e-30",
            "--gpu", "--n-train", str(steps), "--n-skip", str(10 ** 9)]
    res = {}
    dense = AppFD1D(pde_cls=fd.Heat1D, nn_cls=torch_port.DenseSPINN1D, plotter_cls=fd.FDPlotter1D)
    mk_dense = lambda: AppFD1D(pde_cls=fd.Heat1D, nn_cls=torch_port.DenseSPINN1D, plotter_cls=fd.FDPlotter1D)
    with contextlib.redirect_stdout(io.StringIO()):          # one-time costs (module loads) out of the way
        for mk, extra in ((fd.make_app, []), (fd.make_app, ["--graph"]), (mk_dense, [])):
            mk().run(args=["--nodes", "10", "--samples", "100", "--dt", "0.01", "-T", "0.02", "--tol", "1e-30",
                           "--gpu", "--n-train", "5"] + extra)
    for tag, app, extra in (("fused", fd.make_app(), []), ("fused_graphed", fd.make_app(), ["--graph"]),
                            ("dense_eager_same_gpu", dense, [])):
        np.random.seed(0)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            app.run(args=args + extra)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res[tag + "_ms_per_step"] = 1e3 * dt / (10 * steps)
        res["linf_" + tag] = float(app.plotter.get_error()[2])
    return res


if __name__ == "__main__":
    main()
