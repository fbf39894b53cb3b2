#!/usr/bin/env python
"""Parity report on the GPU: every reference-generated op fixture (tests/golden/op_*.npz) through
the C ABI, compared with the reference's own fp64 evaluation, next to the reference's own fp32
result on the same inputs.  Prints one JSON object (-> profiles/rNN_parity_report.json).

Two metrics per quantity: `mixed` = max |new-ref| / (|ref| + 1e-2 max|ref|) (tests/helpers.py, the
bar the tests assert) and `rel_p99` = 99th percentile of the strict per-element relative error
|new-ref|/|ref| (the literal "1e-5 relative per element" reading; its maximum is dominated by zero
crossings for the reference's fp32 path and for ours alike).

    python tests/perf/parity_report.py > gpurun_out/parity_report.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from helpers import OP_CASES_1D, OP_CASES_2D, load_case, mixed_err  # noqa: E402
from spinn_b200 import ops  # noqa: E402

KIND = {"gaussian": 0, "softplus": 1}
DEV = torch.device("cuda", 0)
T = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


def rel_p99(new, ref):
    new, ref = np.asarray(new, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    nz = np.abs(ref) > 0
    if not nz.any():
        return 0.0
    return float(np.percentile(np.abs(new[nz] - ref[nz]) / np.abs(ref[nz]), 99))


def both(new, ref32, ref64):
    return {"ours_mixed": mixed_err(new, ref64), "reference_fp32_mixed": mixed_err(ref32, ref64),
            "ours_rel_p99": rel_p99(new, ref64), "reference_fp32_rel_p99": rel_p99(ref32, ref64)}


def main():
    be = ops._BACKEND
    assert isinstance(be, ops.CudaBackend)
    out = {}
    for name in OP_CASES_2D:
        c = load_case(name)
        args = (T(c["x"]), T(c["y"]), T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]),
                T(c["h"]).reshape(-1), T(c["W"]), T(c["b"]))
        kind = KIND[str(c["kind"])]
        jets = be.forward(2, kind, 2, *args).cpu().numpy()
        grads = [g.cpu().numpy() for g in be.backward(2, kind, 2, *args[:-1], T(c["gjets"][:5]), True)]
        r = {"N": int(len(c["x"])), "M": int(len(c["xf"]) + len(c["xfix"])), "K": int(c["K"]), "kind": str(c["kind"])}
        for a, nm in enumerate(("u", "u_x", "u_y", "u_xx", "u_yy")):
            r[nm] = both(jets[a], c["jets_f32"][a], c["jets_f64"][a])
        for g, key in zip(grads, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
            r[key] = both(g, c[key + "_f32"], c[key + "_f64"])
        out[name] = r
    for name in OP_CASES_1D:
        c = load_case(name)
        args = (T(c["x"]), None, T(c["xf"]), None, T(c["xfix"]), None, T(c["h"]), T(c["W"]), T(c["b"]))
        kind = KIND[str(c["kind"])]
        jets = be.forward(1, kind, 2, *args).cpu().numpy()
        d_xc, _, d_h, d_W, d_b = be.backward(1, kind, 2, *args[:-1], T(c["gjets"][:3]), True)
        r = {"N": int(len(c["x"])), "M": int(len(c["xf"]) + len(c["xfix"])), "K": int(c["K"]), "kind": str(c["kind"])}
        for a, nm in enumerate(("u", "u_x", "u_xx")):
            r[nm] = both(jets[a], c["jets_f32"][a], c["jets_f64"][a])
        for g, key in ((d_xc, "d_xc"), (d_h, "d_h"), (d_W, "d_W"), (d_b, "d_b")):
            r[key] = both(g.cpu().numpy(), c[key + "_f32"], c[key + "_f64"])
        out[name] = r
    worst = {"ours_mixed": 0.0, "reference_fp32_mixed": 0.0, "ours_rel_p99": 0.0, "reference_fp32_rel_p99": 0.0}
    for r in out.values():
        for v in r.values():
            if isinstance(v, dict):
                for k in worst:
                    worst[k] = max(worst[k], v[k])
    out["worst_over_all_cases"] = worst
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
