#!/usr/bin/env python
"""Throughput of every kernel variant (dim, kernel, K, level) through the C ABI:
point-node evaluations/s for forward and backward separately.

    python tools/bench_variants.py [--points 1048576] [--nodes 1024]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinn_b200 import ops  # noqa: E402


def time_ms(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1048576)
    ap.add_argument("--nodes", type=int, default=1024)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    N, M = a.points, a.nodes
    r = lambda *s: torch.rand(*s, generator=g).to(dev)
    be = ops.CudaBackend()
    rows = []
    for dim in (2, 1):
        for kind in (0, 1):
            for K in (1, 3):
                x, y = r(N), (r(N) if dim == 2 else None)
                Mf = M - M // 8
                xf, yf = r(Mf), (r(Mf) if dim == 2 else None)
                xx, yx = r(M - Mf), (r(M - Mf) if dim == 2 else None)
                h = (1.0 / np.sqrt(M) if dim == 2 else 1.0 / M) * (1 + 0.3 * r(M))
                W, b = 0.3 * (r(K, M) - 0.5), 0.1 * r(K)
                for level in ((2, 3) if dim == 2 else (2,)):
                    nj = ops.num_jets(dim, level)
                    gj = (r(nj, N, K) - 0.5).contiguous()
                    f = lambda: be.forward(dim, kind, level, x, y, xf, yf, xx, yx, h, W, b)
                    bw = lambda: be.backward(dim, kind, level, x, y, xf, yf, xx, yx, h, W, gj, True)
                    f(); bw()
                    mf, mb = time_ms(f), time_ms(bw)
                    row = {"dim": dim, "kind": ["gaussian", "softplus"][kind], "K": K, "level": level,
                           "fwd_ms": mf, "bwd_ms": mb, "fwd_evals_per_s": N * M / mf * 1e3,
                           "bwd_evals_per_s": N * M / mb * 1e3,
                           "fwdbwd_evals_per_s": N * M / (mf + mb) * 1e3}
                    rows.append(row)
                    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
