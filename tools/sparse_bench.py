#!/usr/bin/env python
"""Compact-support (cut-off) step vs the dense step on the BASELINE workload: time, evaluated
pair fraction, deviation.   python tools/sparse_bench.py [--cutoffs 6,7,8] [--samples N]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinn_b200 import synthetic  # noqa: E402
from spinn_b200.fused_step import PoissonInteriorStep  # noqa: E402


def time_ms(fn, reps=5):
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
    ap.add_argument("--cutoffs", default="6,7,8")
    ap.add_argument("--samples", type=int, default=4194304)
    a = ap.parse_args()
    pde, nn = synthetic.make_state(samples=a.samples, device="cuda")
    xs, ys = (t.detach() for t in pde.p_samples)
    N, M = xs.numel(), nn.layer1.n
    dense = PoissonInteriorStep(nn, xs, ys)
    ref = dense.run(all_reduce=False).clone()
    ms_dense = time_ms(lambda: dense.run(all_reduce=False))
    n = dense.pack.slices["loss"][0]
    print(json.dumps({"mode": "dense", "ms_per_step": ms_dense, "dense_pairs_per_s": N * M / ms_dense * 1e3}))
    for c in [float(t) for t in a.cutoffs.split(",")]:
        st = PoissonInteriorStep(nn, xs, ys, cutoff=c)
        out = st.run(all_reduce=False).clone()
        st.stats.zero_()
        st.run(all_reduce=False)
        torch.cuda.synchronize()
        frac = [float(v) / (N * M) for v in st.stats]
        ms = time_ms(lambda: st.run(all_reduce=False))
        print(json.dumps({"mode": "compact_support", "cutoff_widths": c, "ms_per_step": ms,
                          "speedup_vs_dense": ms_dense / ms, "evaluated_pair_fraction_fwd": frac[0],
                          "evaluated_pair_fraction_bwd": frac[1],
                          "evaluated_pairs_per_s": (frac[0] + frac[1]) / 2 * N * M / ms * 1e3,
                          "max_rel_grad_diff_vs_dense": float((out[:n] - ref[:n]).abs().max() / ref[:n].abs().max()),
                          "rel_loss_diff_vs_dense": abs(float(out[n]) - float(ref[n])) / abs(float(ref[n]))}),
              flush=True)


if __name__ == "__main__":
    main()
