#!/usr/bin/env python
"""Time the fast-kernel variants (SPINN_TUNE) on the BASELINE workload on this GPU.
    python tools/tune.py [--samples N] [--fwd ids] [--bwd ids] [--waves list]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinn_b200 import synthetic  # noqa: E402
from spinn_b200.fused_step import PoissonInteriorStep  # noqa: E402


def time_ms(fn, reps=4):
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


MASK = [0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=4194304)
    ap.add_argument("--fwd", default="0,1,2,3")
    ap.add_argument("--bwd", default="0,1,2,3")
    ap.add_argument("--waves", default="2,4,8")
    ap.add_argument("--mask", type=lambda v: int(v, 0), default=0)
    a = ap.parse_args()
    MASK[0] = a.mask
    pde, nn = synthetic.make_state(samples=a.samples, device="cuda")
    xs, ys = (t.detach() for t in pde.p_samples)
    N, M = xs.numel(), nn.layer1.n
    step = PoissonInteriorStep(nn, xs, ys, structured=a.mask != 0)
    step.lib.spinn_set_tuning(b"fwd=0,bwd=0,waves=2")
    ref = step.run(all_reduce=False, separate=True).clone()
    ref_jets = step.jets.clone()
    lib, P = step.lib, (lambda t: None if t is None else t.data_ptr())
    l1, l2 = nn.layer1, nn.layer2
    S = lambda: torch.cuda.current_stream().cuda_stream

    def fwd():
        assert lib.spinn2d_jets_fwd(0, 2, N, step.Mf, step.Mx, 1, P(xs), P(ys), P(l1.x), P(l1.y), P(l1.xf),
                                    P(l1.yf), P(l1.h), 0, P(l2.weight), P(l2.bias), P(step.jets),
                                    P(step.ws_f), step.ws_f.numel(), S()) == 0

    def bwd():
        p = step.pack
        assert lib.spinn2d_jets_bwd(0, 2, MASK[0], N, step.Mf, step.Mx, 1, P(xs), P(ys), P(l1.x), P(l1.y), P(l1.xf),
                                    P(l1.yf), P(l1.h), 0, P(l2.weight), P(step.gjets), P(p.view("d_x")),
                                    P(p.view("d_y")), P(p.view("d_h")), P(p.view("d_W")), P(p.view("d_b")),
                                    P(step.ws_b), step.ws_b.numel(), S()) == 0

    print(f"N={N} M={M}")
    for v in [int(t) for t in a.fwd.split(",") if t]:
        lib.spinn_set_tuning(f"fwd={v},bwd=0,waves=2".encode())
        step.jets.zero_()
        ms = time_ms(fwd)
        err = float((step.jets - ref_jets).abs().max() / ref_jets.abs().max())
        print(f"fwd variant {v}: {ms:8.3f} ms   {N * M * 23 / ms / 1e9:7.2f} TFLOP/s   maxdiff {err:.2e}", flush=True)
    for v in [int(t) for t in a.bwd.split(",") if t]:
        for w in [int(t) for t in a.waves.split(",") if t]:
            lib.spinn_set_tuning(f"fwd=0,bwd={v},waves={w}".encode())
            step.pack.flat.zero_()
            ms = time_ms(bwd)
            n = step.pack.slices["loss"][0]
            err = float((step.pack.flat[:n] - ref[:n]).abs().max() / ref[:n].abs().max())
            flop = {0x18: 25, 0x08: 25, 0x10: 25, 0x01: 15}.get(a.mask, 53)    # per pair, by backward mode
            print(f"bwd variant {v} waves {w}: {ms:8.3f} ms   {N * M * flop / ms / 1e9:7.2f} TFLOP/s "
                  f"({flop} flop/pair)   maxdiff {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
