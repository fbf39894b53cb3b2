#!/usr/bin/env python
"""Benchmark of the SPINN hot path on B200: point-node evaluations/s of the fused
residual forward+backward on the BASELINE 'synthetic scaled Poisson2D' workload
(4 194 304 collocation points x 4096 nodes, Gaussian kernel, fp32), point-sharded over
the GPUs of one box (strong scaling: the global point set is fixed).

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference's CPU algorithm (oracle port)

One "step" = interior forward jets (u, ux, uy, uxx, uyy) -> residual/loss/upstream
gradients -> backward to all node parameters -> one all-reduce of the packed gradient.
`value` times it with inputs resident in HBM; `e2e` times the same quantity through
the public plug-in API (SPINN2D.forward + the problem class's autograd recipe +
loss.backward) with points coming from pinned host memory and the gradient+loss going
back to the host every step.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "point-node evals/sec for fused residual fwd+bwd"
UNIT = "point-node evals/s"
# algorithmic flops per point-node pair (FMA = 2).  General path: SURVEY.md 8(d): 23 fwd + 53 bwd
# (13 recompute + 40) = 76.  The benchmarked Poisson residual only has u_xx,u_yy upstream
# gradients, so its backward runs the structured variant: 8 (geometry) + 17 = 25 flops in 16
# instructions (spinn_math.cuh:g2_bwd_pair_lapl) -> 48 per pair.
SMS, LANES, MUFU_LANES, NOMINAL_MHZ = 148, 128, 16, 1965.0
# Executed work per point-node pair of each kernel, counted in the SASS of the inner loops
# (tools/sass_census.py -> profiles/r02_sass_inner_loops.txt): flop (FMA = 2), FP32-pipe instruction
# slots (a packed FFMA2/FMUL2/FADD2 = one slot per lane-pair = 1 per pair), MUFU operations.
KERNEL_COST = {
    # flop: ALGORITHMIC count the roofline uses (Gaussian: SURVEY.md 8d, 23 forward + 53 general backward; the
    # structured backward and the softplus kernels have no survey figure, so theirs is the executed count);
    # flop_sass: flops the inner loop actually executes (the Gaussian kernels fold scalings into per-node
    # weights and need 21 / 50 where the survey counts 23 / 53).  tests/test_sass_contract.py pins
    # instr, mufu and flop_sass against the SASS of the shipped library.
    ("gaussian", "fwd"): dict(flop=23.0, flop_sass=21.0, instr=14, mufu=1),
    ("gaussian", "bwd_structured"): dict(flop=25.0, flop_sass=25.0, instr=16, mufu=1),
    ("gaussian", "bwd_general"): dict(flop=53.0, flop_sass=50.0, instr=32, mufu=1),
    ("softplus", "fwd"): dict(flop=49.0, flop_sass=49.0, instr=34, mufu=5),
    ("softplus", "bwd_structured"): dict(flop=74.0, flop_sass=74.0, instr=51, mufu=4),
    ("softplus", "bwd_general"): dict(flop=102.0, flop_sass=102.0, instr=67, mufu=5),
}
FLOP_FWD, FLOP_BWD_GENERAL, FLOP_BWD_STRUCTURED = 23.0, 53.0, 25.0      # Gaussian (23 + 53 = 76)
NCU_KERNELS_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_ncu_kernels.json")


def kernel_report(kind, which, pairs, ms, clk_mhz):
    """Achieved rates of one kernel launch against the nominal pipes at the SM clock sampled during
    the timed region: FP32 148 SM x 128 lanes x 2 flop, MUFU 148 SM x 16 lanes."""
    c = KERNEL_COST[(kind, which)]
    t = ms * 1e-3
    fp32_peak = SMS * LANES * 2 * clk_mhz * 1e6
    lane_rate = SMS * LANES * clk_mhz * 1e6
    mufu_rate = SMS * MUFU_LANES * clk_mhz * 1e6
    slots = pairs * c["instr"] / t / lane_rate
    mufu = pairs * c["mufu"] / t / mufu_rate
    return {"ms": ms, "pairs_per_s": pairs / t, "tflops": pairs * c["flop"] / t / 1e12,
            "frac": pairs * c["flop"] / t / fp32_peak, "flop_per_pair": c["flop"],
            "flop_per_pair_executed": c["flop_sass"],
            "fp32_instr_per_pair": c["instr"], "fp32_slot_frac": slots,
            "mufu_per_pair": c["mufu"], "mufu_frac": mufu,
            "binding_pipe": "mufu" if mufu > slots else "fp32", "binding_pipe_frac": max(mufu, slots)}


def ncu_traffic(kernel_prefix, N, M):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture of this workload (tools/ncu_summary.py json); None if the file
    does not cover it."""
    try:
        d = json.load(open(NCU_KERNELS_JSON))
        if d.get("points") != N or d.get("nodes") != M:
            return None
        for name, k in d["kernels"].items():
            if name.startswith(kernel_prefix):
                return {"bytes": k["dram_bytes_read"] + k["dram_bytes_write"], "kernel": name,
                        "source": os.path.relpath(NCU_KERNELS_JSON, os.path.dirname(NCU_KERNELS_JSON) + "/..")}
    except Exception:
        pass
    return None


def workload_name(kind, n_points, n_nodes):
    return (f"synthetic scaled Poisson2D: {n_points} collocation points x {n_nodes} nodes, "
            f"{kind} kernel, K=1, 5 jets (u,ux,uy,uxx,uyy)")


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--samples", type=int, default=4194304)
    p.add_argument("--nodes", type=int, default=3600)
    p.add_argument("--b-nodes", type=int, default=124)
    p.add_argument("--kind", default="gaussian", choices=["gaussian", "softplus"])
    p.add_argument("--cpu-chunk", type=int, default=16384)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-softplus", action="store_true", help="skip the softplus-hat leg of the default line")
    p.add_argument("--general-backward", action="store_true",
                   help="time the general 5-row backward instead of the structured (u_xx,u_yy) one")
    return p.parse_args()


# ------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# the reference ITSELF (baseline/_ref/code, unmodified) -- cpu_baseline, --impl reference and the
# same-GPU eager leg.  Falls back to the oracle port only when baseline/_ref is absent.
# ------------------------------------------------------------------------------------------
def _perturb_reference_(nn, seed=0):
    """The seeded parameter state of SURVEY.md 8(d) (same draws as spinn_b200.synthetic.perturb_,
    restated here so that the reference arm imports nothing of spinn_b200)."""
    g = torch.Generator().manual_seed(seed)
    l1, l2 = nn.layer1, nn.layer2
    dev = l2.weight.device
    with torch.no_grad():
        l2.weight.copy_((0.3 * torch.randn(l2.weight.shape, generator=g)).to(dev))
        l1.h.mul_((1.0 + 0.3 * torch.rand(l1.h.shape, generator=g)).to(dev))
        l1.x.add_((0.01 * torch.randn(l1.x.shape, generator=g)).to(dev))
        l1.y.add_((0.01 * torch.randn(l1.y.shape, generator=g)).to(dev))
        l2.bias.copy_((0.1 * torch.randn(l2.bias.shape, generator=g)).to(dev))


def host_threads():
    """All host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to its children)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def reference_run(args, device, chunk, steps, warmup):
    """One step = the reference's own interior path on `chunk` of the workload's points:
    SPINN2D.forward (code/spinn2d.py:169-179) -> RegularPDE._get_residue / _compute_derivatives
    (code/pde2d_base.py:46-62,132-137) -> Poisson2D.pde (code/poisson2d_sine.py:16-18) ->
    interior_loss (pde2d_base.py:139-141) -> backward (common.py:246), executed from
    baseline/_ref/code.  Rate is size-independent."""
    from baseline import reference_loader as rl
    ref = rl.load("common", "spinn2d", "pde2d_base", "poisson2d_sine")
    ref.common.device(device)
    pde = ref.poisson2d_sine.Poisson2D(args.nodes, args.samples, args.b_nodes, None)
    act = ref.spinn2d.gaussian if args.kind == "gaussian" else ref.spinn2d.SoftPlus()
    nn = ref.spinn2d.SPINN2D(pde, act).to(ref.common.device())
    _perturb_reference_(nn)
    xs, ys = pde.p_samples
    n_all = xs.numel()
    chunk = min(chunk, n_all)
    stride = max(1, n_all // chunk)                        # spread the sample over the whole grid
    pde.p_samples = tuple(t.detach()[::stride][:chunk].clone().requires_grad_(True) for t in (xs, ys))
    n = pde.p_samples[0].numel()
    M = nn.layer1.n
    sync = torch.cuda.synchronize if torch.device(device).type == "cuda" else (lambda: None)

    def one():
        nn.zero_grad()
        for t in pde.p_samples:
            t.grad = None
        loss = pde.interior_loss(nn)
        loss.backward()
        return loss

    for _ in range(warmup):
        one()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = one()
    sync()
    dt = time.perf_counter() - t0
    return {"value": n * M * steps / dt, "ms_per_step": 1e3 * dt / steps, "points": n, "nodes": M,
            "cores": torch.get_num_threads(), "nproc": os.cpu_count(), "loss": float(loss.detach()), "kind": "reference",
            "what": "baseline/_ref/code (unmodified reference): SPINN2D.forward + RegularPDE._compute_derivatives "
                    "(3x autograd.grad) + Poisson2D.pde + interior_loss + backward, dense (N,M) fp32 eager"}


def port_run(args, steps, warmup):
    """Oracle port (oracle/torch_port.py) -- only when baseline/_ref did not travel."""
    from oracle import torch_port as tp
    from spinn_b200 import common, synthetic
    common.device("cpu")
    n_side = int(round(np.sqrt(args.cpu_chunk)))
    pde, nn = synthetic.make_state(args.nodes, args.b_nodes, n_side * n_side, args.kind, device="cpu")
    xs, ys = pde.p_samples
    n = xs.numel()
    l1, l2 = nn.layer1, nn.layer2
    M = l1.n

    def one():
        for t in (xs, ys, l1.x, l1.y, l1.h, l2.weight, l2.bias):
            t.grad = None
        xc, yc = l1.centers()
        return tp.poisson_interior_step(args.kind, xs, ys, xc, yc, l1.h, l2.weight, l2.bias, float(args.samples))

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {"value": n * M * steps / dt, "ms_per_step": 1e3 * dt / steps, "points": n, "nodes": M,
            "cores": torch.get_num_threads(), "nproc": os.cpu_count(), "kind": "port",
            "what": "oracle/torch_port.py (dense eager port; baseline/_ref absent)"}


def cpu_reference_run(args, steps, warmup):
    from baseline import reference_loader as rl
    host_threads()
    if rl.available():
        return reference_run(args, "cpu", args.cpu_chunk, steps, warmup)
    return port_run(args, steps, warmup)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args, args.steps, args.warmup)
    sample = (f"{r['points']} of {args.samples} points x {r['nodes']} nodes per step, {r['what']}; "
              f"{r['cores']} torch threads of {r['nproc']} cpus")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.kind, args.samples, r["nodes"]), "points": args.samples,
                   "nodes": r["nodes"], "kind": args.kind,
                   "sample": f"each step = {r['points']} of the {args.samples} points (bounded CPU sample)"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def microbench_peaks(lib, dev):
    """FP32 and MUFU.EX2 instruction rates measured on this GPU right now (lane-ops/s) with the
    probes of spinn_b200/csrc/microbench.cu (register operands, 64 resident warps per SM)."""
    import ctypes
    sink = torch.zeros(4, device=dev)
    src = (1.0 + 1e-3 * torch.rand(64, device=dev)).float()
    out = {}
    s = torch.cuda.current_stream().cuda_stream
    probes = ((1, "ffma_3reg"), (0, "ffma_shared"), (2, "ffma2_shared"), (3, "ffma2_3pair"),
              (4, "ffma2_bcast"), (10, "ex2"))
    for which, name in probes:
        ops = ctypes.c_double(0.0)
        best = None
        for rep in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            rc = lib.spinn_microbench(which, 2048, SMS * 8, 256, src.data_ptr(), sink.data_ptr(),
                                      ctypes.byref(ops), s)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        out[name] = ops.value / (best * 1e-3)
    return out


def run_ours(args):
    from spinn_b200 import _cabi, common, distributed, synthetic
    from spinn_b200.fused_step import PoissonInteriorStep
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the spinn_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU algorithm)")
    rank, ws, local = distributed.init()
    if ws != args.gpus:
        if ws == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _cabi.load()
    common.device(dev)

    pde, nn = synthetic.make_state(args.nodes, args.b_nodes, args.samples, args.kind, device=dev)
    xs_all, ys_all = (t.detach() for t in pde.p_samples)
    N = xs_all.numel()
    M = nn.layer1.n
    lo, hi = distributed.shard_range(N, rank, ws)
    xs, ys = xs_all[lo:hi].contiguous(), ys_all[lo:hi].contiguous()
    del xs_all, ys_all
    n_loc = hi - lo
    step = PoissonInteriorStep(nn, xs, ys, n_total=N, structured=not args.general_backward)

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- device-resident timing ("value") ----
    for _ in range(args.warmup):
        step.run()
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step.run()
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = N * M / (ms_step * 1e-3)
    loss = float(step.pack.view("loss")[0])

    # ---- per-kernel timing for the roofline (fwd, bwd incl. their tiny pack kernels) ----
    def time_call(fn, reps=5):
        best = None
        for _ in range(reps):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            torch.cuda.synchronize()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        return best

    l1, l2 = nn.layer1, nn.layer2
    P = lambda t: None if t is None else t.data_ptr()
    S = lambda: torch.cuda.current_stream().cuda_stream
    hs = int(l1.h.numel() == 1)

    def fwd_only():
        rc = lib.spinn2d_jets_fwd(nn.kind, 2, n_loc, step.Mf, step.Mx, 1, P(xs), P(ys), P(l1.x), P(l1.y),
                                  P(l1.xf), P(l1.yf), P(l1.h), hs, P(l2.weight), P(l2.bias), P(step.jets),
                                  P(step.ws_f), step.ws_f.numel(), S())
        assert rc == 0

    def bwd_only():
        p = step.pack
        rc = lib.spinn2d_jets_bwd(nn.kind, 2, step.upstream_mask, n_loc, step.Mf, step.Mx, 1, P(xs), P(ys), P(l1.x), P(l1.y),
                                  P(l1.xf), P(l1.yf), P(l1.h), hs, P(l2.weight), P(step.gjets),
                                  P(p.view("d_x")), P(p.view("d_y")), P(p.view("d_h")), P(p.view("d_W")),
                                  P(p.view("d_b")), P(step.ws_b), step.ws_b.numel(), S())
        assert rc == 0

    step.run(all_reduce=False, separate=True)       # leaves dLoss/dJets in step.gjets for the per-kernel timing
    ms_fwd = max_over_ranks(time_call(fwd_only))
    ms_bwd = max_over_ranks(time_call(bwd_only))
    pairs_loc = float(n_loc) * M

    # ---- end to end through the public API, host buffers ----
    e2e = None
    if not args.no_e2e:
        xh = xs.cpu().pin_memory()
        yh = ys.cpu().pin_memory()
        pack = step.pack
        out_host = torch.empty(pack.flat.numel(), dtype=torch.float32).pin_memory()
        params = [l1.x, l1.y, l1.h, l2.weight, l2.bias]

        # Host->device copies run on a side stream, one step ahead (what a data loader does);
        # every step still uploads its own inputs and downloads its own result inside the
        # timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        staged = {}

        def upload():
            with torch.cuda.stream(copy_stream):
                x = xh.to(dev, non_blocking=True)
                y = yh.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            staged["next"] = (x, y, ev)

        def e2e_step():
            if "next" not in staged:
                upload()
            x, y, ev = staged.pop("next")
            torch.cuda.current_stream().wait_event(ev)
            x.record_stream(torch.cuda.current_stream())
            y.record_stream(torch.cuda.current_stream())
            upload()                                                   # inputs of the following step
            x.requires_grad_(True)
            y.requires_grad_(True)
            for p_ in params:
                p_.grad = None
            u = nn(x, y)                                               # public plug-in call
            u, ux, uy, uxx, uyy = pde._compute_derivatives(u, x, y)    # the problem class's ag.grad recipe
            res = pde.pde(x, y, u, ux, uy, uxx, uyy)
            lossv = (res ** 2).sum() / N
            lossv.backward()
            flat = pack.flat
            flat.zero_()
            pack.view("d_x").copy_(l1.x.grad)
            pack.view("d_y").copy_(l1.y.grad)
            pack.view("d_h").copy_(l1.h.grad.reshape(-1))
            pack.view("d_W").copy_(l2.weight.grad.reshape(-1))
            pack.view("d_b").copy_(l2.bias.grad)
            pack.view("loss").copy_(lossv.detach().reshape(1))
            pack.all_reduce()
            out_host.copy_(flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(max(2, args.warmup)):
            e2e_step()
        barrier()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(args.steps):
            e2e_step()
        b.record()
        barrier()
        ms_e2e_eager = max_over_ranks(a.elapsed_time(b)) / args.steps
        loss_eager = float(out_host[pack.slices["loss"][0]])
        staged.clear()

        # The same step as a CUDA graph (what `--graph` does for the training loop, spinn_b200/
        # graphed.py): identical public calls, captured once per input buffer; every timed step
        # still uploads its inputs from pinned host memory (prefetched into the OTHER static
        # buffer on the copy stream), replays, downloads the packed result and synchronises.
        def body(x, y):
            for p_ in params:
                p_.grad = None
            u = nn(x, y)
            u, ux, uy, uxx, uyy = pde._compute_derivatives(u, x, y)
            res = pde.pde(x, y, u, ux, uy, uxx, uyy)
            lossv = (res ** 2).sum() / N
            lossv.backward(inputs=params)
            flat = pack.flat
            flat.zero_()
            pack.view("d_x").copy_(l1.x.grad)
            pack.view("d_y").copy_(l1.y.grad)
            pack.view("d_h").copy_(l1.h.grad.reshape(-1))
            pack.view("d_W").copy_(l2.weight.grad.reshape(-1))
            pack.view("d_b").copy_(l2.bias.grad)
            pack.view("loss").copy_(lossv.detach().reshape(1))
            pack.all_reduce()

        bufs = [(torch.zeros_like(xs).requires_grad_(True), torch.zeros_like(ys).requires_grad_(True))
                for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for x, y in bufs:
                with torch.no_grad():
                    x.copy_(xs)
                    y.copy_(ys)
                body(x, y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for x, y in bufs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=graphs[0].pool() if graphs else None):
                body(x, y)
            graphs.append(g)

        def prefetch(k):
            with torch.cuda.stream(copy_stream), torch.no_grad():
                bufs[k][0].copy_(xh, non_blocking=True)
                bufs[k][1].copy_(yh, non_blocking=True)
                evs[k].record(copy_stream)

        turn = [0]

        def e2e_graph_step():
            k = turn[0] & 1
            turn[0] += 1
            prefetch(1 - k)                      # inputs of the following step (its buffer is idle)
            torch.cuda.current_stream().wait_event(evs[k])
            graphs[k].replay()
            out_host.copy_(pack.flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        prefetch(0)
        for _ in range(max(2, args.warmup)):
            e2e_graph_step()
        barrier()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(args.steps):
            e2e_graph_step()
        b.record()
        barrier()
        ms_e2e = max_over_ranks(a.elapsed_time(b)) / args.steps
        e2e = {"value": N * M / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": 2 * 4 * n_loc, "d2h_bytes_per_step": pack.nbytes(),
               "loss": float(out_host[pack.slices["loss"][0]]),
               "api": "SPINN2D.forward -> RegularPDE._compute_derivatives -> pde() -> loss.backward(), "
                      "captured once as a CUDA graph and replayed per step (as --graph does)",
               "eager_loop": {"value": N * M / (ms_e2e_eager * 1e-3), "ms_per_step": ms_e2e_eager,
                              "loss": loss_eager,
                              "what": "the same calls issued eagerly every step (no graph)"}}
        for p_ in params:            # gradients allocated inside the captures stay with the graphs
            p_.grad = None

    # ---- one full optimisation step through the plug-in surface: interior (this rank's points) +
    # boundary loss + backward + (N > 1: one packed gradient all-reduce) + Adam + the per-step
    # loss.item() sync, exactly Optimizer.closure / opt.step of common.py (N > 1: the
    # data-parallel twin, spinn_b200/dist_optimizer.py) ----
    train_step = None
    if not args.no_e2e:
        from spinn_b200.common import Optimizer
        from spinn_b200.dist_optimizer import DistributedOptimizer
        from spinn_b200.spinn2d import Plotter2D
        prev_w = [p_.detach().clone() for p_ in params]
        cls = Optimizer if ws == 1 else DistributedOptimizer       # the latter shards pde.p_samples
        solver = cls(pde, nn, Plotter2D(pde, nn), n_train=1, n_skip=10 ** 9, lr=1e-3, plot=False)
        solver.opt = solver.make_optimizer()                        # Adam (fused kernel on CUDA parameters)
        for _ in range(2):
            solver.opt.step(solver.closure)
        barrier()
        t0 = time.perf_counter()
        nts = max(3, args.steps // 2)
        for _ in range(nts):
            solver.opt.step(solver.closure)
        torch.cuda.synchronize()
        ms_ts = max_over_ranks(1e3 * (time.perf_counter() - t0) / nts)
        train_step = {"ms": ms_ts, "steps": nts, "steps_per_s": 1e3 / ms_ts,
                      "what": "opt.step(closure): interior %d pts (%d per GPU) + boundary %d pts + backward + "
                              "%sAdam + loss.item(), points resident on the GPU"
                              % (N, n_loc, pde.b_samples[0].numel(), "gradient all-reduce + " if ws > 1 else "")}
        with torch.no_grad():
            for p_, w_ in zip(params, prev_w):
                p_.copy_(w_)
        # the same loop with --graph: one captured step (incl. Adam and, for N > 1, the NCCL
        # all-reduce) replayed; losses stay on the device until the end of solve()
        import contextlib
        import io
        from spinn_b200.common import Plotter
        from spinn_b200.graphed import GraphedOptimizer
        gsolver = GraphedOptimizer(pde, nn, Plotter(pde, nn), n_train=4, n_skip=10 ** 9, tol=-1.0, lr=1e-3,
                                   plot=False)
        with contextlib.redirect_stdout(io.StringIO()):
            gsolver.solve()                                           # 3 eager steps + capture + 1 replay
            gsolver.n_train = nts
            barrier()
            t0 = time.perf_counter()
            gsolver.solve()
            torch.cuda.synchronize()
        ms_g = max_over_ranks(1e3 * (time.perf_counter() - t0) / nts)
        train_step["graphed_ms"] = ms_g
        train_step["graphed_steps_per_s"] = 1e3 / ms_g
        with torch.no_grad():
            for p_, w_ in zip(params, prev_w):
                p_.copy_(w_)
        for p_ in params:
            p_.grad = None

    if rank != 0:
        if ws > 1:
            dist.barrier()
        return

    peaks = microbench_peaks(lib, dev)
    clocks = clk.summary()
    clk_mhz = clocks.get("sm_mhz") or NOMINAL_MHZ
    kind = args.kind
    bwd_which = "bwd_general" if args.general_backward else "bwd_structured"
    k_fwd = kernel_report(kind, "fwd", pairs_loc, ms_fwd, clk_mhz)
    k_bwd = kernel_report(kind, bwd_which, pairs_loc, ms_bwd, clk_mhz)
    FLOP_PER_PAIR = k_fwd["flop_per_pair"] + k_bwd["flop_per_pair"]
    fp32_peak = SMS * LANES * 2 * clk_mhz * 1e6                 # flop/s at the clock seen under load
    ach_step = N * M * FLOP_PER_PAIR / (ms_step * 1e-3) / ws
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per step and point: x,y read twice (8+8 B), jets written (20 B), point-pair records written by the
    # forward epilogue and read by the backward (32 B each way)
    hbm_bytes_step = (16.0 + 20.0 + 64.0) * n_loc
    dom_prefix = ("g2k1_bwd_kernel" if kind == "gaussian" else "s2k1_bwd_kernel")
    traffic = ncu_traffic(dom_prefix, N, M) if (ws == 1 and not args.general_backward) else None
    probe_fp32 = 2.0 * max(peaks["ffma_3reg"], peaks["ffma_shared"], peaks["ffma2_shared"])
    roofline = {
        "bound": "fp32", "kernel": "%s (%s upstream)" % (dom_prefix, "general 5-row" if args.general_backward
                                                          else "structured u_xx,u_yy"),
        "achieved": k_bwd["tflops"], "peak": fp32_peak / 1e12, "unit": "TFLOP/s", "frac": k_bwd["frac"],
        "peak_source": "nominal FP32 rate 148 SM x 128 lanes x 2 flop x %.0f MHz (SM clock sampled during the timed "
                       "region); MEASURED_PEAKS.json holds no FP32/MUFU figure.  The in-run FMA probe "
                       "(register-only FFMA chains, 64 warps/SM) is listed under `probe`." % clk_mhz,
        "probe": {"fp32_tflops": probe_fp32 / 1e12, "frac_of_probe": k_bwd["tflops"] * 1e12 / probe_fp32,
                  "mufu_ex2_gops": peaks["ex2"] / 1e9},
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full`
        # capture of this very workload (profiles/r02_ncu_kernels.json, regenerated by tools/ncu_summary.py)
        "traffic": traffic["bytes"] if traffic else None,
        "traffic_source": traffic,
        # point-pair records (4 x 16 B) read once + per-chunk partial sums (4 rows x M floats) written;
        # chunks = 2 waves x SMs x 2 CTAs/SM / node groups of 1024
        "traffic_algorithmic": (64.0 * (n_loc / 2) + 16.0 * M * (2 * SMS * 2 // max(1, -(-M // 1024)))
                                if not args.general_backward else None),
        "algorithmic_flop_per_pair": {"fwd": k_fwd["flop_per_pair"], "bwd": k_bwd["flop_per_pair"],
                                      "step": FLOP_PER_PAIR, "survey_8d_general_gaussian": 76.0},
        "kernels": {
            "fwd": k_fwd, "bwd": k_bwd,
            "step": {"ms": ms_step, "tflops_per_gpu": ach_step / 1e12, "frac": ach_step / fp32_peak},
        },
        "mufu": {"achieved_gops": pairs_loc * (k_fwd["mufu_per_pair"] + k_bwd["mufu_per_pair"])
                 / ((ms_fwd + ms_bwd) * 1e-3) / 1e9,
                 "peak_gops": SMS * MUFU_LANES * clk_mhz * 1e-3,
                 "frac": pairs_loc * (k_fwd["mufu_per_pair"] + k_bwd["mufu_per_pair"])
                 / ((ms_fwd + ms_bwd) * 1e-3) / (SMS * MUFU_LANES * clk_mhz * 1e6)},
        "hbm": {"achieved_gbs": hbm_bytes_step / (ms_step * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": hbm_bytes_step / (ms_step * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
        "microbench_lane_ops_per_s": peaks,
        "notes": "fp32_slot_frac = FP32-pipe instruction slots issued / (148 SM x 128 lanes x clock): packed "
                 "FMUL2/FADD2 cost the pipe as much as FFMA2 (profiles/r02a_microbench_probes.txt, probes "
                 "15-18: 92-95 % of the lane rate each), so the flop-based `frac` understates pipe occupancy "
                 "by the add/mul share of the mix; a register-only probe of the forward's own mix reaches "
                 "0.776 (probe 19).",
    }

    # the softplus-hat kernel on the same point set: same fused step, MUFU-bound forward
    softplus = None
    if ws == 1 and kind == "gaussian" and not args.general_backward and not args.no_softplus:
        try:
            pde_s, nn_s = synthetic.make_state(args.nodes, args.b_nodes, args.samples, "softplus", device=dev)
            sstep = PoissonInteriorStep(nn_s, xs, ys, n_total=N)
            for _ in range(2):
                sstep.run(all_reduce=False)
            ms_s = time_call(lambda: sstep.run(all_reduce=False), reps=3)
            sstep.run(all_reduce=False, separate=True)
            ls1, ls2 = nn_s.layer1, nn_s.layer2

            def s_fwd():
                assert lib.spinn2d_jets_fwd(1, 2, n_loc, sstep.Mf, sstep.Mx, 1, P(xs), P(ys), P(ls1.x), P(ls1.y),
                                            P(ls1.xf), P(ls1.yf), P(ls1.h), 0, P(ls2.weight), P(ls2.bias),
                                            P(sstep.jets), P(sstep.ws_f), sstep.ws_f.numel(), S()) == 0

            def s_bwd():
                q = sstep.pack
                assert lib.spinn2d_jets_bwd(1, 2, 0x18, n_loc, sstep.Mf, sstep.Mx, 1, P(xs), P(ys), P(ls1.x),
                                            P(ls1.y), P(ls1.xf), P(ls1.yf), P(ls1.h), 0, P(ls2.weight),
                                            P(sstep.gjets), P(q.view("d_x")), P(q.view("d_y")), P(q.view("d_h")),
                                            P(q.view("d_W")), P(q.view("d_b")), P(sstep.ws_b), sstep.ws_b.numel(),
                                            S()) == 0

            sf = kernel_report("softplus", "fwd", pairs_loc, time_call(s_fwd, reps=3), clk_mhz)
            sb = kernel_report("softplus", "bwd_structured", pairs_loc, time_call(s_bwd, reps=3), clk_mhz)
            softplus = {"value": N * M / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s,
                        "loss": float(sstep.pack.view("loss")[0]), "kernels": {"fwd": sf, "bwd": sb},
                        "what": "the same fused interior step with the softplus-hat kernel (spinn2d.py:195-205): "
                                "s2k1_fwd_kernel is MUFU-bound (5 MUFU + 34 FP32 slots per pair), s2k1_bwd_kernel "
                                "FP32-bound (51 slots + 4 MUFU); reference CPU rate for this kernel: 4.4e6 evals/s "
                                "on 8 threads (BASELINE.md section 2)"}
            del sstep, pde_s, nn_s
        except Exception as exc:  # pragma: no cover
            softplus = {"error": repr(exc)[:300]}

    # the general 5-row backward on the same inputs (what a residual using all jets would cost)
    general = None
    if ws == 1 and not args.general_backward:
        gstep = PoissonInteriorStep(nn, xs, ys, n_total=N, structured=False)
        gstep.gjets.copy_(step.gjets)

        def bwd_general():
            p = gstep.pack
            rc = lib.spinn2d_jets_bwd(nn.kind, 2, 0, n_loc, gstep.Mf, gstep.Mx, 1, P(xs), P(ys), P(l1.x), P(l1.y),
                                      P(l1.xf), P(l1.yf), P(l1.h), hs, P(l2.weight), P(gstep.gjets),
                                      P(p.view("d_x")), P(p.view("d_y")), P(p.view("d_h")), P(p.view("d_W")),
                                      P(p.view("d_b")), P(gstep.ws_b), gstep.ws_b.numel(), S())
            assert rc == 0

        ms_g = time_call(bwd_general)
        nloss = step.pack.slices["loss"][0]
        diff = float((gstep.pack.flat[:nloss] - step.pack.flat[:nloss]).abs().max()
                     / step.pack.flat[:nloss].abs().max())
        general = {"bwd_ms": ms_g, "bwd_tflops": pairs_loc * KERNEL_COST[(kind, "bwd_general")]["flop"] / (ms_g * 1e-3) / 1e12,
                   "kernel": kernel_report(kind, "bwd_general", pairs_loc, ms_g, clk_mhz),
                   "step_ms_equivalent": ms_step - ms_bwd + ms_g,
                   "evals_per_s_equivalent": N * M / ((ms_step - ms_bwd + ms_g) * 1e-3),
                   "max_rel_diff_vs_structured": diff,
                   "note": "same gradients; the structured variant skips terms whose upstream gradient "
                           "(u, u_x, u_y) is structurally absent for this residual"}
        del gstep

    # OPT-IN compact-support mode, reported separately: it skips (node block, point group)
    # combinations beyond 6.5 widths, so its pairs/s is NOT the dense metric
    compact = None
    if ws == 1 and args.kind == "gaussian":
        try:
            cstep = PoissonInteriorStep(nn, xs, ys, n_total=N, cutoff=6.5)
            cstep.run(all_reduce=False)
            cstep.stats.zero_()
            cstep.run(all_reduce=False)
            torch.cuda.synchronize()
            fr = [float(v) / (N * M) for v in cstep.stats]
            ms_c = time_call(lambda: cstep.run(all_reduce=False))
            nloss = step.pack.slices["loss"][0]
            step.run(all_reduce=False)
            dev_g = float((cstep.pack.flat[:nloss] - step.pack.flat[:nloss]).abs().max()
                          / step.pack.flat[:nloss].abs().max())
            compact = {"cutoff_widths": 6.5, "ms_per_step": ms_c, "speedup_vs_dense_step": ms_step / ms_c,
                       "evaluated_pair_fraction": {"fwd": fr[0], "bwd": fr[1]},
                       "skipped_pair_fraction": {"fwd": 1 - fr[0], "bwd": 1 - fr[1]},
                       "max_rel_grad_diff_vs_dense": dev_g,
                       "note": "same formulas on the pairs within reach; skipped work is skipped work -- "
                               "not comparable with the dense evals/s headline"}
            del cstep
        except Exception as exc:  # pragma: no cover
            compact = {"error": repr(exc)[:200]}

    # the reference ITSELF (baseline/_ref/code) on THIS GPU, eager, at the largest point count that
    # fits in HBM (~100 B per pair -> ~400K points at 4096 nodes)
    gpu_eager = None
    if ws == 1 and not args.no_cpu_baseline:
        from baseline import reference_loader as rl
        del step
        torch.cuda.empty_cache()
        if rl.available():
            tried = []
            for chunk in (393216, 262144, 131072, 65536):
                try:
                    r = reference_run(args, dev, chunk, steps=3, warmup=1)
                    gpu_eager = {"value": r["value"], "unit": UNIT, "ms_per_chunk": r["ms_per_step"],
                                 "points_per_chunk": r["points"], "did_not_fit": tried, "kind": "reference",
                                 "what": r["what"] + " -- on the same B200 (--gpu path), 3 reps after 1 warm-up"}
                    break
                except torch.cuda.OutOfMemoryError:
                    tried.append(chunk)
                except Exception as exc:  # pragma: no cover
                    gpu_eager = {"error": repr(exc)[:200]}
                    break
                finally:
                    import gc
                    gc.collect()
                    torch.cuda.empty_cache()
        else:
            gpu_eager = {"unavailable": "baseline/_ref/code did not travel"}

    cpu_baseline = None
    if ws == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args, steps=3, warmup=1)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                        "ms_per_sample_step": r["ms_per_step"],
                        "sample": f"{r['points']} of {N} points x {r['nodes']} nodes, 3 reps after 1 warm-up, "
                                  f"{r['what']}, {r['cores']} torch threads of {r['nproc']} cpus"}
        common.device(dev)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.kind, N, M), "parallelism": f"points sharded over {ws} GPU(s)",
                   "points": N, "nodes": M, "points_per_gpu": n_loc, "kind": args.kind,
                   "forward": "all 5 jets for every point-node pair (dense, no cut-off)",
                   "backward": ("general: upstream gradients for all 5 jets" if args.general_backward else
                                "structured: this residual has upstream gradients for u_xx,u_yy only (what "
                                "autograd hands the op through the plug-in API too); every pair evaluated; "
                                "see general_upstream for the 5-row variant"),
                   "l2_policy": "inputs larger than L2: per step and point x,y 8 B (read by the forward and its "
                                "epilogue), jets 20 B written, packed point records 32 B written + 32 B read "
                                "(>400 MB at 4M points vs 126 MB L2)",
                   "collective": "one NCCL all_reduce(sum) of the packed [grads|loss] vector per step"
                                 if ws > 1 else "none (1 GPU)"},
        "loss": loss, "train_step_ms": (train_step or {}).get("ms", ms_step), "train_step": train_step,
        "clocks": clocks, "e2e": e2e, "gpu_launches": (PoissonInteriorStep.LAUNCHES_SEPARATE if args.general_backward
                                                        else PoissonInteriorStep.LAUNCHES) * args.steps,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "general_upstream": general, "softplus": softplus,
        "compact_support_opt_in": compact,
        "reference_algorithm_on_this_gpu": gpu_eager,
    }
    print(json.dumps(line), flush=True)
    if ws > 1:
        dist.barrier()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    run_ours(args)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
