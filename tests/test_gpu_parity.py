"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI,
against the oracle on the same seeded inputs, against the reference-generated golden
fixtures, and -- at BASELINE's full size -- through size-independent properties and
sampled oracle checks.

Parity metric: helpers.mixed_err = max |new-ref| / (|ref| + 1e-2*max|ref|)
against the fp64 oracle; bar 2e-5 (the reference's OWN fp32 path sits at 0.5-1.5e-5 on the
same cases, see test_oracle_golden.py::test_reference_fp32_noise_floor_documented), and
in addition new-vs-fp64 must stay within 3x of reference-fp32-vs-fp64 (+5e-6).
"""
import os

import numpy as np
import pytest
import torch

from helpers import OP_CASES_1D, OP_CASES_2D, load_case, maxnorm_err, mixed_err
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

TOL = 2e-5
# softplus-hat: exp / reciprocal / log come from MUFU (2^-22 relative each, five per pair), so a pair is
# ~4e-7 accurate where the reference's libm-grade softplus is ~1e-7; on the reference-generated
# fixtures the fast path sits at <= 1.4e-5 (reference fp32: <= 1.4e-5, profiles/r02f_parity_report.json) and at
# <= 1.98e-5 at BASELINE density (reference fp32: <= 0.9e-5); round 1's bar was 4e-5.
TOL_HAT = 2.5e-5


def tol_for(kind):
    return TOL if kind in ("gaussian", 0) else TOL_HAT


def dev():
    return torch.device("cuda", 0)


def T(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev())


def backend():
    from spinn_b200 import ops
    assert isinstance(ops._BACKEND, ops.CudaBackend), "tests must run on the native backend"
    return ops._BACKEND


def test_native_library_is_loaded_and_exports_abi():
    from spinn_b200 import _cabi
    lib = _cabi.load()
    assert lib.spinn_abi_version() == _cabi.ABI_VERSION
    for name in _cabi.SIGNATURES:
        assert hasattr(lib, name)


def _run_2d(c, level, kind=None, gj=None):
    be = backend()
    kind = cf.kind_id(c["kind"]) if kind is None else kind
    args = (T(c["x"]), T(c["y"]), T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]), T(c["h"]).reshape(-1),
            T(c["W"]), T(c["b"]))
    jets = be.forward(2, kind, level, *args)
    grads = None
    if gj is not None:
        grads = be.backward(2, kind, level, *args[:-1], T(gj), True)
    torch.cuda.synchronize()
    return jets.cpu().numpy(), None if grads is None else [g.cpu().numpy() for g in grads]


@pytest.mark.parametrize("name", OP_CASES_2D)
@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_golden_2d_all_levels(name, level):
    c = load_case(name)
    nj = (1, 3, 5, 6)[level]
    K = int(c["K"])
    rs = np.random.RandomState(level)
    gj = c["gjets"][:nj] if level >= 2 else rs.randn(nj, len(c["x"]), K).astype(np.float32)
    jets, grads = _run_2d(c, level, gj=gj)
    assert jets.shape == (nj, len(c["x"]), K)
    for a in range(nj):
        e_new = mixed_err(jets[a], c["jets_f64"][a])
        e_ref = mixed_err(c["jets_f32"][a], c["jets_f64"][a])
        assert e_new < tol_for(c["kind"]) and e_new < 3 * e_ref + 5e-6, (name, level, a, e_new, e_ref)
    xc = np.concatenate([c["xf"], c["xfix"]])
    yc = np.concatenate([c["yf"], c["yfix"]])
    g6 = np.zeros((6,) + gj.shape[1:])
    g6[:nj] = gj
    G = cf.grads2d(c["kind"], c["x"], c["y"], xc, yc, c["h"], c["W"], g6, n_free=len(c["xf"]))
    for got, key in zip(grads, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
        assert got.shape == np.asarray(G[key]).shape, (name, key, got.shape)
        assert mixed_err(got, G[key]) < tol_for(c["kind"]), (name, level, key, mixed_err(got, G[key]))
    if level >= 2 and (level == 3 or True):
        # the very gradients the reference's autograd produced (gjets[5]=0 in the fixtures)
        for got, key in zip(grads, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
            e_new = mixed_err(got, c[key + "_f64"])
            e_ref = mixed_err(c[key + "_f32"], c[key + "_f64"])
            assert e_new < tol_for(c["kind"]) and e_new < 3 * e_ref + 1e-5, (name, key, e_new, e_ref)


@pytest.mark.parametrize("name", OP_CASES_1D)
@pytest.mark.parametrize("level", [0, 1, 2])
def test_golden_1d_all_levels(name, level):
    c = load_case(name)
    be = backend()
    kind = cf.kind_id(c["kind"])
    K = int(c["K"])
    nj = level + 1
    gj = c["gjets"][:nj] if level == 2 else np.random.RandomState(level).randn(nj, len(c["x"]), K).astype(np.float32)
    args = (T(c["x"]), None, T(c["xf"]), None, T(c["xfix"]), None, T(c["h"]), T(c["W"]), T(c["b"]))
    jets = be.forward(1, kind, level, *args).cpu().numpy()
    grads = be.backward(1, kind, level, *args[:-1], T(gj), True)
    for a in range(nj):
        assert mixed_err(jets[a], c["jets_f64"][a]) < tol_for(kind), (name, level, a)
    xc = np.concatenate([c["xf"], c["xfix"]])
    g3 = np.zeros((3,) + gj.shape[1:])
    g3[:nj] = gj
    G = cf.grads1d(kind, c["x"], xc, c["h"], c["W"], g3, n_free=len(c["xf"]))
    d_xc, _, d_h, d_W, d_b = grads
    for got, key in ((d_xc, "d_xc"), (d_h, "d_h"), (d_W, "d_W"), (d_b, "d_b")):
        assert mixed_err(got.cpu().numpy(), G[key]) < tol_for(kind), (name, level, key)


def _random_problem(seed, N, Mf, Mx, K, span=1.0):
    rs = np.random.RandomState(seed)
    M = Mf + Mx
    h0 = span / np.sqrt(M)
    return dict(
        x=(span * rs.rand(N)).astype(np.float32), y=(span * rs.rand(N)).astype(np.float32),
        xf=(span * rs.rand(Mf)).astype(np.float32), yf=(span * rs.rand(Mf)).astype(np.float32),
        xfix=(span * rs.rand(Mx)).astype(np.float32), yfix=(span * rs.rand(Mx)).astype(np.float32),
        h=(h0 * (1 + 0.5 * rs.rand(M))).astype(np.float32),
        W=(0.3 * rs.randn(K, M)).astype(np.float32), b=(0.1 * rs.randn(K)).astype(np.float32),
        gjets=rs.randn(6, N, K).astype(np.float32), kind="gaussian", K=K)


@pytest.mark.parametrize("kind", ["gaussian", "softplus"])
@pytest.mark.parametrize("N,Mf,Mx,K", [
    (3001, 450, 63, 1),      # odd N, odd M: pair padding on both axes, several point tiles/chunks
    (1, 5, 0, 1),            # single point, no fixed nodes
    (2, 1, 0, 1),            # single node
    (777, 0, 40, 1),         # no free nodes at all
    (5000, 4200, 100, 1),    # M > 4096: two node tiles in the fast forward
    (1500, 300, 33, 3),      # cavity-like K=3
    (640, 50, 10, 4),        # maximum K
])
def test_seeded_medium_vs_oracle(kind, N, Mf, Mx, K):
    c = _random_problem(N + Mf, N, Mf, Mx, K)
    c["kind"] = kind
    for level in (2, 3):
        nj = (5, 6)[level - 2]
        jets, grads = _run_2d(c, level, gj=c["gjets"][:nj])
        xc = np.concatenate([c["xf"], c["xfix"]])
        yc = np.concatenate([c["yf"], c["yfix"]])
        ref = cf.jets2d(kind, c["x"], c["y"], xc, yc, c["h"], c["W"], c["b"])
        g6 = np.zeros((6, N, K))
        g6[:nj] = c["gjets"][:nj]
        G = cf.grads2d(kind, c["x"], c["y"], xc, yc, c["h"], c["W"], g6, n_free=Mf)
        # bar: 2e-5, or twice the reference's own fp32 error on these inputs where that is larger
        r32 = reference_fp32_error(2, kind, c["x"], c["y"], c["xf"], c["yf"], c["xfix"], c["yfix"], c["h"], c["W"],
                                   c["b"], c["gjets"][:nj], ref, G) if N * (Mf + Mx) <= 4_000_000 else {"jets": [0.0] * nj}
        for a in range(nj):
            e = mixed_err(jets[a], ref[a])
            assert e < max(tol_for(kind), 2.0 * r32["jets"][a]), (kind, level, a, e, r32["jets"][a])
        for got, key in zip(grads, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
            e = mixed_err(got, G[key])
            assert e < max(tol_for(kind), 2.0 * r32.get(key, 0.0)), (kind, level, key, e, r32.get(key))


def test_empty_point_set():
    c = _random_problem(3, 0, 7, 3, 1)
    jets, grads = _run_2d(c, 2, gj=np.zeros((5, 0, 1), np.float32))
    assert jets.shape == (5, 0, 1)
    assert all(np.all(g == 0) for g in grads)


def test_scalar_width_fixed_h():
    c = _random_problem(5, 400, 30, 6, 1)
    c["h"] = np.array([0.17], np.float32)
    jets, grads = _run_2d(c, 2, gj=c["gjets"][:5])
    xc = np.concatenate([c["xf"], c["xfix"]])
    yc = np.concatenate([c["yf"], c["yfix"]])
    G = cf.grads2d("gaussian", c["x"], c["y"], xc, yc, c["h"], c["W"], c["gjets"][:5], n_free=30)
    assert grads[2].shape == (1,)
    assert mixed_err(grads[2], G["d_h"]) < TOL


def test_unaligned_pointers_take_the_scalar_path():
    from spinn_b200 import ops
    c = _random_problem(9, 4099, 60, 4, 1)
    be = backend()
    big_x, big_y = T(np.concatenate([[0.0], c["x"]])), T(np.concatenate([[0.0], c["y"]]))
    x, y = big_x[1:], big_y[1:]                       # 4-byte offset: not 16-byte aligned
    assert x.data_ptr() % 16 != 0
    args = (x, y, T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]), T(c["h"]), T(c["W"]), T(c["b"]))
    jets = be.forward(2, ops.GAUSSIAN, 2, *args).cpu().numpy()
    xc = np.concatenate([c["xf"], c["xfix"]])
    yc = np.concatenate([c["yf"], c["yfix"]])
    ref = cf.jets2d("gaussian", c["x"], c["y"], xc, yc, c["h"], c["W"], c["b"])
    for a in range(5):
        assert mixed_err(jets[a], ref[a]) < TOL


def test_bad_arguments_are_reported_not_crashed():
    from spinn_b200 import _cabi
    lib = _cabi.load()
    rc = lib.spinn2d_jets_fwd(7, 2, 10, 1, 0, 1, *([None] * 7), 0, *([None] * 4), 0, None)
    assert rc < 0 and b"kind" in lib.spinn_last_error()
    rc = lib.spinn2d_jets_fwd(0, 2, 10, 1, 0, 9, *([None] * 7), 0, *([None] * 4), 0, None)
    assert rc == -4


# ----- BASELINE full size: 4M points x 4096 nodes --------------------------------------
@pytest.fixture(scope="module")
def full_state():
    from spinn_b200 import synthetic
    pde, nn = synthetic.make_state(device="cuda")
    return pde, nn


def test_full_size_forward_sampled_vs_oracle_and_linearity(full_state):
    pde, nn = full_state
    xs, ys = (t.detach() for t in pde.p_samples)
    assert xs.numel() == 4194304 and nn.layer2.weight.shape == (1, 4096)
    with torch.no_grad():
        jets = torch.stack(nn.jets(xs, ys, level=2)).squeeze(-1)          # (5, N)
        idx = torch.randint(0, xs.numel(), (96,), generator=torch.Generator().manual_seed(1))
        l1, l2 = nn.layer1, nn.layer2
        xc, yc = (t.cpu().numpy() for t in l1.centers())
        ref = cf.jets2d("gaussian", xs[idx.cuda()].cpu().numpy(), ys[idx.cuda()].cpu().numpy(), xc, yc,
                        l1.h.detach().cpu().numpy(), l2.weight.detach().cpu().numpy(),
                        l2.bias.detach().cpu().numpy())[:5, :, 0]
        got = jets[:, idx.cuda()].cpu().numpy()
        for a in range(5):
            assert mixed_err(got[a], ref[a]) < TOL, (a, mixed_err(got[a], ref[a]))
        # north_star's literal metric -- strict per-element relative error -- as a percentile table over
        # 2048 sampled points of the full-size run (test artefact; near zero crossings of a jet any fp32
        # evaluation, the reference's included, is off by 1e-4..1e-2 in this metric: SURVEY.md 7.3)
        idx2 = torch.randint(0, xs.numel(), (2048,), generator=torch.Generator().manual_seed(2)).cuda()
        ref2 = cf.jets2d("gaussian", xs[idx2].cpu().numpy(), ys[idx2].cpu().numpy(), xc, yc,
                         l1.h.detach().cpu().numpy(), l2.weight.detach().cpu().numpy(),
                         l2.bias.detach().cpu().numpy())[:5, :, 0]
        got2 = jets[:, idx2].cpu().numpy().astype(np.float64)
        table = {}
        for a, name in enumerate(("u", "u_x", "u_y", "u_xx", "u_yy")):
            rel = np.abs(got2[a] - ref2[a]) / np.maximum(np.abs(ref2[a]), 1e-300)
            table[name] = {"p50": float(np.percentile(rel, 50)), "p90": float(np.percentile(rel, 90)),
                           "p99": float(np.percentile(rel, 99)), "max": float(rel.max()),
                           "frac_within_1e-5": float((rel < 1e-5).mean()), "mixed": mixed_err(got2[a], ref2[a])}
            assert table[name]["p90"] < 1e-5, (name, table[name])
        _dump_report("parity_full_size_strict_percentiles.json", table)
        # linearity in the weights: J(2W, 2b) = 2 J(W, b)  (bit-exact scaling by a power of two)
        l2.weight.mul_(2.0)
        l2.bias.mul_(2.0)
        jets2 = torch.stack(nn.jets(xs, ys, level=2)).squeeze(-1)
        l2.weight.mul_(0.5)
        l2.bias.mul_(0.5)
        assert torch.equal(jets2, 2.0 * jets)


def test_full_size_backward_sampled_nodes_vs_oracle(full_state):
    from spinn_b200.fused_step import PoissonInteriorStep
    pde, nn = full_state
    xs, ys = (t.detach() for t in pde.p_samples)
    step = PoissonInteriorStep(nn, xs, ys)
    flat = step.run(all_reduce=False).clone()
    torch.cuda.synchronize()
    l1, l2 = nn.layer1, nn.layer2
    # oracle on a subset of NODES over a subset of POINTS is not a valid check of a sum over all
    # points, so: all 4M points, 6 nodes, fp64, chunked (gradients of node j depend on node j only
    # once the upstream g = dLoss/dJets is given).
    g = step.upstream_from_jets().cpu().numpy().astype(np.float64)[:, :, None]            # (5, N, 1)
    x, y = xs.cpu().numpy(), ys.cpu().numpy()
    xc, yc = (t.detach().cpu().numpy() for t in l1.centers())
    h, W = l1.h.detach().cpu().numpy(), l2.weight.detach().cpu().numpy()
    nodes = np.array([0, 1777, 3599, 3600, 3900, 4095])
    G = cf.grads2d("gaussian", x, y, xc[nodes], yc[nodes], h[nodes], W[:, nodes], g, chunk=65536)
    p = step.pack
    d_W = p.view("d_W").cpu().numpy()
    d_h = p.view("d_h").cpu().numpy()
    d_x = p.view("d_x").cpu().numpy()
    scale_W, scale_h, scale_x = np.abs(d_W).max(), np.abs(d_h).max(), np.abs(d_x).max()
    assert np.max(np.abs(d_W[nodes] - G["d_W"][0])) < 1e-5 * scale_W + 1e-5 * np.abs(G["d_W"][0]).max()
    assert np.max(np.abs(d_h[nodes] - G["d_h"])) < 1e-5 * scale_h + 1e-5 * np.abs(G["d_h"]).max()
    free = nodes[nodes < 3600]
    assert np.max(np.abs(d_x[free] - G["d_xc"][:len(free)])) < 1e-5 * scale_x + 1e-5 * np.abs(G["d_xc"]).max()
    # bias gradient = sum of g0 (=0 for this residual); loss equals mean(res^2) recomputed in fp64
    jets = step.jets.cpu().numpy().astype(np.float64)
    f = 20 * np.pi ** 2 * np.sin(2 * np.pi * x.astype(np.float64)) * np.sin(4 * np.pi * y.astype(np.float64))
    loss_ref = np.mean((0.1 * (jets[3] + jets[4] + f)) ** 2)
    assert abs(float(p.view("loss")[0]) - loss_ref) < 2e-5 * loss_ref
    # determinism: no atomics anywhere -> bit-identical repeat
    flat2 = step.run(all_reduce=False)
    assert torch.equal(flat, flat2)


@pytest.mark.parametrize("kind", ["gaussian", "softplus"])
@pytest.mark.parametrize("samples,nodes,b_nodes", [(40000, 400, 20), (1000 * 1000 + 1, 900, 31), (169, 25, 5)])
def test_fused_interior_step_equals_the_separate_calls(kind, samples, nodes, b_nodes):
    """spinn2d_poisson_interior_step (5 launches, no dLoss/dJets array) against forward + residual
    + backward issued one by one: same jets, bit-identical gradients, same loss; odd point counts."""
    from spinn_b200 import synthetic
    from spinn_b200.fused_step import PoissonInteriorStep
    pde, nn = synthetic.make_state(nodes=nodes, b_nodes=b_nodes, samples=samples, kind=kind, device="cuda", seed=11)
    xs, ys = (t.detach() for t in pde.p_samples)
    if samples % 2:                                    # odd N: drop to an odd count
        xs, ys = xs[:samples].contiguous(), ys[:samples].contiguous()
    step = PoissonInteriorStep(nn, xs, ys)
    nl = step.pack.slices["loss"][0]
    a = step.run(all_reduce=False).clone()
    jets_a = step.jets.clone()
    b = step.run(all_reduce=False, separate=True).clone()
    assert torch.equal(jets_a, step.jets)
    assert torch.equal(a[:nl], b[:nl])
    assert abs(float(a[nl]) - float(b[nl])) <= 2e-6 * abs(float(b[nl]))
    assert torch.allclose(step.gjets[3], step.upstream_from_jets()[3], rtol=1e-5, atol=0)


def test_full_size_point_sharding_is_additive(full_state):
    """Point shards are independent and gradients are sums over points: the 8-way sharded
    result (what 8 ranks + one all-reduce produce) equals the unsharded one."""
    from spinn_b200.distributed import shard_range
    from spinn_b200.fused_step import PoissonInteriorStep
    pde, nn = full_state
    xs, ys = (t.detach() for t in pde.p_samples)
    N = xs.numel()
    whole = PoissonInteriorStep(nn, xs, ys).run(all_reduce=False).double()
    acc = torch.zeros_like(whole)
    for r in range(8):
        lo, hi = shard_range(N, r, 8)
        acc += PoissonInteriorStep(nn, xs[lo:hi], ys[lo:hi], n_total=N).run(all_reduce=False).double()
    scale = whole.abs().max()
    assert float((acc - whole).abs().max() / scale) < 5e-6


def test_beyond_baseline_size_16m_points_16384_nodes():
    """4x the points and 4x the nodes of the BASELINE workload (2.7e11 pairs; the node table no
    longer fits one shared-memory tile, point-record offsets pass 2^31 bytes): sampled forward
    points and sampled backward nodes against the fp64 oracle, bit-exact repeatability."""
    from spinn_b200 import synthetic
    from spinn_b200.fused_step import PoissonInteriorStep
    pde, nn = synthetic.make_state(nodes=124 * 124, b_nodes=252, samples=4096 * 4096, device="cuda")
    xs, ys = (t.detach() for t in pde.p_samples)
    N = xs.numel()
    assert N == 16777216 and nn.layer2.weight.shape == (1, 16384)
    l1, l2 = nn.layer1, nn.layer2
    xc, yc = (t.detach().cpu().numpy() for t in l1.centers())
    h, W, b = (t.detach().cpu().numpy() for t in (l1.h, l2.weight, l2.bias))
    with torch.no_grad():
        jets = torch.stack(nn.jets(xs, ys, level=2)).squeeze(-1)
    idx = torch.cat((torch.tensor([0, 1, N - 2, N - 1]),
                     torch.randint(0, N, (60,), generator=torch.Generator().manual_seed(3)))).cuda()
    ref = cf.jets2d("gaussian", xs[idx].cpu().numpy(), ys[idx].cpu().numpy(), xc, yc, h, W, b)[:5, :, 0]
    got = jets[:, idx].cpu().numpy()
    for a in range(5):
        assert mixed_err(got[a], ref[a]) < TOL, (a, mixed_err(got[a], ref[a]))
    del jets
    step = PoissonInteriorStep(nn, xs, ys)
    flat = step.run(all_reduce=False).clone()
    g = step.upstream_from_jets().cpu().numpy().astype(np.float64)[:, :, None]
    nodes = np.array([0, 15375, 15376, 16383])
    G = cf.grads2d("gaussian", xs.cpu().numpy(), ys.cpu().numpy(), xc[nodes], yc[nodes], h[nodes], W[:, nodes], g,
                   chunk=262144)
    p = step.pack
    for name, key, sel in (("d_W", "d_W", nodes), ("d_h", "d_h", nodes), ("d_x", "d_xc", nodes[:2])):
        got = p.view(name).cpu().numpy()
        want = np.asarray(G[key]).reshape(-1)[:len(sel)]
        assert np.max(np.abs(got[sel] - want)) < 1e-5 * np.abs(got).max() + 1e-5 * np.abs(want).max(), name
    assert torch.equal(flat, step.run(all_reduce=False))


def reference_fp32_error(dim, kind, x, y, xf, yf, xfix, yfix, h, W, b, g_masked, oracle_jets, oracle_grads):
    """How far the REFERENCE's fp32 arithmetic (dense (N,M) tensors + nested autograd.grad: oracle/
    torch_port.py, pinned against reference-generated fixtures to <= 5e-6) is from the fp64 oracle on
    the same inputs, per quantity, in the mixed metric.  The earned slack of a parity bar."""
    from oracle import torch_port as tp
    nj = g_masked.shape[0]
    full = 6 if dim == 2 else 3
    g = np.zeros((full,) + g_masked.shape[1:], np.float32)
    g[:nj] = g_masked
    if dim == 2:
        out = tp.jets_and_grads2d(kind, x, y, xf, yf, xfix, yfix, h, W, b, g)
    else:
        out = tp.jets_and_grads1d(kind, x, xf, xfix, h, W, b, g)
    err = {"jets": [mixed_err(out["jets"][a], oracle_jets[a]) for a in range(nj)]}
    for key in ("d_xc", "d_yc", "d_h", "d_W", "d_b"):
        if key in out and out[key] is not None and np.size(out[key]):
            err[key] = mixed_err(np.asarray(out[key]).reshape(np.asarray(oracle_grads[key]).shape), oracle_grads[key])
    return err


def bar(kind, ref_err):
    """2e-5, or twice the reference's own fp32 error on the same inputs where that is larger."""
    return max(TOL, 2.0 * ref_err)


# how far from the reference's fp64 values the CUDA path may be, in units of the reference's OWN fp32
# error on the same inputs: the Gaussian kernels are as accurate as the reference's fp32 arithmetic;
# the softplus-hat fast path (five MUFU approximations per pair, 2^-22 each) is up to ~4.5x less
# accurate than torch's libm-grade softplus on sums at BASELINE density, still <= 2.5e-5
REL_TO_REF = {"gaussian": 1.5, "softplus": 5.0}


def _dump_report(name, obj):
    """Test artefact (numbers behind the assertions) into gpurun_out/, which gpurun brings back."""
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        import json
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, name), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


def load_dense(kind):
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"dense_{kind}.npz"))
    return {k: (d[k].item() if d[k].ndim == 0 else d[k]) for k in d.files}


@pytest.mark.parametrize("copies", [1, 96])
@pytest.mark.parametrize("kind", ["gaussian", "softplus"])
def test_baseline_density_vs_reference_generated_values(kind, copies):
    """BASELINE density (h = 1/64, the 4096-node synthetic state): jets and parameter gradients at
    2048 points of the collocation grid against values produced BY THE REFERENCE in fp32 and fp64
    (tests/golden/dense_*.npz).  copies = 96 tiles the point set to 196 608 points so that the
    two-points-per-thread main span of the persistent forward (and its one-point tail) is what
    runs; upstream gradients are divided by the number of copies (gradients are sums over points)."""
    c = load_dense(kind)
    be = backend()
    rep = lambda a: np.tile(a, copies)
    x, y = rep(c["x"]), rep(c["y"])
    n0 = c["x"].size
    args = (T(x), T(y), T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]), T(c["h"]), T(c["W"]))
    report, checks = {}, []
    for level in (2, 3):
        jets = be.forward(2, cf.kind_id(kind), level, *args, T(c["b"])).cpu().numpy()[:, :, 0]
        for a in range(5):
            got = jets[a].reshape(copies, n0)
            assert np.array_equal(got, np.broadcast_to(got[0], got.shape)), (level, a)       # every copy identical
            e64 = mixed_err(got[0], c["jets_f64"][a])
            e32 = mixed_err(got[0], c["jets_f32"][a])
            eref = mixed_err(c["jets_f32"][a], c["jets_f64"][a])
            report[f"jet{a}_L{level}"] = (e64, e32, eref)
    for suffix, g, mask in (("", c["gjets"], 0), ("_lapl", c["gjets_lapl"], 0x18)):
        gd = np.tile(g / copies, (1, copies)).astype(np.float32)
        if mask:
            gd[:3] = np.nan                                   # absent rows are never read
        grads = be.backward(2, cf.kind_id(kind), 2, *args, T(gd[:, :, None]), True, mask)
        for got, key in zip(grads, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
            got = got.cpu().numpy().reshape(-1)
            r64, r32 = c[key + suffix + "_f64"].reshape(-1), c[key + suffix + "_f32"].reshape(-1)
            report[key + suffix] = (mixed_err(got, r64), mixed_err(got, r32), mixed_err(r32, r64))
    _dump_report(f"parity_dense_{kind}_x{copies}.json",
                 {k: {"ours_vs_ref_fp64": v[0], "ours_vs_ref_fp32": v[1], "ref_fp32_vs_ref_fp64": v[2]}
                  for k, v in report.items()})
    for key, (e64, e32, eref) in report.items():
        assert e64 < (TOL if kind == "gaussian" else 2.5e-5), (kind, key, e64)      # measured: <= 9.7e-6 / <= 1.98e-5
        assert e64 < REL_TO_REF[kind] * eref + 3e-6 and e32 < 3e-5, (kind, key, e64, e32, eref)


@pytest.mark.parametrize("mask", [0x18, 0x08, 0x10, 0x01, 0x19, 0x06, 0x0C, 0x0F, 0x1F])
@pytest.mark.parametrize("kind,K", [("gaussian", 1), ("gaussian", 3), ("softplus", 1)])
def test_present_mask_rows_are_zero_and_never_read(mask, kind, K):
    """present_mask: absent rows of gjets count as zero and are never read (they hold NaN here);
    the structured fast paths (0x18 Laplacian-type, 0x01 u-only) agree with the full formula."""
    c = _random_problem(77 + mask, 2049, 301, 40, K)
    c["kind"] = kind
    be = backend()
    g = c["gjets"][:5].copy()
    g_dev = g.copy()
    for a in range(5):
        if not (mask >> a) & 1:
            g[a] = 0.0
            g_dev[a] = np.nan
    args = (T(c["x"]), T(c["y"]), T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]), T(c["h"]), T(c["W"]))
    grads = be.backward(2, cf.kind_id(kind), 2, *args, T(g_dev), True, mask)
    xc = np.concatenate([c["xf"], c["xfix"]])
    yc = np.concatenate([c["yf"], c["yfix"]])
    G = cf.grads2d(kind, c["x"], c["y"], xc, yc, c["h"], c["W"], g, n_free=len(c["xf"]))
    # seeded random clouds (301+40 random centres, heavy cancellation in d_h): the slack over 2e-5 is
    # whatever the reference's own fp32 arithmetic needs on these very inputs
    J = cf.jets2d(kind, c["x"], c["y"], xc, yc, c["h"], c["W"], np.zeros(K))
    ref = reference_fp32_error(2, kind, c["x"], c["y"], c["xf"], c["yf"], c["xfix"], c["yfix"], c["h"], c["W"],
                               np.zeros(K, np.float32), g, J, G)
    for got, key in zip(grads, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
        got = got.cpu().numpy()
        assert np.all(np.isfinite(got)), (mask, key)
        e = mixed_err(got, G[key])
        assert e < bar(kind, ref.get(key, 0.0)), (mask, kind, K, key, e, ref.get(key))


def _sweep_cases():
    rs = np.random.RandomState(2024)
    cases = []
    for i in range(28):
        dim = 2 if i % 4 else 1
        kind = "gaussian" if i % 2 else "softplus"
        K = int(rs.choice([1, 1, 2, 3, 4]))
        N = int(rs.choice([1, 2, 3, 31, 64, 65, 257, 1000, 2049]))
        Mf = int(rs.choice([0, 1, 2, 31, 63, 129, 300]))
        Mx = int(rs.choice([0, 1, 5, 64])) if Mf else int(rs.choice([1, 7, 130]))
        level = int(rs.randint(0, 4 if dim == 2 else 3))
        cases.append((i, dim, kind, K, N, Mf, Mx, level))
    return cases


@pytest.mark.parametrize("i,dim,kind,K,N,Mf,Mx,level", _sweep_cases())
def test_random_shape_sweep(i, dim, kind, K, N, Mf, Mx, level):
    """Seeded sweep over awkward sizes (tiny / odd / tile-boundary N and M, every K, level, kernel,
    1-D and 2-D, random present-masks) against the fp64 oracle."""
    rs = np.random.RandomState(100 + i)
    M = Mf + Mx
    h0 = (1.0 / np.sqrt(M) if dim == 2 else 1.0 / M) * 1.5
    x, y = rs.rand(N).astype(np.float32), rs.rand(N).astype(np.float32)
    xf, yf = rs.rand(Mf).astype(np.float32), rs.rand(Mf).astype(np.float32)
    xx, yx = rs.rand(Mx).astype(np.float32), rs.rand(Mx).astype(np.float32)
    h = (h0 * (1 + 0.5 * rs.rand(M))).astype(np.float32)
    W, b = (0.3 * rs.randn(K, M)).astype(np.float32), (0.1 * rs.randn(K)).astype(np.float32)
    from spinn_b200 import ops
    nj = ops.num_jets(dim, level)
    g = rs.randn(nj, N, K).astype(np.float32)
    mask = int(rs.randint(1, 1 << nj))
    gm = g.copy()
    for a in range(nj):
        if not (mask >> a) & 1:
            gm[a] = 0.0
    be = backend()
    kid = cf.kind_id(kind)
    if dim == 2:
        args = (T(x), T(y), T(xf), T(yf), T(xx), T(yx), T(h), T(W))
        J = cf.jets2d(kind, x, y, np.concatenate([xf, xx]), np.concatenate([yf, yx]), h, W, b)
        g6 = np.zeros((6, N, K)); g6[:nj] = gm
        G = cf.grads2d(kind, x, y, np.concatenate([xf, xx]), np.concatenate([yf, yx]), h, W, g6, n_free=Mf)
        keys = ("d_xc", "d_yc", "d_h", "d_W", "d_b")
    else:
        args = (T(x), None, T(xf), None, T(xx), None, T(h), T(W))
        J = cf.jets1d(kind, x, np.concatenate([xf, xx]), h, W, b)
        g3 = np.zeros((3, N, K)); g3[:nj] = gm
        G = cf.grads1d(kind, x, np.concatenate([xf, xx]), h, W, g3, n_free=Mf)
        keys = ("d_xc", None, "d_h", "d_W", "d_b")
    jets = be.forward(dim, kid, level, *args, T(b)).cpu().numpy()
    # a handful of wide, randomly placed nodes -> sums with heavy cancellation (|sum| << sum|terms|):
    # fp32 round-off of any evaluation order shows up as a few 1e-5 here; the bar is 2e-5 or twice
    # what the reference's own fp32 arithmetic is off on the same inputs
    ref = reference_fp32_error(dim, kind, x, y, xf, yf, xx, yx, h, W, b, gm, J, G)
    for a in range(nj):
        e = mixed_err(jets[a], J[a])
        assert e < bar(kind, ref["jets"][a]), (a, e, ref["jets"][a])
    grads = be.backward(dim, kid, level, *args, T(g), True, mask)
    for got, key in zip(grads, keys):
        if key is None:
            continue
        got = got.cpu().numpy()
        assert got.shape == np.asarray(G[key]).shape
        e = mixed_err(got, G[key])
        assert e < bar(kind, ref.get(key, 0.0)), (key, e, ref.get(key))
