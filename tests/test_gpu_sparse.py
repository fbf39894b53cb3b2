"""GPU tests of the compact-support ("cut-off") kernels: they evaluate the dense formulas on a
subset of (node block, point group) combinations, so with a generous cut-off they must agree with
the dense kernels (and the fp64 oracle) to within the Gaussian tail, for ANY node permutation and
ANY point order (coherence only affects speed)."""
import numpy as np
import pytest
import torch

from helpers import maxnorm_err, mixed_err
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu


def T(a, dtype=torch.float32):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def _problem(seed, N, Mf, Mx, coherent):
    rs = np.random.RandomState(seed)
    M = Mf + Mx
    h0 = 1.0 / np.sqrt(M)
    if coherent:
        n = int(np.ceil(np.sqrt(N)))
        gx, gy = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
        x, y = gx.ravel()[:N], gy.ravel()[:N]
    else:
        x, y = rs.rand(N), rs.rand(N)
    return dict(x=x.astype(np.float32), y=y.astype(np.float32),
                xf=rs.rand(Mf).astype(np.float32), yf=rs.rand(Mf).astype(np.float32),
                xfix=rs.rand(Mx).astype(np.float32), yfix=rs.rand(Mx).astype(np.float32),
                h=(h0 * (1 + 0.3 * rs.rand(M))).astype(np.float32),
                W=(0.3 * rs.randn(1, M)).astype(np.float32), b=(0.1 * rs.randn(1)).astype(np.float32),
                g=rs.randn(5, N, 1).astype(np.float32))


@pytest.mark.parametrize("N,Mf,Mx,coherent,perm_kind", [
    (4096, 900, 124, True, "morton"), (4099, 901, 122, True, "random"), (3000, 500, 37, False, "none"),
    (1, 40, 0, True, "morton"), (65, 1, 0, True, "none"), (2000, 0, 33, False, "morton")])
@pytest.mark.parametrize("mask", [0, 0x18, 0x01])
def test_sparse_matches_dense_and_oracle(N, Mf, Mx, coherent, perm_kind, mask):
    from spinn_b200 import ops
    be = ops.CudaBackend()
    c = _problem(N + Mf, N, Mf, Mx, coherent)
    M = Mf + Mx
    args = (T(c["x"]), T(c["y"]), T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]), T(c["h"]), T(c["W"]))
    xc = torch.cat((args[2], args[4])); yc = torch.cat((args[3], args[5]))
    perm = {"morton": lambda: ops.morton_order(xc, yc),
            "random": lambda: T(np.random.RandomState(1).permutation(M), torch.int32),
            "none": lambda: None}[perm_kind]()
    cutoff = 7.5
    stats = torch.zeros(1, dtype=torch.int64, device="cuda")
    for level in (2, 3):
        js = be.forward_sparse(level, *args, T(c["b"]), perm, cutoff, stats).cpu().numpy()
        jd = be.forward(2, 0, level, *args, T(c["b"])).cpu().numpy()
        ref = cf.jets2d("gaussian", c["x"], c["y"], xc.cpu().numpy(), yc.cpu().numpy(), c["h"], c["W"], c["b"])
        for a in range(js.shape[0]):
            # vs dense: only the summation order differs (fp32 round-off) + the tail beyond 7.5 widths
            assert maxnorm_err(js[a], jd[a]) < 1e-5, (level, a, maxnorm_err(js[a], jd[a]))
            assert mixed_err(js[a], ref[a]) < 3e-5, (level, a, mixed_err(js[a], ref[a]))
    assert 0 < int(stats[0]) <= 2 * (((N + 63) // 64) * 64) * (((M + 31) // 32) * 32)
    g = c["g"].copy()
    gdev = g.copy()
    if mask:
        for a in range(5):
            if not (mask >> a) & 1:
                g[a] = 0.0
                gdev[a] = np.nan
    gs = be.backward_sparse(2, *args, T(gdev), True, mask, perm, cutoff)
    gd = be.backward(2, 0, 2, *args, T(gdev), True, mask)
    G = cf.grads2d("gaussian", c["x"], c["y"], xc.cpu().numpy(), yc.cpu().numpy(), c["h"], c["W"], g, n_free=Mf)
    for a, b_, key in zip(gs, gd, ("d_xc", "d_yc", "d_h", "d_W", "d_b")):
        a, b_ = a.cpu().numpy(), b_.cpu().numpy()
        assert a.shape == b_.shape and np.all(np.isfinite(a))
        assert maxnorm_err(a, b_) < 1e-5, (key, maxnorm_err(a, b_))
        assert mixed_err(a, G[key]) < 3e-5, (key, mixed_err(a, G[key]))
    # deterministic (no atomics on the data path)
    gs2 = be.backward_sparse(2, *args, T(gdev), True, mask, perm, cutoff)
    assert all(torch.equal(u, v) for u, v in zip(gs, gs2) if u is not None)


def test_sparse_scalar_width_and_errors():
    from spinn_b200 import _cabi, ops
    be = ops.CudaBackend()
    c = _problem(5, 1000, 200, 24, True)
    c["h"] = np.array([0.07], np.float32)
    args = (T(c["x"]), T(c["y"]), T(c["xf"]), T(c["yf"]), T(c["xfix"]), T(c["yfix"]), T(c["h"]), T(c["W"]))
    gs = be.backward_sparse(2, *args, T(c["g"]), True, 0, None, 8.0)
    gd = be.backward(2, 0, 2, *args, T(c["g"]), True, 0)
    assert gs[2].shape == (1,) and maxnorm_err(gs[2].cpu().numpy(), gd[2].cpu().numpy()) < 5e-6
    lib = _cabi.load()
    rc = lib.spinn2d_jets_fwd_sparse(1, 2, 10, 1, 0, 1, *([None] * 7), 0, *([None] * 3), 7.0, *([None] * 3), 0, None)
    assert rc == -4 and b"Gaussian" in lib.spinn_last_error()
    big = _problem(6, 64, 6000, 0, True)          # too many nodes for the shared-memory table
    a2 = (T(big["x"]), T(big["y"]), T(big["xf"]), T(big["yf"]), T(big["xfix"]), T(big["yfix"]), T(big["h"]), T(big["W"]))
    with pytest.raises(_cabi.SpinnNativeError):
        be.forward_sparse(2, *a2, T(big["b"]), None, 7.0)


def test_sparse_full_size_step_matches_dense_and_skips_most_pairs():
    """BASELINE size: the compact-support step gives the dense step's loss and gradients (to the
    Gaussian tail) while evaluating a small fraction of the 1.7e10 pairs."""
    from spinn_b200 import synthetic
    from spinn_b200.fused_step import PoissonInteriorStep
    pde, nn = synthetic.make_state(device="cuda")
    xs, ys = (t.detach() for t in pde.p_samples)
    dense = PoissonInteriorStep(nn, xs, ys).run(all_reduce=False).clone()
    step = PoissonInteriorStep(nn, xs, ys, cutoff=7.0)
    sparse = step.run(all_reduce=False).clone()
    torch.cuda.synchronize()
    n = step.pack.slices["loss"][0]
    scale = dense[:n].abs().max()
    assert float((sparse[:n] - dense[:n]).abs().max() / scale) < 2e-6
    assert abs(float(sparse[n]) - float(dense[n])) < 1e-6 * abs(float(dense[n]))
    total = xs.numel() * nn.layer1.n
    frac_f, frac_b = float(step.stats[0]) / total, float(step.stats[1]) / total
    assert 0.01 < frac_f < 0.5 and 0.01 < frac_b < 0.5, (frac_f, frac_b)


def test_module_cutoff_through_autograd():
    from spinn_b200 import common, synthetic
    from spinn_b200.problems import derivs
    pde, nn = synthetic.make_state(nodes=400, b_nodes=20, samples=10000, device="cuda", seed=2)
    xs, ys = pde.p_samples

    def grads(cutoff):
        nn.cutoff = cutoff
        for p in nn.parameters():
            p.grad = None
        u = nn(xs, ys)
        u, ux, uy, uxx, uyy = derivs.jets_2d(u, xs, ys)
        loss = ((uxx + uyy + u * ux) ** 2).mean()
        loss.backward()
        return float(loss), [p.grad.clone() for p in nn.parameters()]

    l0, g0 = grads(None)
    l1, g1 = grads(7.5)
    assert abs(l1 - l0) < 2e-6 * abs(l0)
    for a, b in zip(g0, g1):
        assert float((a - b).abs().max() / a.abs().max()) < 5e-6
    common.device("cpu")
