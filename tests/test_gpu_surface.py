"""GPU tests of the plug-in surface: the autograd tower over the real CUDA op, and the
four BASELINE configurations driven through App.run(--gpu), compared step by step with
loss histories recorded from the reference itself (tests/golden/traj_*.npz)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_case, mixed_err

pytestmark = pytest.mark.gpu


class _StubPDE:
    def __init__(self, nodes, fixed, k):
        self._n, self._f, self._k = nodes, fixed, k

    def nodes(self):
        return self._n

    def fixed_nodes(self):
        return self._f

    def n_vars(self):
        return self._k


@pytest.fixture()
def on_gpu():
    from spinn_b200 import common, ops
    assert isinstance(ops._BACKEND, ops.CudaBackend)
    common.device("cuda")
    yield
    common.device("cpu")


@pytest.mark.parametrize("name", ["g2d_k1", "s2d_k1", "g2d_k3", "s2d_k3", "g2d_nofixed", "g2d_fixed_h",
                                  "g2d_one_point", "g2d_wide", "s2d_wide"])
def test_tower_on_gpu_matches_reference_autograd(on_gpu, name):
    from spinn_b200 import spinn2d
    from spinn_b200.common import tensor
    from spinn_b200.problems import derivs
    c = load_case(name)
    K = int(c["K"])
    act = spinn2d.gaussian if c["kind"] == "gaussian" else spinn2d.SoftPlus()
    nn = spinn2d.SPINN2D(_StubPDE((c["xf"], c["yf"]), (c["xfix"], c["yfix"]), K), act,
                         fixed_h=c["h"].size == 1).to("cuda")
    with torch.no_grad():
        nn.layer1.x.copy_(tensor(c["xf"]))
        nn.layer1.y.copy_(tensor(c["yf"]))
        nn.layer1.h.copy_(tensor(c["h"]).reshape(nn.layer1.h.shape))
        nn.layer2.weight.copy_(tensor(c["W"]))
        nn.layer2.bias.copy_(tensor(c["b"]))
    x = tensor(c["x"], requires_grad=True)
    y = tensor(c["y"], requires_grad=True)
    U = nn(x, y).reshape(x.numel(), K)
    g = tensor(c["gjets"])
    L = 0.0
    for k in range(K):
        got = torch.stack(derivs.jets_2d(U[:, k], x, y))
        for a in range(5):
            assert mixed_err(got[a].detach().cpu().numpy(), c["jets_f64"][a, :, k]) < 2e-5
        L = L + (g[:5, :, k] * got).sum()
    L.backward()
    pairs = [(nn.layer1.x.grad, "d_xc"), (nn.layer1.y.grad, "d_yc"), (nn.layer1.h.grad.reshape(-1), "d_h"),
             (nn.layer2.weight.grad, "d_W"), (nn.layer2.bias.grad, "d_b")]
    for got, key in pairs:
        e_new = mixed_err(got.cpu().numpy(), c[key + "_f64"])
        e_ref = mixed_err(c[key + "_f32"], c[key + "_f64"])
        assert e_new < 2e-5 and e_new < 3 * e_ref + 5e-6, (name, key, e_new, e_ref)


def test_plot_grid_forward_is_u_only_and_matches(on_gpu):
    from spinn_b200 import spinn2d
    from spinn_b200.common import tensor
    c = load_case("g2d_k1")
    nn = spinn2d.SPINN2D(_StubPDE((c["xf"], c["yf"]), (c["xfix"], c["yfix"]), 1), spinn2d.gaussian).to("cuda")
    with torch.no_grad():
        nn.layer2.weight.copy_(tensor(c["W"]))
        nn.layer1.h.copy_(tensor(c["h"]))
    u = nn(tensor(c["x"]), tensor(c["y"]))             # no requires_grad -> level 0
    assert u.shape == (len(c["x"]),) and not u.requires_grad or u.grad_fn is not None


def _traj(name):
    d = np.load(os.path.join(GOLDEN, f"traj_{name}.npz"))
    return d["loss"], [str(a) for a in d["args"]], d["errors"]


@pytest.mark.parametrize("name,modname", [
    ("ode3", "ode3"), ("ode3_softplus", "ode3"),
    ("poisson2d_sine", "poisson2d_sine"), ("poisson2d_sine_softplus", "poisson2d_sine"),
    ("irreg", "poisson2d_irreg_dom"), ("cavity", "cavity"),
    ("heat1d", "heat1d"), ("burgers1d", "burgers1d"),
    ("poisson2d_sine_pu", "poisson2d_sine"), ("poisson2d_sine_fixed_h", "poisson2d_sine")])
def test_config_trajectory_on_gpu_matches_reference(name, modname):
    import importlib
    mod = importlib.import_module(f"spinn_b200.problems.{modname}")
    ref, args, ref_err = _traj(name)
    np.random.seed(0)
    torch.manual_seed(0)
    app = mod.make_app()
    app.run(args=args + ["--gpu"])
    loss = np.asarray(app.solver.loss)
    n = min(len(loss), len(ref))
    assert n == len(ref)
    rel = np.abs(loss[:n] - ref[:n]) / np.abs(ref[:n])
    assert rel[:10].max() < 5e-5, (name, rel[:10].max())
    assert rel.max() < 3e-3, (name, rel.max())
    err = app.plotter.get_error()
    if ref_err[2] > 0:
        assert abs(err[2] - ref_err[2]) < 0.01 * ref_err[2] + 1e-6, (name, err, ref_err)


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("name,modname", [("heat1d_fd", "heat1d_fd_spinn"), ("burgers1d_fd", "burgers1d_fd_spinn"),
                                          ("advection1d_fd", "advection1d_fd_spinn")])
def test_fd_spinn_time_stepping_on_gpu_matches_reference(name, modname, graph):
    """FD-SPINN outer loop (reference fd_spinn1d_base.py:142-190) on the 1-D fused op: eager, and
    with ONE captured training step replayed across all time levels (--graph; the advection case
    sub-samples half of the points per step through the static index buffer)."""
    import importlib
    mod = importlib.import_module(f"spinn_b200.problems.{modname}")
    d = np.load(os.path.join(GOLDEN, f"traj_{name}.npz"))
    args = [str(a) for a in d["args"]]
    np.random.seed(0)
    torch.manual_seed(0)
    app = mod.make_app()
    app.run(args=args + ["--gpu"] + (["--graph"] if graph else []))
    loss, ref = np.asarray(app.solver.loss), d["loss"]
    assert len(loss) == len(ref), (len(loss), len(ref))
    rel = np.abs(loss - ref) / (np.abs(ref) + 1e-3 * np.abs(ref).max())
    assert rel[:10].max() < 5e-5 and rel.max() < 5e-3, (name, rel[:10].max(), rel.max())
    assert app.pde.iter == int(d["iter"])
    u0 = app.pde.u0.detach().cpu().numpy()
    # burgers: sum(u0) = 0 up to rounding, so Adam moves the bias by +-lr per step along the SIGN of
    # rounding noise for the first steps (in the reference too) -> absolute slack of a few lr
    lr = float(args[args.index("--lr") + 1])
    assert np.abs(u0 - d["u0"]).max() < max(3e-3 * np.abs(d["u0"]).max(), 10 * lr)
    if graph:
        assert getattr(app.solver, "_captured", (None,))[0] is not None


@pytest.mark.parametrize("name,modname,steps", [
    ("ode3", "ode3", 120), ("poisson2d_sine", "poisson2d_sine", 80),
    ("irreg", "poisson2d_irreg_dom", 40), ("cavity", "cavity", 60)])
def test_cuda_graphed_step_follows_the_eager_trajectory(name, modname, steps):
    """--graph (one training step captured in a CUDA graph and replayed) must reproduce the
    reference's recorded loss history like the eager loop does, including the cavity's
    per-step random sub-sampling (same np.random stream)."""
    import importlib
    mod = importlib.import_module(f"spinn_b200.problems.{modname}")
    ref, args, _ = _traj(name)
    i = args.index("--n-train")
    args[i + 1] = str(steps)
    np.random.seed(0)
    torch.manual_seed(0)
    app = mod.make_app()
    app.run(args=args + ["--gpu", "--graph"])
    from spinn_b200.graphed import GraphedOptimizer
    assert isinstance(app.solver, GraphedOptimizer)
    loss = np.asarray(app.solver.loss)
    assert len(loss) == steps
    rel = np.abs(loss - ref[:steps]) / np.abs(ref[:steps])
    assert rel[:10].max() < 5e-5, (name, rel[:10].max())
    assert rel.max() < 3e-3, (name, rel.max())


@pytest.mark.parametrize("name", ["g2d_k1", "s2d_k1", "g2d_k3"])
def test_partition_of_unity_on_gpu_matches_reference(on_gpu, name):
    """--pu: u = sum W phi / sum phi through one fused call with an extra all-ones output and the
    quotient rule on the jets; jets and gradients vs the reference's autograd (fp64 fixtures)."""
    from spinn_b200 import spinn2d
    from spinn_b200.common import tensor
    from spinn_b200.problems import derivs
    d = np.load(os.path.join(GOLDEN, f"pu_{name}.npz"))
    c = {k: (d[k].item() if d[k].ndim == 0 else d[k]) for k in d.files}
    K = int(c["K"])
    act = spinn2d.gaussian if c["kind"] == "gaussian" else spinn2d.SoftPlus()
    nn = spinn2d.SPINN2D(_StubPDE((c["xf"], c["yf"]), (c["xfix"], c["yfix"]), K), act, use_pu=True).to("cuda")
    with torch.no_grad():
        nn.layer1.x.copy_(tensor(c["xf"]))
        nn.layer1.y.copy_(tensor(c["yf"]))
        nn.layer1.h.copy_(tensor(c["h"]))
        nn.layer2.weight.copy_(tensor(c["W"]))
    x = tensor(c["x"], requires_grad=True)
    y = tensor(c["y"], requires_grad=True)
    U = nn(x, y).reshape(x.numel(), K)
    g = tensor(c["gjets"])
    L = 0.0
    for k in range(K):
        got = torch.stack(derivs.jets_2d(U[:, k], x, y))
        for a in range(5):
            e_new = mixed_err(got[a].detach().cpu().numpy(), c["jets_f64"][a, :, k])
            e_ref = mixed_err(c["jets_f32"][a, :, k], c["jets_f64"][a, :, k])
            assert e_new < 4e-5 and e_new < 3 * e_ref + 1e-5, (name, a, e_new, e_ref)
        L = L + (g[:5, :, k] * got).sum()
    L.backward()
    for got, key in [(nn.layer1.x.grad, "d_xc"), (nn.layer1.y.grad, "d_yc"), (nn.layer1.h.grad, "d_h"),
                     (nn.layer2.weight.grad, "d_W")]:
        e_new = mixed_err(got.cpu().numpy(), c[key + "_f64"])
        e_ref = mixed_err(c[key + "_f32"], c[key + "_f64"])
        assert e_new < 5e-5 and e_new < 3 * e_ref + 1e-5, (name, key, e_new, e_ref)


def test_cutoff_and_graph_together_follow_the_reference_trajectory():
    """--cutoff 7 (compact-support kernels) + --graph on poisson2d_sine: with Gaussians 7 widths
    wide of a 1/12-wide node the cut-off drops nothing measurable, so the reference's loss history
    must still be reproduced."""
    from spinn_b200.problems import poisson2d_sine
    ref, args, _ = _traj("poisson2d_sine")
    i = args.index("--n-train")
    args[i + 1] = "120"
    np.random.seed(0)
    torch.manual_seed(0)
    app = poisson2d_sine.make_app()
    app.run(args=args + ["--gpu", "--graph", "--cutoff", "7"])
    assert app.nn.cutoff == 7.0
    loss = np.asarray(app.solver.loss)
    rel = np.abs(loss - ref[:120]) / np.abs(ref[:120])
    assert rel[:10].max() < 5e-5 and rel.max() < 3e-3, rel.max()


@pytest.mark.parametrize("modname,args", [
    ("ode3", ["--nodes", "5", "--samples", "60", "--sample-frac", "0.5", "--lr", "1e-3", "--tol", "1e-30",
              "--n-train", "60"]),
    ("heat1d", ["--nodes", "50", "--samples", "200", "--sample-frac", "0.5", "--lr", "1e-2", "--n-train", "40"]),
    ("poisson2d_irreg_dom", ["--sample-frac", "0.25", "--lr", "1e-3", "--n-train", "30"])])
def test_subsampling_under_graph_replays_the_eager_draws(modname, args):
    """`--sample-frac < 1` with `--graph`: every replay gathers through the static index buffer that
    is refilled from the SAME np.random stream the eager loop consumes (one draw per step), for all
    problem bases (BasicODE, IBVP2D, irregular-domain Poisson2D; RegularPDE is covered by the cavity)."""
    import importlib
    mod = importlib.import_module(f"spinn_b200.problems.{modname}")
    losses = []
    for extra in ([], ["--graph"]):
        np.random.seed(0)
        torch.manual_seed(0)
        app = mod.make_app()
        app.run(args=args + ["--gpu"] + extra)
        losses.append(np.asarray(app.solver.loss))
    eager, graphed = losses
    assert len(eager) == len(graphed)
    rel = np.abs(eager - graphed) / np.abs(eager)
    assert rel[:10].max() < 5e-5 and rel.max() < 3e-3, (modname, rel[:10].max(), rel.max())
