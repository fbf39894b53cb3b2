#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed and are what travels to the GPU box.  Nothing in tests/, smoke()
or bench.py imports this file or reads /root/reference at run time.

    python tests/golden/make_golden.py            # regenerate everything

What it records
  op_*.npz     op-level: reference SPINN1D/SPINN2D (code/spinn1d.py, code/spinn2d.py)
               with seeded, perturbed parameters -> nn(x[,y]) -> the reference's
               own _compute_derivatives (code/pde2d_base.py:46-62, code/ode_base.py:32-42)
               -> L = sum g*J -> autograd parameter gradients, in fp32 (the
               reference's arithmetic) and in fp64 (same graph, cast by hand).
  traj_*.npz   trajectory-level: the four BASELINE configs driven through the
               reference's App.run for a few hundred Adam steps: loss history,
               final parameters, final error norms.
  mesh_data.npz  the irregular-domain point sets of code/mesh_data/*.dat.
  constants.npz  SoftPlus.k / SoftPlus.fac as the reference builds them.
"""
import io
import os
import sys
import types
import contextlib

import numpy as np
import torch
import torch.autograd as ag

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """matplotlib / mayavi are not installed here; the reference imports them
    at module top level (code/spinn1d.py:4, code/spinn2d.py:3) -> empty stubs."""
    for name in ("matplotlib", "matplotlib.pyplot", "mayavi", "mayavi.mlab"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mayavi"].mlab = sys.modules["mayavi.mlab"]
    if REF not in sys.path:
        sys.path.insert(0, REF)


class StubPDE:
    def __init__(self, nodes, fixed, k):
        self._n, self._f, self._k = nodes, fixed, k

    def nodes(self):
        return self._n

    def fixed_nodes(self):
        return self._f

    def n_vars(self):
        return self._k


def _to_double(nn, act):
    """.double() does not touch plain-attribute tensors."""
    nn = nn.double()
    l1 = nn.layer1
    for name in ("xf", "yf", "fixed"):
        if hasattr(l1, name):
            setattr(l1, name, getattr(l1, name).double())
    if hasattr(act, "k"):
        act.k = act.k.double()
        act.fac = act.fac.double()
    return nn


def _perturb(nn, gen, dim):
    """zero-initialised weights make centre/width gradients vanish -> perturb."""
    l1, l2 = nn.layer1, nn.layer2
    with torch.no_grad():
        l2.weight.copy_(0.3 * torch.randn(l2.weight.shape, generator=gen))
        if l2.bias is not None:
            l2.bias.copy_(0.1 * torch.randn(l2.bias.shape, generator=gen))
        l1.h.mul_(1.0 + 0.3 * torch.rand(l1.h.shape, generator=gen))
        if dim == 2:
            l1.x.add_(0.01 * torch.randn(l1.x.shape, generator=gen))
            l1.y.add_(0.01 * torch.randn(l1.y.shape, generator=gen))
        else:
            l1.center.add_(0.01 * torch.randn(l1.center.shape, generator=gen))


def op_case_2d(name, kind, n_side, nb, npts, K, seed, fixed_h=False, lo=0.0, hi=1.0, use_pu=False):
    import spinn2d
    import pde2d_base
    gen = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    g1 = np.linspace(lo + 0.05, hi - 0.05, n_side)
    xn, yn = (a.ravel() for a in np.meshgrid(g1, g1, indexing="ij"))
    if nb:
        e = np.linspace(lo, hi, nb)
        o = np.ones_like(e)
        fx = np.hstack((e, hi * o, e, lo * o))
        fy = np.hstack((lo * o, e, hi * o, e))
        fixed = (fx, fy)
    else:
        fixed = ([], [])
    xs = (lo + (hi - lo) * rs.rand(npts)).astype(np.float32)
    ys = (lo + (hi - lo) * rs.rand(npts)).astype(np.float32)
    gj = rs.randn(6, npts, K).astype(np.float32)
    gj[5] = 0.0  # the reference never consumes u_xy
    out = {"kind": kind, "K": K, "dim": 2}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        act = spinn2d.gaussian if kind == "gaussian" else spinn2d.SoftPlus()
        nn = spinn2d.SPINN2D(StubPDE((xn, yn), fixed, K), act, fixed_h=fixed_h, use_pu=use_pu)
        _perturb(nn, torch.Generator().manual_seed(seed + 1), 2)
        if tag == "f32":
            out.update(
                x=xs, y=ys, gjets=gj,
                xf=nn.layer1.x.detach().numpy().copy(), yf=nn.layer1.y.detach().numpy().copy(),
                xfix=nn.layer1.xf.numpy().copy(), yfix=nn.layer1.yf.numpy().copy(),
                h=nn.layer1.h.detach().numpy().reshape(-1).copy(),
                W=nn.layer2.weight.detach().numpy().copy(),
                b=(np.zeros(K, np.float32) if use_pu else nn.layer2.bias.detach().numpy().copy()),
            )
        else:
            nn = _to_double(nn, act)
        x = torch.tensor(xs, dtype=dt, requires_grad=True)
        y = torch.tensor(ys, dtype=dt, requires_grad=True)
        U = nn(x, y)
        U2 = U.reshape(npts, K)
        jets = []
        for k in range(K):
            u = U2[:, k]
            # the reference routine, verbatim call (it is self-free)
            u_, ux, uy, uxx, uyy = pde2d_base.RegularPDE._compute_derivatives(None, u, x, y)
            uxy = ag.grad(ux, y, torch.ones_like(ux), retain_graph=True, create_graph=True)[0]
            jets.append(torch.stack((u_, ux, uy, uxx, uyy, uxy)))
        J = torch.stack(jets, -1)                                   # (6, N, K)
        L = (torch.tensor(gj, dtype=dt) * J).sum()
        params = [nn.layer1.x, nn.layer1.y, nn.layer1.h, nn.layer2.weight]
        if not use_pu:
            params.append(nn.layer2.bias)
        gr = ag.grad(L, params)
        out["jets_" + tag] = J.detach().numpy()
        for nme, gv in zip(("d_xc", "d_yc", "d_h", "d_W", "d_b"), gr):
            out[nme + "_" + tag] = gv.numpy().reshape(-1) if nme == "d_h" else gv.numpy()
    out["use_pu"] = use_pu
    np.savez_compressed(os.path.join(HERE, f"{'pu' if use_pu else 'op'}_{name}.npz"), **out)
    print("wrote", name)


def dense_baseline_case(kind, npts=2048, seed=41):
    """BASELINE density (judge round 1, weak 1-ii): the 4096-node state of the synthetic workload --
    what `poisson2d_sine.py --nodes 3600 --b-nodes 124` builds (60x60 free + 4x124 fixed nodes,
    h = 1/64), with the seeded perturbation of SURVEY.md 8(d) (spinn_b200.synthetic.perturb_) --
    evaluated BY THE REFERENCE at `npts` points drawn from the 2048 x 2048 collocation grid, in fp32
    and fp64: jets through RegularPDE._compute_derivatives, parameter gradients through autograd for
    (i) a random 5-row upstream and (ii) a Laplacian-type upstream (rows u_xx = u_yy only, as the
    Poisson residual produces)."""
    import spinn2d
    import poisson2d_sine
    import pde2d_base
    rs = np.random.RandomState(seed)
    ij = rs.randint(0, 2048, size=(2, npts))
    xs = ((ij[0] + 0.5) / 2048.0).astype(np.float32)
    ys = ((ij[1] + 0.5) / 2048.0).astype(np.float32)
    gj = (rs.randn(5, npts) / npts).astype(np.float32)
    gl = np.zeros((5, npts), np.float32)
    gl[3] = gl[4] = (rs.randn(npts) / npts).astype(np.float32)
    out = {"kind": kind, "K": 1, "dim": 2, "x": xs, "y": ys, "gjets": gj, "gjets_lapl": gl}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        pde = poisson2d_sine.Poisson2D(3600, 64, 124, None)
        act = spinn2d.gaussian if kind == "gaussian" else spinn2d.SoftPlus()
        nn = spinn2d.SPINN2D(pde, act)
        g = torch.Generator().manual_seed(0)                  # == spinn_b200.synthetic.perturb_(nn, 0)
        l1, l2 = nn.layer1, nn.layer2
        with torch.no_grad():
            l2.weight.copy_(0.3 * torch.randn(l2.weight.shape, generator=g))
            l1.h.mul_(1.0 + 0.3 * torch.rand(l1.h.shape, generator=g))
            l1.x.add_(0.01 * torch.randn(l1.x.shape, generator=g))
            l1.y.add_(0.01 * torch.randn(l1.y.shape, generator=g))
            l2.bias.copy_(0.1 * torch.randn(l2.bias.shape, generator=g))
        if tag == "f32":
            out.update(xf=l1.x.detach().numpy().copy(), yf=l1.y.detach().numpy().copy(),
                       xfix=l1.xf.numpy().copy(), yfix=l1.yf.numpy().copy(),
                       h=l1.h.detach().numpy().copy(), W=l2.weight.detach().numpy().copy(),
                       b=l2.bias.detach().numpy().copy())
        else:
            nn = _to_double(nn, act)
        x = torch.tensor(xs, dtype=dt, requires_grad=True)
        y = torch.tensor(ys, dtype=dt, requires_grad=True)
        u = nn(x, y)
        J = torch.stack(pde2d_base.RegularPDE._compute_derivatives(None, u, x, y))      # (5, N)
        out["jets_" + tag] = J.detach().numpy()
        params = [nn.layer1.x, nn.layer1.y, nn.layer1.h, nn.layer2.weight, nn.layer2.bias]
        for suffix, gg in (("", gj), ("_lapl", gl)):
            L = (torch.tensor(gg, dtype=dt) * J).sum()
            gr = ag.grad(L, params, retain_graph=True)
            for nme, gv in zip(("d_xc", "d_yc", "d_h", "d_W", "d_b"), gr):
                out[nme + suffix + "_" + tag] = gv.numpy().reshape(-1) if nme == "d_h" else gv.numpy()
    np.savez_compressed(os.path.join(HERE, f"dense_{kind}.npz"), **out)
    print("wrote dense", kind)


def op_case_1d(name, kind, n, npts, K, seed, fixed=True, use_pu=False):
    import spinn1d
    import ode_base
    rs = np.random.RandomState(seed)
    xn = np.asarray([i / (n + 1) for i in range(1, n + 1)])
    fx = np.array([0.0, 1.0]) if fixed else np.array([])
    xs = rs.rand(npts).astype(np.float32)
    gj = rs.randn(3, npts, K).astype(np.float32)
    out = {"kind": kind, "K": K, "dim": 1}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        act = spinn1d.gaussian if kind == "gaussian" else spinn1d.SoftPlus()
        nn = spinn1d.SPINN1D(StubPDE(xn, fx, K), act, use_pu=use_pu)
        _perturb(nn, torch.Generator().manual_seed(seed + 1), 1)
        if tag == "f32":
            out.update(
                x=xs, gjets=gj, xf=nn.layer1.center.detach().numpy().copy(),
                xfix=nn.layer1.fixed.numpy().copy(), h=nn.layer1.h.detach().numpy().copy(),
                W=nn.layer2.weight.detach().numpy().copy(),
                b=(np.zeros(K, np.float32) if use_pu else nn.layer2.bias.detach().numpy().copy()),
            )
        else:
            nn = _to_double(nn, act)
        x = torch.tensor(xs, dtype=dt, requires_grad=True)
        if tag == "f64":
            # SPINN1D.forward builds an fp32 tensor(1.0) (code/spinn1d.py:154); dividing a
            # double tensor by a 0-dim float tensor keeps double, so no patching needed.
            pass
        U2 = nn(x).reshape(npts, K)
        jets = []
        for k in range(K):
            u_, ux, uxx = ode_base.BasicODE._compute_derivatives(None, U2[:, k], x)
            jets.append(torch.stack((u_, ux, uxx)))
        J = torch.stack(jets, -1)
        L = (torch.tensor(gj, dtype=dt) * J).sum()
        params = [nn.layer1.center, nn.layer1.h, nn.layer2.weight]
        if not use_pu:
            params.append(nn.layer2.bias)
        gr = ag.grad(L, params)
        out["jets_" + tag] = J.detach().numpy()
        for nme, gv in zip(("d_xc", "d_h", "d_W", "d_b"), gr):
            out[nme + "_" + tag] = gv.numpy()
    out["use_pu"] = use_pu
    np.savez_compressed(os.path.join(HERE, f"{'pu' if use_pu else 'op'}_{name}.npz"), **out)
    print("wrote", name)


def trajectory(name, module, app_attr, args, seed=0):
    """Drive a reference config through its own App.run for a few steps."""
    import importlib
    np.random.seed(seed)
    torch.manual_seed(seed)
    mod = importlib.import_module(module)
    app = app_attr(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        app.run(args=args)
    sd = {k: v.detach().cpu().numpy() for k, v in app.nn.state_dict().items()}
    err = app.plotter.get_error()
    out = {"loss": np.asarray(app.solver.loss, dtype=np.float64),
           "errors": np.asarray(err, dtype=np.float64),
           "args": np.asarray(args)}
    out.update({"sd_" + k: v for k, v in sd.items()})
    if hasattr(app.pde, "u0"):            # FD-SPINN time stepping (code/fd_spinn1d_base.py)
        out.update(t=app.pde.t, iter=app.pde.iter, u0=app.pde.u0.detach().cpu().numpy())
    np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), **out)
    print("wrote traj", name, "final loss", app.solver.loss[-1], "err", err)


def mesh_data():
    out = {}
    for f in ("interior_nodes", "boundary_nodes", "interior_samples", "boundary_samples"):
        with open(os.path.join(REF, "mesh_data", f + ".dat")) as fh:
            n = int(fh.readline().split()[0])
            rows = [fh.readline().split()[:2] for _ in range(n)]
        out[f] = np.asarray(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "mesh_data.npz"), **out)
    print("wrote mesh_data", {k: v.shape for k, v in out.items()})


def constants():
    import spinn1d
    import spinn2d
    a1, a2 = spinn1d.SoftPlus(), spinn2d.SoftPlus()
    np.savez(os.path.join(HERE, "constants.npz"),
             k1=a1.k.numpy(), fac1=a1.fac.numpy(), k2=a2.k.numpy(), fac2=a2.fac.numpy())


def main():
    load_reference()
    torch.set_num_threads(8)
    constants()
    mesh_data()
    op_case_2d("g2d_k1", "gaussian", 4, 3, 61, 1, 11)
    op_case_2d("s2d_k1", "softplus", 4, 3, 61, 1, 12)
    op_case_2d("g2d_k3", "gaussian", 3, 4, 50, 3, 13)
    op_case_2d("s2d_k3", "softplus", 3, 4, 50, 3, 14)
    op_case_2d("g2d_nofixed", "gaussian", 5, 0, 33, 1, 15)
    op_case_2d("g2d_one_point", "gaussian", 3, 2, 1, 1, 16)
    op_case_2d("g2d_fixed_h", "gaussian", 4, 3, 40, 1, 17, fixed_h=True)
    op_case_2d("s2d_wide", "softplus", 6, 5, 130, 1, 18, lo=-1.5, hi=2.0)
    op_case_2d("g2d_wide", "gaussian", 6, 5, 130, 2, 19, lo=-1.5, hi=2.0)
    op_case_1d("g1d_k1", "gaussian", 7, 45, 1, 21)
    op_case_1d("s1d_k1", "softplus", 7, 45, 1, 22)
    op_case_1d("g1d_k2", "gaussian", 5, 30, 2, 23)
    op_case_1d("s1d_nofixed", "softplus", 6, 1, 1, 24, fixed=False)

    from_mod = lambda cls_name, app_name, nn_name, pl_name: (
        lambda m: getattr(m, app_name)(pde_cls=getattr(m, cls_name), nn_cls=getattr(m, nn_name),
                                       plotter_cls=getattr(m, pl_name)))
    trajectory("ode3", "ode3", from_mod("ODESimple", "App1D", "SPINN1D", "Plotter1D"),
               ["--nodes", "3", "--samples", "60", "--n-train", "300", "--lr", "1e-3", "--tol", "1e-3"])
    trajectory("ode3_softplus", "ode3", from_mod("ODESimple", "App1D", "SPINN1D", "Plotter1D"),
               ["--nodes", "3", "--samples", "60", "--n-train", "200", "--lr", "1e-3", "--tol", "1e-3",
                "--activation", "softplus"])
    trajectory("poisson2d_sine", "poisson2d_sine", from_mod("Poisson2D", "App2D", "SPINN2D", "Plotter2D"),
               ["--nodes", "100", "--samples", "400", "--n-train", "300", "--lr", "1e-3", "--tol", "1e-3"])
    trajectory("poisson2d_sine_softplus", "poisson2d_sine",
               from_mod("Poisson2D", "App2D", "SPINN2D", "Plotter2D"),
               ["--nodes", "100", "--samples", "400", "--n-train", "150", "--lr", "1e-3", "--tol", "1e-3",
                "--activation", "softplus"])
    trajectory("irreg", "poisson2d_irreg_dom", from_mod("Poisson2D", "App2D", "SPINN2D", "PointCloud"),
               ["--n-train", "60", "--lr", "1e-3"])
    # two more callers on the same surface (space-time problems on IBVP2D, code/ibvp2d_base.py)
    trajectory("heat1d", "heat1d", from_mod("Heat1D", "App2D", "SPINN2D", "HeatPlotter"),
               ["--nodes", "50", "--samples", "200", "--lr", "1e-2", "--n-train", "60"])
    trajectory("burgers1d", "burgers1d", from_mod("Burgers1D", "App2D", "SPINN2D", "MyPlotter"),
               ["--nodes", "40", "--samples", "200", "--lr", "1e-3", "--n-train", "60"])
    trajectory("cavity", "cavity", from_mod("CavityPDE", "App2D", "SPINN2D", "CavityPlotter"),
               ["--nodes", "100", "--samples", "2500", "--sample-frac", "0.1", "--lr", "5e-4",
                "--n-train", "100"])
    only(["dense", "pu", "traj:ode3_lbfgs", "traj:sine_pu", "traj:sine_fixed_h", "traj:ode3_tolstop",
          "traj:heat1d_fd", "traj:burgers1d_fd", "traj:advection1d_fd"])


def only(names):
    """Regenerate a subset:  python tests/golden/make_golden.py traj:heat1d traj:burgers1d"""
    load_reference()
    from_mod = lambda cls_name, app_name, nn_name, pl_name: (
        lambda m: getattr(m, app_name)(pde_cls=getattr(m, cls_name), nn_cls=getattr(m, nn_name),
                                       plotter_cls=getattr(m, pl_name)))
    table = {
        "traj:heat1d": lambda: trajectory("heat1d", "heat1d", from_mod("Heat1D", "App2D", "SPINN2D", "HeatPlotter"),
                                          ["--nodes", "50", "--samples", "200", "--lr", "1e-2", "--n-train", "60"]),
        "traj:burgers1d": lambda: trajectory("burgers1d", "burgers1d",
                                             from_mod("Burgers1D", "App2D", "SPINN2D", "MyPlotter"),
                                             ["--nodes", "40", "--samples", "200", "--lr", "1e-3", "--n-train", "60"]),
    }
    table["dense"] = lambda: (dense_baseline_case("gaussian"), dense_baseline_case("softplus"))
    table["pu"] = lambda: (op_case_2d("g2d_k1", "gaussian", 4, 3, 61, 1, 31, use_pu=True),
                           op_case_2d("s2d_k1", "softplus", 4, 3, 61, 1, 32, use_pu=True),
                           op_case_2d("g2d_k3", "gaussian", 3, 4, 50, 3, 33, use_pu=True),
                           op_case_1d("g1d_k1", "gaussian", 7, 45, 1, 34, use_pu=True),
                           op_case_1d("s1d_k1", "softplus", 7, 45, 1, 35, use_pu=True))
    table["traj:ode3_lbfgs"] = lambda: trajectory(
        "ode3_lbfgs", "ode3", from_mod("ODESimple", "App1D", "SPINN1D", "Plotter1D"),
        ["--nodes", "3", "--samples", "60", "--n-train", "6", "--lr", "1e-2", "--tol", "1e-30",
         "--optimizer", "LBFGS"])
    table["traj:sine_pu"] = lambda: trajectory(
        "poisson2d_sine_pu", "poisson2d_sine", from_mod("Poisson2D", "App2D", "SPINN2D", "Plotter2D"),
        ["--nodes", "100", "--samples", "400", "--n-train", "60", "--lr", "1e-3", "--tol", "1e-30", "--pu"])
    table["traj:sine_fixed_h"] = lambda: trajectory(
        "poisson2d_sine_fixed_h", "poisson2d_sine", from_mod("Poisson2D", "App2D", "SPINN2D", "Plotter2D"),
        ["--nodes", "100", "--samples", "400", "--n-train", "60", "--lr", "1e-3", "--tol", "1e-30", "--fixed-h"])
    table["traj:ode3_tolstop"] = lambda: trajectory(
        "ode3_tolstop", "ode3", from_mod("ODESimple", "App1D", "SPINN1D", "Plotter1D"),
        ["--nodes", "3", "--samples", "60", "--n-train", "2000", "--lr", "1e-3", "--tol", "2e-2",
         "--n-skip", "20"])
    # FD-SPINN outer loop: a few time levels, each a short solve() on the same optimiser
    fd = lambda cls_name: from_mod(cls_name, "AppFD1D", "SPINN1D", "FDPlotter1D")
    fd_mod = lambda m: (setattr(m, "AppFD1D", __import__("fd_spinn1d_base").AppFD1D),
                        setattr(m, "FDPlotter1D", __import__("fd_spinn1d_base").FDPlotter1D), m)[-1]
    table["traj:heat1d_fd"] = lambda: trajectory(
        "heat1d_fd", "heat1d_fd_spinn", lambda m: fd("Heat1D")(fd_mod(m)),
        ["--nodes", "10", "--samples", "100", "--dt", "0.01", "-T", "0.03", "--lr", "1e-3", "--tol", "1.2e-2",
         "--n-train", "150", "--n-skip", "50"])
    table["traj:burgers1d_fd"] = lambda: trajectory(
        "burgers1d_fd", "burgers1d_fd_spinn", lambda m: fd("Burgers1D")(fd_mod(m)),
        ["--nodes", "40", "--samples", "400", "--dt", "0.01", "-T", "0.02", "--lr", "1e-4", "--tol", "5e-6",
         "--n-train", "100", "--n-skip", "50"])
    table["traj:advection1d_fd"] = lambda: trajectory(
        "advection1d_fd", "advection1d_fd_spinn", lambda m: fd("Advection1D")(fd_mod(m)),
        ["--nodes", "20", "--samples", "500", "--sample-frac", "0.5", "--dt", "0.0025", "-T", "0.005",
         "--lr", "1e-4", "--tol", "1e-6", "--n-train", "100"])
    for n in names:
        table[n]()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        only(sys.argv[1:])
    else:
        main()
