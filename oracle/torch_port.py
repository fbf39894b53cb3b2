"""CPU oracle (TEST INFRASTRUCTURE, not product code) -- a dense, eager PyTorch
restatement of the reference's own algorithm for the hot path.

This is the "port" that ``bench.py`` times as ``cpu_baseline`` / ``--impl
reference``: it does what the reference does, the way the reference does it --
materialise the (N points x M nodes) tensors, apply the kernel, read out with a
linear layer, differentiate with three nested ``torch.autograd.grad`` calls and
back-propagate through that graph -- so its cost is the reference's cost.

Followed reference code (relative to /root/reference):
  shift/scale      code/spinn2d.py:130-135 (multiply by 1/h), code/spinn1d.py:110-112 (divide by h)
  kernels          code/spinn2d.py:191-205, code/spinn1d.py:176-189
  read-out         code/spinn2d.py:169-179, code/spinn1d.py:147-156
  jets             code/pde2d_base.py:46-62, code/ode_base.py:32-42
  backward         code/common.py:242-248

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this.
Pinned against reference-generated fixtures by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.autograd as ag
import torch.nn.functional as F


def _hat_constants(dim, dtype, device):
    # fp32 tensors in the reference (code/spinn2d.py:198-199); cast afterwards.
    k = torch.tensor([1.0 + 2.0 * dim * np.log(2.0)], dtype=torch.float32)
    fac = F.softplus(torch.tensor([1.0], dtype=torch.float32))
    return k.to(device=device, dtype=dtype), fac.to(device=device, dtype=dtype)


def _kernel(kind, *hs):
    if kind in ("gaussian", 0):
        s = hs[0] * hs[0]
        for t in hs[1:]:
            s = s + t * t
        return torch.exp(-0.5 * s)
    k, fac = _hat_constants(len(hs), hs[0].dtype, hs[0].device)
    z = k
    for t in hs:
        z = z - F.softplus(2.0 * t) - F.softplus(-2.0 * t)
    return F.softplus(z) / fac


def forward2d(kind, x, y, xc, yc, h, W, b):
    """(N,) x2, (M,) x3, (K,M), (K,)|None -> (N,K)"""
    fac = 1.0 / h
    xh = (x.unsqueeze(1) - xc) * fac
    yh = (y.unsqueeze(1) - yc) * fac
    return F.linear(_kernel(kind, xh, yh), W, b)


def forward1d(kind, x, xc, h, W, b):
    return F.linear(_kernel(kind, (x.unsqueeze(1) - xc) / h), W, b)


def _d(u, inputs):
    return ag.grad(u, inputs, grad_outputs=torch.ones_like(u),
                   retain_graph=True, create_graph=True)


def jets2d(kind, x, y, xc, yc, h, W, b):
    """x, y must require grad.  Returns list over k of (u,ux,uy,uxx,uyy,uxy)."""
    U = forward2d(kind, x, y, xc, yc, h, W, b)
    out = []
    for k in range(U.shape[1]):
        u = U[:, k]
        ux, uy = _d(u, (x, y))
        uxx, uxy = _d(ux, (x, y))
        _, uyy = _d(uy, (x, y))
        out.append((u, ux, uy, uxx, uyy, uxy))
    return out


def jets1d(kind, x, xc, h, W, b):
    U = forward1d(kind, x, xc, h, W, b)
    out = []
    for k in range(U.shape[1]):
        u = U[:, k]
        (ux,) = _d(u, (x,))
        (uxx,) = _d(ux, (x,))
        out.append((u, ux, uxx))
    return out


def jets_and_grads2d(kind, x, y, xf, yf, xfix, yfix, h, W, b, gjets, dtype=torch.float32):
    """Reference-style evaluation of jets and of the parameter gradients of
    ``L = sum g*J``.  All inputs array-like; xf/yf are the free centres,
    xfix/yfix the fixed ones; gjets (5|6, N, K).  Returns numpy dict."""
    t = lambda a, rg=False: torch.tensor(np.asarray(a), dtype=dtype).requires_grad_(rg)
    x, y = t(x, True), t(y, True)
    xf, yf, h, W = t(xf, True), t(yf, True), t(h, True), t(W, True)
    b = None if b is None else t(b, True)
    xc = torch.cat((xf, t(xfix)))
    yc = torch.cat((yf, t(yfix)))
    g = torch.tensor(np.asarray(gjets), dtype=dtype)
    J = jets2d(kind, x, y, xc, yc, h, W, b)
    L = 0.0
    for k, jk in enumerate(J):
        for a in range(g.shape[0]):
            L = L + (g[a, :, k] * jk[a]).sum()
    params = [xf, yf, h, W] + ([] if b is None else [b])
    grads = ag.grad(L, params, allow_unused=True)
    jets = torch.stack([torch.stack(jk, 0) for jk in J], -1)      # (6, N, K)
    out = {"jets": jets.detach().numpy(), "d_xc": grads[0].numpy(), "d_yc": grads[1].numpy(),
           "d_h": grads[2].numpy().reshape(-1), "d_W": grads[3].numpy()}
    if b is not None:
        out["d_b"] = grads[4].numpy()
    return out


def jets_and_grads1d(kind, x, xf, xfix, h, W, b, gjets, dtype=torch.float32):
    t = lambda a, rg=False: torch.tensor(np.asarray(a), dtype=dtype).requires_grad_(rg)
    x = t(x, True)
    xf, h, W = t(xf, True), t(h, True), t(W, True)
    b = None if b is None else t(b, True)
    xc = torch.cat((xf, t(xfix)))
    g = torch.tensor(np.asarray(gjets), dtype=dtype)
    J = jets1d(kind, x, xc, h, W, b)
    L = 0.0
    for k, jk in enumerate(J):
        for a in range(3):
            L = L + (g[a, :, k] * jk[a]).sum()
    params = [xf, h, W] + ([] if b is None else [b])
    grads = ag.grad(L, params, allow_unused=True)
    jets = torch.stack([torch.stack(jk, 0) for jk in J], -1)
    out = {"jets": jets.detach().numpy(), "d_xc": grads[0].numpy(),
           "d_h": grads[1].numpy().reshape(-1), "d_W": grads[2].numpy()}
    if b is not None:
        out["d_b"] = grads[3].numpy()
    return out


# --------------------------------------------------------------------------- #
# the timed leg: one interior fwd+bwd of the scaled Poisson2D residual on a
# chunk of points (code/pde2d_base.py:132-141 + code/poisson2d_sine.py:16-18)
# --------------------------------------------------------------------------- #
def poisson_interior_step(kind, x, y, xc, yc, h, W, b, n_total):
    """x, y: leaf fp32 tensors with requires_grad (as code/pde2d_base.py:91);
    xc.. W, b: leaf parameters.  Accumulates .grad like ``loss.backward()`` and
    returns the chunk's contribution to ``(res**2).mean()``."""
    U = forward2d(kind, x, y, xc, yc, h, W, b)
    u = U.squeeze()
    ux, uy = _d(u, (x, y))
    uxx, _ = _d(ux, (x, y))
    _, uyy = _d(uy, (x, y))
    f = 20 * np.pi ** 2 * torch.sin(2 * np.pi * x) * torch.sin(4 * np.pi * y)
    res = 0.1 * (uxx + uyy + f)
    loss = (res ** 2).sum() / n_total
    loss.backward()
    return float(loss.detach())


def ode3_interior_step(kind, x, xc, h, W, b, n_total, Kc=0.01):
    """code/ode3.py:12-16 + code/ode_base.py:32-42,83-92: dense 1-D forward, two nested
    autograd.grad, res = u_xx + f(x), loss = sum(res^2)/n_total, backward.  Returns the loss."""
    u = forward1d(kind, x, xc, h, W, b).squeeze(-1)
    (ux,) = _d(u, (x,))
    (uxx,) = _d(ux, (x,))
    f = -(Kc * (-6 * x + 4 / 3.) + x * (2 * x - 2. / 3) ** 2) * torch.exp(-(x - 1 / 3.) ** 2 / Kc) / Kc ** 2
    res = uxx + f
    loss = (res ** 2).sum() / n_total
    loss.backward()
    return float(loss.detach())


def cavity_interior_step(kind, x, y, xc, yc, h, W, b, nu=0.01):
    """code/cavity.py:84-100: U = nn(xs, ys) with 3 outputs; jets of u and v through
    _compute_derivatives, gradient of p; continuity + two momentum residuals, SUMMED; backward."""
    U = forward2d(kind, x, y, xc, yc, h, W, b)
    u, v, p = U[:, 0], U[:, 1], U[:, 2]
    ux, uy = _d(u, (x, y))
    uxx, _ = _d(ux, (x, y))
    _, uyy = _d(uy, (x, y))
    vx, vy = _d(v, (x, y))
    vxx, _ = _d(vx, (x, y))
    _, vyy = _d(vy, (x, y))
    px, py = _d(p, (x, y))
    ce = ((ux + vy) ** 2).sum()
    mex = ((u * ux + v * uy + px - nu * (uxx + uyy)) ** 2).sum()
    mey = ((u * vx + v * vy + py - nu * (vxx + vyy)) ** 2).sum()
    loss = ce + (mex + mey)
    loss.backward()
    return float(loss.detach())


# --------------------------------------------------------------------------- #
# nn.Module form of the same dense algorithm (same parameter names as the
# reference: layer1.x/.y/.h, layer2.weight/.bias; 1-D layer1.center), so that the
# problem classes / Optimizer of spinn_b200 can drive the REFERENCE ALGORITHM on
# any device.  Used by tests and by tests/perf/train_step_bench.py as the "reference
# eager path on the same GPU" (SURVEY.md section 6).  Follows code/spinn2d.py:97-188
# and code/spinn1d.py:88-171.
# --------------------------------------------------------------------------- #
class _Shift2D(torch.nn.Module):
    def __init__(self, points, fixed, device):
        super().__init__()
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float32, device=device)
        px, fx = np.asarray(points[0], float), np.asarray(fixed[0], float)
        self.n = n = px.size + fx.size
        span = np.ptp(np.hstack((px, fx))) if fx.size else np.ptp(px)
        self.x = torch.nn.Parameter(t(points[0]))
        self.y = torch.nn.Parameter(t(points[1]))
        self.xf, self.yf = t(fixed[0]), t(fixed[1])
        self.h = torch.nn.Parameter((span / np.sqrt(n)) * torch.ones(n, device=device))

    def centers(self):
        return torch.cat((self.x, self.xf)), torch.cat((self.y, self.yf))


class DenseSPINN2D(torch.nn.Module):
    """Dense eager 2-D SPINN: materialises (N, M) tensors; autograd does the rest."""

    @classmethod
    def from_args(cls, pde, activation, args):
        return cls(pde, activation)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        parser.add_argument("--fixed-h", dest="fixed_h", action="store_true", default=False)
        parser.add_argument("--pu", dest="pu", action="store_true", default=False)

    def __init__(self, pde, activation, device=None):
        super().__init__()
        from spinn_b200 import common
        device = common.device() if device is None else device
        self.kind = "gaussian" if getattr(activation, "__name__", "") == "gaussian" else "softplus"
        self.layer1 = _Shift2D(pde.nodes(), pde.fixed_nodes(), device)
        self.layer2 = torch.nn.Linear(self.layer1.n, pde.n_vars()).to(device)
        self.layer2.weight.data.fill_(0.0)
        self.layer2.bias.data.fill_(0.0)

    def forward(self, x, y):
        xc, yc = self.layer1.centers()
        return forward2d(self.kind, x, y, xc, yc, self.layer1.h, self.layer2.weight, self.layer2.bias).squeeze()

    def centers(self):
        return self.layer1.centers()

    def widths(self):
        return self.layer1.h

    def weights(self):
        return self.layer2.weight


class _Shift1D(torch.nn.Module):
    def __init__(self, points, fixed, device):
        super().__init__()
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float32, device=device)
        p, f = np.asarray(points, float).ravel(), np.asarray(fixed, float).ravel()
        self.n = n = p.size + f.size
        allx = np.hstack((p, f)) if f.size else p
        self.center = torch.nn.Parameter(t(p))
        self.fixed = t(f)
        self.h = torch.nn.Parameter(((allx.max() - allx.min()) / n) * torch.ones(n, device=device))


class DenseSPINN1D(torch.nn.Module):
    @classmethod
    def from_args(cls, pde, activation, args):
        return cls(pde, activation)

    @classmethod
    def setup_argparse(cls, parser, **kw):
        parser.add_argument("--fixed-h", dest="fixed_h", action="store_true", default=False)
        parser.add_argument("--pu", dest="pu", action="store_true", default=False)

    def __init__(self, pde, activation, device=None):
        super().__init__()
        from spinn_b200 import common
        device = common.device() if device is None else device
        self.kind = "gaussian" if getattr(activation, "__name__", "") == "gaussian" else "softplus"
        self.layer1 = _Shift1D(pde.nodes(), pde.fixed_nodes(), device)
        self.layer2 = torch.nn.Linear(self.layer1.n, pde.n_vars()).to(device)
        self.layer2.weight.data.fill_(0.0)
        self.layer2.bias.data.fill_(0.0)

    def forward(self, x):
        xc = torch.cat((self.layer1.center, self.layer1.fixed))
        return forward1d(self.kind, x, xc, self.layer1.h, self.layer2.weight, self.layer2.bias).squeeze()

    def centers(self):
        return torch.cat((self.layer1.center, self.layer1.fixed))

    def widths(self):
        return self.layer1.h

    def weights(self):
        return self.layer2.weight
