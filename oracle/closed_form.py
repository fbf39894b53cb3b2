"""CPU oracle (TEST INFRASTRUCTURE, not product code) -- closed-form restatement
of the SPINN hot path in NumPy.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import anything under ``oracle/``.  The product
package ``spinn_b200`` never does.

What is restated (reference files relative to /root/reference):

* the shift/scale layer ``xh=(x-xc)*(1/h)``            code/spinn2d.py:130-135,
  ``(x-c)/h``                                           code/spinn1d.py:110-112
* the kernels ``gaussian``                              code/spinn2d.py:191-192, code/spinn1d.py:176-177
  and softplus-hat ``SoftPlus.__call__``                code/spinn2d.py:195-205, code/spinn1d.py:180-189
  (fp32-rounded constants ``k`` and ``fac`` as the reference builds them)
* the read-out ``layer2`` (``z @ W.T + b``)             code/spinn2d.py:178, code/spinn1d.py:155
* the jets ``u, u_x, u_y, u_xx, u_yy`` (and the mixed ``u_xy`` the reference
  computes and discards) that ``_compute_derivatives`` obtains with three nested
  ``torch.autograd.grad`` calls                         code/pde2d_base.py:46-62, code/ode_base.py:32-42
* the parameter gradients ``loss.backward()`` produces code/common.py:242-248

The reference has NO golden vectors or tests for this path,
so parity is pinned by fixtures generated from the reference itself, executed
in the build container: ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against them.

The derivative formulas are analytic (no autograd): for a kernel
``phi((x-c)/h, (y-d)/h)`` translation invariance gives ``d/dc = -d/dx`` and Euler
scaling gives ``d/dh phi_a = -(1/h)(|a| phi_a + X phi_ax + Y phi_ay)``.
"""
from __future__ import annotations

import numpy as np

GAUSSIAN = 0
SOFTPLUS = 1

_KINDS = {"gaussian": GAUSSIAN, "softplus": SOFTPLUS, GAUSSIAN: GAUSSIAN, SOFTPLUS: SOFTPLUS}


def kind_id(kind):
    return _KINDS[kind]


# --------------------------------------------------------------------------- #
# constants exactly as the reference builds them (fp32-rounded, then used as-is)
# --------------------------------------------------------------------------- #
def softplus_constants(dim):
    """``k`` and ``fac`` of ``SoftPlus.__init__`` (code/spinn2d.py:198-199,
    code/spinn1d.py:183-184): both live in fp32 tensors in the reference."""
    k = np.float32(1.0 + 2.0 * dim * np.log(2.0))
    # torch's fp32 softplus(1.0): log1p(e) = 1.31326168751822 lies half-way between two
    # floats and torch lands on the lower one, 0x3FA818F5 (recorded from the reference
    # in tests/golden/constants.npz).
    fac = np.array([0x3FA818F5], dtype=np.uint32).view(np.float32)[0]
    return float(k), float(fac)


def _softplus(z):
    # nn.Softplus(beta=1, threshold=20): linear above the threshold.
    return np.where(z > 20.0, z, np.log1p(np.exp(np.minimum(z, 20.0))))


def _sigmoid(z):
    e = np.exp(-np.abs(z))
    return np.where(z >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


# --------------------------------------------------------------------------- #
# kernel value and all partial derivatives w.r.t. the point up to order three
# --------------------------------------------------------------------------- #
def _phi_derivs_2d(kind, X, Y, r):
    """X, Y: (n, M) offsets x-c, y-d; r: (M,) = 1/h.
    Returns dict with keys '', x, y, xx, yy, xy, xxx, xyy, xxy, yyy."""
    rho = r * r
    r3 = rho * r
    if kind == GAUSSIAN:
        a = rho * X
        b = rho * Y
        phi = np.exp(-0.5 * (a * X + b * Y))
        A2 = a * a - rho
        B2 = b * b - rho
        return {
            "": phi, "x": -a * phi, "y": -b * phi,
            "xx": A2 * phi, "yy": B2 * phi, "xy": a * b * phi,
            "xxx": (3.0 * rho * a - a ** 3) * phi, "yyy": (3.0 * rho * b - b ** 3) * phi,
            "xyy": -a * B2 * phi, "xxy": -A2 * b * phi,
        }
    k, fac = softplus_constants(2)
    s = X * r
    t = Y * r
    z = k - _softplus(2 * s) - _softplus(-2 * s) - _softplus(2 * t) - _softplus(-2 * t)
    S = _softplus(z) / fac
    sg = _sigmoid(z) / fac
    s1 = _sigmoid(z) * (1 - _sigmoid(z)) / fac
    s2 = s1 * (1 - 2 * _sigmoid(z))
    Ts, Tt = np.tanh(s), np.tanh(t)
    Cs, Ct = 1 - Ts * Ts, 1 - Tt * Tt
    zx, zy = -2 * r * Ts, -2 * r * Tt
    zxx, zyy = -2 * rho * Cs, -2 * rho * Ct
    zxxx, zyyy = 4 * r3 * Ts * Cs, 4 * r3 * Tt * Ct
    return {
        "": S, "x": sg * zx, "y": sg * zy,
        "xx": s1 * zx * zx + sg * zxx, "yy": s1 * zy * zy + sg * zyy, "xy": s1 * zx * zy,
        "xxx": s2 * zx ** 3 + 3 * s1 * zx * zxx + sg * zxxx,
        "yyy": s2 * zy ** 3 + 3 * s1 * zy * zyy + sg * zyyy,
        "xyy": s2 * zx * zy * zy + s1 * zx * zyy,
        "xxy": s2 * zx * zx * zy + s1 * zxx * zy,
    }


def _phi_derivs_1d(kind, X, r):
    rho = r * r
    r3 = rho * r
    if kind == GAUSSIAN:
        a = rho * X
        phi = np.exp(-0.5 * a * X)
        return {"": phi, "x": -a * phi, "xx": (a * a - rho) * phi,
                "xxx": (3.0 * rho * a - a ** 3) * phi}
    k, fac = softplus_constants(1)
    s = X * r
    z = k - _softplus(2 * s) - _softplus(-2 * s)
    sig = _sigmoid(z)
    S = _softplus(z) / fac
    sg = sig / fac
    s1 = sig * (1 - sig) / fac
    s2 = s1 * (1 - 2 * sig)
    T = np.tanh(s)
    C = 1 - T * T
    zx, zxx, zxxx = -2 * r * T, -2 * rho * C, 4 * r3 * T * C
    return {"": S, "x": sg * zx, "xx": s1 * zx * zx + sg * zxx,
            "xxx": s2 * zx ** 3 + 3 * s1 * zx * zxx + sg * zxxx}


def _prep(dtype, *arrs):
    return [None if a is None else np.asarray(a, dtype=dtype) for a in arrs]


def _widths(h, M, dtype):
    h = np.asarray(h, dtype=dtype).reshape(-1)
    if h.size == 1:  # --fixed-h: one shared scalar width (code/spinn2d.py:116-117)
        h = np.full(M, h[0], dtype=dtype)
    return h


JETS_2D = ("", "x", "y", "xx", "yy", "xy")
JETS_1D = ("", "x", "xx")


# --------------------------------------------------------------------------- #
# forward jets
# --------------------------------------------------------------------------- #
def jets2d(kind, x, y, xc, yc, h, W, b=None, dtype=np.float64, chunk=1024):
    """Jets of the 2-D SPINN at N points.

    x, y: (N,) points.  xc, yc: (M,) centres (free ++ fixed, code/spinn2d.py:121-122).
    h: (M,) or scalar.  W: (K, M).  b: (K,) or None (``--pu`` has no bias).
    Returns (6, N, K): u, u_x, u_y, u_xx, u_yy, u_xy.
    """
    kind = kind_id(kind)
    x, y, xc, yc, W, b = _prep(dtype, x, y, xc, yc, W, b)
    M = xc.shape[0]
    W = W.reshape(-1, M)
    K = W.shape[0]
    r = 1.0 / _widths(h, M, dtype)
    N = x.shape[0]
    out = np.zeros((6, N, K), dtype=dtype)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        X = x[s:e, None] - xc[None, :]
        Y = y[s:e, None] - yc[None, :]
        d = _phi_derivs_2d(kind, X, Y, r)
        for a, key in enumerate(JETS_2D):
            out[a, s:e] = d[key] @ W.T
    if b is not None:
        out[0] += b[None, :]
    return out


def jets1d(kind, x, xc, h, W, b=None, dtype=np.float64, chunk=4096):
    """Jets of the 1-D SPINN: returns (3, N, K): u, u_x, u_xx."""
    kind = kind_id(kind)
    x, xc, W, b = _prep(dtype, x, xc, W, b)
    M = xc.shape[0]
    W = W.reshape(-1, M)
    K = W.shape[0]
    r = 1.0 / _widths(h, M, dtype)
    N = x.shape[0]
    out = np.zeros((3, N, K), dtype=dtype)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        d = _phi_derivs_1d(kind, x[s:e, None] - xc[None, :], r)
        for a, key in enumerate(JETS_1D):
            out[a, s:e] = d[key] @ W.T
    if b is not None:
        out[0] += b[None, :]
    return out


# --------------------------------------------------------------------------- #
# parameter gradients of L = sum_{a,i,k} g[a,i,k] * J[a,i,k]
# --------------------------------------------------------------------------- #
def grads2d(kind, x, y, xc, yc, h, W, gjets, n_free=None, dtype=np.float64, chunk=1024):
    """gjets: (5 or 6, N, K) upstream gradients for u,u_x,u_y,u_xx,u_yy[,u_xy].

    Returns dict d_xc (n_free,), d_yc (n_free,), d_h (M,) (or (1,) when h is a
    scalar), d_W (K, M), d_b (K,).  Only the free centres are parameters
    (code/spinn2d.py:109-113)."""
    kind = kind_id(kind)
    x, y, xc, yc, W, g = _prep(dtype, x, y, xc, yc, W, gjets)
    M = xc.shape[0]
    W = W.reshape(-1, M)
    K = W.shape[0]
    hs = np.asarray(h).reshape(-1).size == 1
    r = 1.0 / _widths(h, M, dtype)
    N = x.shape[0]
    g = g.reshape(g.shape[0], N, K)
    if g.shape[0] == 5:
        g = np.concatenate([g, np.zeros((1, N, K), dtype=dtype)], axis=0)
    dW = np.zeros((K, M), dtype=dtype)
    dc = np.zeros(M, dtype=dtype)
    dd = np.zeros(M, dtype=dtype)
    dh = np.zeros(M, dtype=dtype)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        X = x[s:e, None] - xc[None, :]
        Y = y[s:e, None] - yc[None, :]
        p = _phi_derivs_2d(kind, X, Y, r)
        gs = g[:, s:e]                                   # (6, n, K)
        for a, key in enumerate(JETS_2D):
            dW += gs[a].T @ p[key]
        G = [gs[a] @ W for a in range(6)]                # each (n, M)
        dcij = -(G[0] * p["x"] + G[1] * p["xx"] + G[2] * p["xy"]
                 + G[3] * p["xxx"] + G[4] * p["xyy"] + G[5] * p["xxy"])
        ddij = -(G[0] * p["y"] + G[1] * p["xy"] + G[2] * p["yy"]
                 + G[3] * p["xxy"] + G[4] * p["yyy"] + G[5] * p["xyy"])
        order = (G[1] * p["x"] + G[2] * p["y"]
                 + 2 * (G[3] * p["xx"] + G[4] * p["yy"] + G[5] * p["xy"]))
        dc += dcij.sum(0)
        dd += ddij.sum(0)
        dh += (r[None, :] * (X * dcij + Y * ddij - order)).sum(0)
    n_free = M if n_free is None else n_free
    return {
        "d_xc": dc[:n_free], "d_yc": dd[:n_free],
        "d_h": dh.sum(keepdims=True) if hs else dh,
        "d_W": dW, "d_b": g[0].sum(0),
    }


def grads1d(kind, x, xc, h, W, gjets, n_free=None, dtype=np.float64, chunk=4096):
    """1-D twin: gjets (3, N, K) for u, u_x, u_xx."""
    kind = kind_id(kind)
    x, xc, W, g = _prep(dtype, x, xc, W, gjets)
    M = xc.shape[0]
    W = W.reshape(-1, M)
    K = W.shape[0]
    hs = np.asarray(h).reshape(-1).size == 1
    r = 1.0 / _widths(h, M, dtype)
    N = x.shape[0]
    g = g.reshape(3, N, K)
    dW = np.zeros((K, M), dtype=dtype)
    dc = np.zeros(M, dtype=dtype)
    dh = np.zeros(M, dtype=dtype)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        X = x[s:e, None] - xc[None, :]
        p = _phi_derivs_1d(kind, X, r)
        gs = g[:, s:e]
        for a, key in enumerate(JETS_1D):
            dW += gs[a].T @ p[key]
        G = [gs[a] @ W for a in range(3)]
        dcij = -(G[0] * p["x"] + G[1] * p["xx"] + G[2] * p["xxx"])
        order = G[1] * p["x"] + 2 * G[2] * p["xx"]
        dc += dcij.sum(0)
        dh += (r[None, :] * (X * dcij - order)).sum(0)
    n_free = M if n_free is None else n_free
    return {"d_xc": dc[:n_free], "d_h": dh.sum(keepdims=True) if hs else dh,
            "d_W": dW, "d_b": g[0].sum(0)}
