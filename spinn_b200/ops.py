"""PyTorch custom ops over the C ABI: fused SPINN jets forward / backward.

``fused_jets`` replaces, for one call of the network, what the reference does with
a dense forward (code/spinn2d.py:169-179, code/spinn1d.py:147-156) followed by
three nested ``torch.autograd.grad`` calls in the problem class
(code/pde2d_base.py:46-62, code/ode_base.py:32-42): it returns u and its analytic
spatial derivatives, and its autograd backward returns the gradients w.r.t.
centres, widths, weights and bias that ``loss.backward()`` (code/common.py:242-248)
would produce.

``attach_tower`` makes those precomputed jets visible to UNCHANGED problem code:
the problem class still calls ``ag.grad(u, (xs, ys), ..., create_graph=True)``
twice; the tower's custom backward hands it the precomputed ``u_x``/``u_xx``...
instead of re-differentiating an (N, M) graph.

There is no CPU path: inputs must be CUDA fp32 tensors and the native library
must be loadable, otherwise an exception is raised.
"""
from __future__ import annotations

import torch

from . import _cabi

GAUSSIAN = 0
SOFTPLUS = 1

_NJETS = {(2, 0): 1, (2, 1): 3, (2, 2): 5, (2, 3): 6, (1, 0): 1, (1, 1): 2, (1, 2): 3}
MAX_K = 4


def num_jets(dim, level):
    return _NJETS[(dim, level)]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CudaBackend:
    """Thin marshalling layer: torch tensors -> raw device pointers -> C ABI."""

    name = "cuda"

    @staticmethod
    def _check(*tensors):
        dev = None
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda:
                raise _cabi.SpinnNativeError(
                    "spinn_b200 fused ops need CUDA tensors (no CPU fallback); got device "
                    f"{t.device}. Run with --gpu / common.device('cuda').")
            if t.dtype != torch.float32:
                raise TypeError(f"spinn_b200 fused ops are fp32 (reference common.py:28); got {t.dtype}")
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise ValueError("all tensors must live on the same CUDA device")
        return dev

    def forward(self, dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, b):
        lib = _cabi.load()
        dev = self._check(x, y, xf, yf, xfix, yfix, h, W, b)
        N = x.numel()
        Mf, Mx = xf.numel(), xfix.numel()
        K = W.shape[0]
        nj = num_jets(dim, level)
        with torch.cuda.device(dev):
            jets = torch.empty((nj, N, K), dtype=torch.float32, device=dev)
            if N == 0:
                return jets
            if dim == 2:
                nbytes = lib.spinn2d_fwd_workspace_bytes(N, Mf + Mx, K)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                rc = lib.spinn2d_jets_fwd(kind, level, N, Mf, Mx, K, _ptr(x), _ptr(y), _ptr(xf), _ptr(yf),
                                          _ptr(xfix), _ptr(yfix), _ptr(h), int(h.numel() == 1), _ptr(W),
                                          _ptr(b), _ptr(jets), _ptr(ws), nbytes, _stream())
            else:
                nbytes = lib.spinn1d_fwd_workspace_bytes(N, Mf + Mx, K)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                rc = lib.spinn1d_jets_fwd(kind, level, N, Mf, Mx, K, _ptr(x), _ptr(xf), _ptr(xfix), _ptr(h),
                                          int(h.numel() == 1), _ptr(W), _ptr(b), _ptr(jets), _ptr(ws),
                                          nbytes, _stream())
        _cabi.check(rc, f"spinn{dim}d_jets_fwd")
        return jets

    def backward(self, dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, gj, want_bias, mask=0):
        """mask: bit a set <=> gj[a] holds an upstream gradient (0 = all rows of the level)."""
        lib = _cabi.load()
        dev = self._check(x, y, xf, yf, xfix, yfix, h, W, gj)
        N = x.numel()
        Mf, Mx = xf.numel(), xfix.numel()
        M = Mf + Mx
        K = W.shape[0]
        with torch.cuda.device(dev):
            d_xc = torch.empty(Mf, dtype=torch.float32, device=dev)
            d_yc = torch.empty(Mf, dtype=torch.float32, device=dev) if dim == 2 else None
            d_h = torch.empty(h.numel(), dtype=torch.float32, device=dev)
            d_W = torch.empty((K, M), dtype=torch.float32, device=dev)
            d_b = torch.empty(K, dtype=torch.float32, device=dev) if want_bias else None
            if dim == 2:
                nbytes = lib.spinn2d_bwd_workspace_bytes(N, M, K)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                rc = lib.spinn2d_jets_bwd(kind, level, mask, N, Mf, Mx, K, _ptr(x), _ptr(y), _ptr(xf), _ptr(yf),
                                          _ptr(xfix), _ptr(yfix), _ptr(h), int(h.numel() == 1), _ptr(W),
                                          _ptr(gj), _ptr(d_xc), _ptr(d_yc), _ptr(d_h), _ptr(d_W), _ptr(d_b),
                                          _ptr(ws), nbytes, _stream())
            else:
                nbytes = lib.spinn1d_bwd_workspace_bytes(N, M, K)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                rc = lib.spinn1d_jets_bwd(kind, level, mask, N, Mf, Mx, K, _ptr(x), _ptr(xf), _ptr(xfix), _ptr(h),
                                          int(h.numel() == 1), _ptr(W), _ptr(gj), _ptr(d_xc), _ptr(d_h),
                                          _ptr(d_W), _ptr(d_b), _ptr(ws), nbytes, _stream())
        _cabi.check(rc, f"spinn{dim}d_jets_bwd")
        return d_xc, d_yc, d_h, d_W, d_b


    # ---- compact-support variants (2-D Gaussian, K=1); see include/spinn_b200.h -------------
    _ws_cache = {}

    def _sparse_ws(self, nbytes, dev):
        """Backward workspace of the compact-support path, cached per (size, device, stream): it holds
        per-chunk partial sums (MBs at M=4096) and used to be allocated on every backward call."""
        if torch.cuda.is_current_stream_capturing():       # graph capture: memory must come from the graph's pool
            return torch.empty(nbytes, dtype=torch.uint8, device=dev)
        key = (nbytes, dev, torch.cuda.current_stream(dev).cuda_stream)
        ws = self._ws_cache.get(key)
        if ws is None:
            if len(self._ws_cache) > 8:
                self._ws_cache.clear()
            ws = self._ws_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return ws

    def forward_sparse(self, level, x, y, xf, yf, xfix, yfix, h, W, b, perm, cutoff, stats=None):
        lib = _cabi.load()
        dev = self._check(x, y, xf, yf, xfix, yfix, h, W, b)
        N, Mf, Mx = x.numel(), xf.numel(), xfix.numel()
        nj = num_jets(2, level)
        with torch.cuda.device(dev):
            jets = torch.empty((nj, N, 1), dtype=torch.float32, device=dev)
            if N == 0:
                return jets
            nbytes = lib.spinn2d_sparse_fwd_workspace_bytes(N, Mf + Mx)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rc = lib.spinn2d_jets_fwd_sparse(GAUSSIAN, level, N, Mf, Mx, 1, _ptr(x), _ptr(y), _ptr(xf), _ptr(yf),
                                             _ptr(xfix), _ptr(yfix), _ptr(h), int(h.numel() == 1), _ptr(W),
                                             _ptr(b), _ptr(perm), float(cutoff), _ptr(jets), _ptr(stats),
                                             _ptr(ws), nbytes, _stream())
        _cabi.check(rc, "spinn2d_jets_fwd_sparse")
        return jets

    def backward_sparse(self, level, x, y, xf, yf, xfix, yfix, h, W, gj, want_bias, mask, perm, cutoff,
                        stats=None):
        lib = _cabi.load()
        dev = self._check(x, y, xf, yf, xfix, yfix, h, W, gj)
        N, Mf, Mx = x.numel(), xf.numel(), xfix.numel()
        M = Mf + Mx
        with torch.cuda.device(dev):
            d_xc = torch.empty(Mf, dtype=torch.float32, device=dev)
            d_yc = torch.empty(Mf, dtype=torch.float32, device=dev)
            d_h = torch.empty(h.numel(), dtype=torch.float32, device=dev)
            d_W = torch.empty((1, M), dtype=torch.float32, device=dev)
            d_b = torch.empty(1, dtype=torch.float32, device=dev) if want_bias else None
            nbytes = lib.spinn2d_sparse_bwd_workspace_bytes(N, M)
            ws = self._sparse_ws(nbytes, dev)
            rc = lib.spinn2d_jets_bwd_sparse(GAUSSIAN, level, mask, N, Mf, Mx, 1, _ptr(x), _ptr(y), _ptr(xf),
                                             _ptr(yf), _ptr(xfix), _ptr(yfix), _ptr(h), int(h.numel() == 1),
                                             _ptr(W), _ptr(gj), _ptr(perm), float(cutoff), _ptr(d_xc),
                                             _ptr(d_yc), _ptr(d_h), _ptr(d_W), _ptr(d_b), _ptr(stats),
                                             _ptr(ws), nbytes, _stream())
        _cabi.check(rc, "spinn2d_jets_bwd_sparse")
        return d_xc, d_yc, d_h, d_W, d_b


_BACKEND = CudaBackend()


def set_backend(backend):
    """Test hook: tests/ inject an oracle-backed stand-in to exercise the host logic
    (tower, plug-in surface) in the GPU-less build container.  Product code never
    calls this."""
    global _BACKEND
    prev = _BACKEND
    _BACKEND = backend
    return prev


def _level_for(dim, last_index):
    """Smallest level whose upstream-gradient slots cover jet index `last_index`."""
    for level in range(0, 4 if dim == 2 else 3):
        if num_jets(dim, level) > last_index:
            return level
    raise ValueError(last_index)


class _FusedJets(torch.autograd.Function):
    """jets = F(x, y; centres, widths, weights).  Outputs: one (N, K) tensor per jet.

    x, y are treated as constants here (callers pass detached tensors): the
    spatial derivatives ARE the outputs; the tower below wires them to autograd."""

    @staticmethod
    def forward(ctx, dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, b, cutoff=None, perm=None):
        hk = h.reshape(-1)
        sparse = bool(cutoff) and sparse_supported(dim, kind, level, W.shape[0])
        if sparse:
            jets = _BACKEND.forward_sparse(level, x, y, xf, yf, xfix, yfix, hk, W, b, perm, cutoff)
        else:
            jets = _BACKEND.forward(dim, kind, level, x, y, xf, yf, xfix, yfix, hk, W, b)
        ctx.sparse = (float(cutoff), perm) if sparse else None
        ctx.meta = (dim, kind, level)
        ctx.h_shape = h.shape
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, y, xf, yf, xfix, yfix, h, W)
        ctx.set_materialize_grads(False)
        return tuple(jets.unbind(0))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        dim, kind, level = ctx.meta
        x, y, xf, yf, xfix, yfix, h, W = ctx.saved_tensors
        present = [i for i, g in enumerate(gs) if g is not None]
        if not present:
            return (None,) * 14
        lv = _level_for(dim, present[-1])
        ng = num_jets(dim, lv)
        N, K = x.numel(), W.shape[0]
        # rows of absent jets are never read by the kernels (present-mask): no zero fill
        gj = torch.empty((ng, N, K), dtype=torch.float32, device=x.device)
        mask = 0
        i = 0
        while i < len(present):                  # one copy kernel per RUN of adjacent present rows
            j = i
            while j + 1 < len(present) and present[j + 1] == present[j] + 1:
                j += 1
            rows = [gs[r].reshape(N, K) for r in present[i:j + 1]]
            if len(rows) == 1:
                gj[present[i]].copy_(rows[0])
            else:
                torch.stack(rows, 0, out=gj[present[i]:present[j] + 1])
            for r in present[i:j + 1]:
                mask |= 1 << r
            i = j + 1
        if ctx.sparse is not None and lv <= 2:
            cutoff, perm = ctx.sparse
            d_xc, d_yc, d_h, d_W, d_b = _BACKEND.backward_sparse(
                lv, x, y, xf, yf, xfix, yfix, h.reshape(-1), W, gj, ctx.has_bias, mask, perm, cutoff)
        else:
            d_xc, d_yc, d_h, d_W, d_b = _BACKEND.backward(
                dim, kind, lv, x, y, xf, yf, xfix, yfix, h.reshape(-1), W, gj, ctx.has_bias, mask)
        return (None, None, None, None, None, d_xc, d_yc, None, None,
                d_h.reshape(ctx.h_shape), d_W, d_b, None, None)


def fused_jets(dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, b, cutoff=None, perm=None):
    """Tuple of jets, each (N, K).  2-D order u,ux,uy,uxx,uyy[,uxy]; 1-D u,ux,uxx.
    cutoff (in node widths) selects the compact-support kernels where they exist (2-D Gaussian,
    K=1, levels 2-3); perm is the int32 node visiting order for them (see morton_order)."""
    return _FusedJets.apply(dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, b, cutoff, perm)


def sparse_supported(dim, kind, level, K):
    return dim == 2 and kind == GAUSSIAN and K == 1 and level in (2, 3)


def _spread_bits16(v):
    """0000abcd -> 0a0b0c0d for 16-bit integers held in int64 tensors."""
    v = (v | (v << 8)) & 0x00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def morton_order(xc, yc):
    """int32 permutation that visits the nodes along a Z-order curve of their centres, so that
    32 consecutive nodes form a compact block (no host synchronisation)."""
    x = xc.detach().float()
    y = yc.detach().float()
    lo_x, lo_y = x.min(), y.min()
    sx = 65535.0 / (x.max() - lo_x).clamp_min(1e-30)
    sy = 65535.0 / (y.max() - lo_y).clamp_min(1e-30)
    qx = ((x - lo_x) * sx).long().clamp_(0, 65535)
    qy = ((y - lo_y) * sy).long().clamp_(0, 65535)
    code = _spread_bits16(qx) | (_spread_bits16(qy) << 1)
    return torch.argsort(code).to(torch.int32)


_ZEROS = {}


def _zero_scalar(like):
    key = (like.device, like.dtype)
    z = _ZEROS.get(key)
    if z is None:
        z = _ZEROS[key] = torch.zeros((), device=like.device, dtype=like.dtype)
    return z


class _Attach(torch.autograd.Function):
    """out = v, but with DECLARED derivatives d out/dx = vx, d out/dy = vy.

    backward(g) = (g -> v, sum_k g*vx -> x, sum_k g*vy -> y); it is written with
    differentiable torch ops on saved inputs, so ``create_graph=True`` keeps the
    dependence on vx, vy (needed for the second ``ag.grad`` and for
    ``loss.backward()`` to reach the fused op)."""

    @staticmethod
    def forward(ctx, v, x, y, vx, vy):
        ctx.save_for_backward(vx, vy)
        ctx.has_y = y is not None
        return v.view_as(v)

    @staticmethod
    def backward(ctx, g):
        vx, vy = ctx.saved_tensors

        def along(v):                # sum_k g * v; a derivative declared "not provided" reads as zero
            if v is None:            # (a cached scalar zero, broadcast: no kernel at all)
                return _zero_scalar(g).expand(g.shape[0])
            t = g * v
            if t.dim() == 1:
                return t
            return t.squeeze(-1) if t.shape[-1] == 1 else t.sum(-1)      # K = 1: a view, no reduction

        gx = along(vx)
        gy = along(vy) if ctx.has_y else None
        return g, gx, gy, None, None


def attach_tower_2d(x, y, u, ux, uy, uxx, uyy, uxy):
    """(N,K) jets -> u' such that the reference's ``_compute_derivatives``
    (code/pde2d_base.py:46-62) recovers ux, uy, uxx, uyy (and uxy) from it.
    ``uxy=None``: the mixed derivative was not evaluated; the slots the recipe computes and drops
    (d ux/dy, d uy/dx) read as zero."""
    ux_ = _Attach.apply(ux, x, y, uxx, uxy)
    uy_ = _Attach.apply(uy, x, y, uxy, uyy)
    return _Attach.apply(u, x, y, ux_, uy_)


def attach_tower_1d(x, u, ux, uxx):
    """1-D twin for ``BasicODE._compute_derivatives`` (code/ode_base.py:32-42)."""
    ux_ = _Attach1.apply(ux, x, uxx)
    return _Attach1.apply(u, x, ux_)


class _Attach1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, x, vx):
        ctx.save_for_backward(vx)
        return v.view_as(v)

    @staticmethod
    def backward(ctx, g):
        (vx,) = ctx.saved_tensors
        gx = g * vx
        if gx.dim() > 1:
            gx = gx.squeeze(-1) if gx.shape[-1] == 1 else gx.sum(-1)
        return g, gx, None


# --------------------------------------------------------------------------------------
# partition of unity: jets of u = T/S from the jets of T = sum W phi and S = sum phi
# --------------------------------------------------------------------------------------
def quotient_jets_2d(T, S):
    """T: list of (N, K) jets of the weighted sum, S: list of (N, 1) jets of the plain sum, in
    the order u, ux, uy[, uxx, uyy[, uxy]].  Returns the jets of u = T/S (quotient rule)."""
    inv = 1.0 / S[0]
    u = T[0] * inv
    out = [u]
    if len(T) >= 3:
        ux = (T[1] - u * S[1]) * inv
        uy = (T[2] - u * S[2]) * inv
        out += [ux, uy]
    if len(T) >= 5:
        out.append((T[3] - 2.0 * ux * S[1] - u * S[3]) * inv)
        out.append((T[4] - 2.0 * uy * S[2] - u * S[4]) * inv)
    if len(T) >= 6:
        out.append((T[5] - ux * S[2] - uy * S[1] - u * S[5]) * inv)
    return tuple(out)


def quotient_jets_1d(T, S):
    inv = 1.0 / S[0]
    u = T[0] * inv
    out = [u]
    if len(T) >= 2:
        ux = (T[1] - u * S[1]) * inv
        out.append(ux)
    if len(T) >= 3:
        out.append((T[2] - 2.0 * ux * S[1] - u * S[2]) * inv)
    return tuple(out)
