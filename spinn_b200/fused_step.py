"""Fused interior step for the stock Poisson residual: forward jets -> residual/loss/
upstream-gradient epilogue -> backward -> (all-reduce), all through the C ABI with
preallocated buffers.  This is the throughput path the benchmark times; it computes
exactly what ``RegularPDE.interior_loss`` + ``loss.backward()`` compute for
``pde = scale*(u_xx + u_yy + amp*sin(kx x) sin(ky y) + f0)`` (reference
code/pde2d_base.py:132-141, code/poisson2d_sine.py:16-18, code/poisson2d_irreg_dom.py:146-147).
"""
from __future__ import annotations

import math

import torch

from . import _cabi
from .distributed import GradPack
from .ops import _ptr, _stream

SINE = dict(scale=0.1, amp=20.0 * math.pi ** 2, kx=2.0 * math.pi, ky=4.0 * math.pi, f0=0.0)
UNIT_SOURCE = dict(scale=1.0, amp=0.0, kx=0.0, ky=0.0, f0=1.0)


class PoissonInteriorStep:
    #: kernels of ours enqueued by one run() through spinn2d_poisson_interior_step: pack_nodes_step,
    #: forward (with the residual epilogue: loss partials + the backward's point records), backward,
    #: second stage.
    #: structured=False / cutoff take the separate calls: pack_nodes(fwd), fwd, residual,
    #: add_partials, pack_nodes(bwd), pack_points, bwd, bias_partial, finalize
    LAUNCHES = 4
    LAUNCHES_SEPARATE = 9

    def __init__(self, nn, x, y, n_total=None, residual=SINE, structured=True, cutoff=None):
        if nn.layer2.weight.shape[0] != 1:
            raise ValueError("PoissonInteriorStep is for scalar problems (K=1)")
        self.lib = _cabi.load()
        self.nn = nn
        self.x = x.detach().contiguous()
        self.y = y.detach().contiguous()
        self.N = self.x.numel()
        self.n_total = float(self.N if n_total is None else n_total)
        self.res = dict(residual)
        # dLoss/dJets of this residual is structurally zero for u, u_x, u_y: tell the backward
        # (0x18 = rows u_xx, u_yy present); structured=False runs the general 5-row backward
        self.upstream_mask = 0x18 if structured else 0
        l1, l2 = nn.layer1, nn.layer2
        dev = self.x.device
        self.Mf, self.Mx = l1.x.numel(), l1.xf.numel()
        M = self.Mf + self.Mx
        self.pack = GradPack(2, self.Mf, M, 1, l1.h.numel(), l2.bias is not None, dev)
        self.jets = torch.empty((5, self.N), dtype=torch.float32, device=dev)
        self.gjets = torch.empty((5, self.N), dtype=torch.float32, device=dev)
        self.ws_f = torch.empty(self.lib.spinn2d_fwd_workspace_bytes(self.N, M, 1), dtype=torch.uint8, device=dev)
        self.ws_b = torch.empty(self.lib.spinn2d_bwd_workspace_bytes(self.N, M, 1), dtype=torch.uint8, device=dev)
        self.ws_r = torch.empty(self.lib.spinn2d_poisson_residual_workspace_bytes(self.N), dtype=torch.uint8, device=dev)
        # compact-support mode (opt-in): Morton order of the nodes, pair counter for reporting
        self.cutoff = float(cutoff) if cutoff else None
        if self.cutoff:
            from .ops import morton_order
            self.perm = morton_order(*l1.centers())
            self.stats = torch.zeros(2, dtype=torch.int64, device=dev)     # [fwd pairs, bwd pairs]
            self.ws_sf = torch.empty(self.lib.spinn2d_sparse_fwd_workspace_bytes(self.N, M), dtype=torch.uint8, device=dev)
            self.ws_sb = torch.empty(self.lib.spinn2d_sparse_bwd_workspace_bytes(self.N, M), dtype=torch.uint8, device=dev)

    def run(self, all_reduce=True, separate=False):
        """Enqueue one interior fwd+bwd on the current stream.  Afterwards
        ``self.pack.flat`` holds [grads | loss] (summed over ranks if all_reduce).
        separate=True issues the three C-ABI calls (forward, residual, backward) one by one and
        leaves dLoss/dJets in ``self.gjets``; the default is the single fused entry."""
        lib, nn = self.lib, self.nn
        l1, l2 = nn.layer1, nn.layer2
        s = _stream()
        hs = int(l1.h.numel() == 1)
        p = self.pack
        if self.cutoff:
            p.view("loss").zero_()
            return self._run_sparse(all_reduce)
        if self.upstream_mask == 0x18 and not separate:
            r = self.res
            rc = lib.spinn2d_poisson_interior_step(
                nn.kind, self.N, self.n_total, self.Mf, self.Mx, _ptr(self.x), _ptr(self.y), _ptr(l1.x), _ptr(l1.y),
                _ptr(l1.xf), _ptr(l1.yf), _ptr(l1.h), hs, _ptr(l2.weight), _ptr(l2.bias), r["scale"], r["amp"],
                r["kx"], r["ky"], r["f0"], _ptr(self.jets), _ptr(p.view("loss")), _ptr(p.view("d_x")),
                _ptr(p.view("d_y")), _ptr(p.view("d_h")), _ptr(p.view("d_W")),
                _ptr(p.view("d_b")) if p.has("d_b") else None, _ptr(self.ws_f), self.ws_f.numel(),
                _ptr(self.ws_b), self.ws_b.numel(), s)
            _cabi.check(rc, "spinn2d_poisson_interior_step")
            if all_reduce:
                p.all_reduce()
            return p.flat
        p.view("loss").zero_()
        rc = lib.spinn2d_jets_fwd(nn.kind, 2, self.N, self.Mf, self.Mx, 1, _ptr(self.x), _ptr(self.y),
                                  _ptr(l1.x), _ptr(l1.y), _ptr(l1.xf), _ptr(l1.yf), _ptr(l1.h), hs,
                                  _ptr(l2.weight), _ptr(l2.bias), _ptr(self.jets), _ptr(self.ws_f),
                                  self.ws_f.numel(), s)
        _cabi.check(rc, "spinn2d_jets_fwd")
        r = self.res
        rc = lib.spinn2d_poisson_residual(self.N, self.n_total, _ptr(self.x), _ptr(self.y), _ptr(self.jets),
                                          r["scale"], r["amp"], r["kx"], r["ky"], r["f0"], _ptr(self.gjets),
                                          _ptr(p.view("loss")), _ptr(self.ws_r), self.ws_r.numel(), s)
        _cabi.check(rc, "spinn2d_poisson_residual")
        rc = lib.spinn2d_jets_bwd(nn.kind, 2, self.upstream_mask, self.N, self.Mf, self.Mx, 1, _ptr(self.x), _ptr(self.y),
                                  _ptr(l1.x), _ptr(l1.y), _ptr(l1.xf), _ptr(l1.yf), _ptr(l1.h), hs,
                                  _ptr(l2.weight), _ptr(self.gjets), _ptr(p.view("d_x")), _ptr(p.view("d_y")),
                                  _ptr(p.view("d_h")), _ptr(p.view("d_W")),
                                  _ptr(p.view("d_b")) if p.has("d_b") else None,
                                  _ptr(self.ws_b), self.ws_b.numel(), s)
        _cabi.check(rc, "spinn2d_jets_bwd")
        if all_reduce:
            p.all_reduce()
        return p.flat

    def upstream_from_jets(self):
        """dLoss/dJets [5][N] recomputed from ``self.jets`` with plain torch ops (for checks: the fused
        entry never materialises it)."""
        r = self.res
        f = r["amp"] * torch.sin(r["kx"] * self.x) * torch.sin(r["ky"] * self.y) + r["f0"]
        res = r["scale"] * (self.jets[3] + self.jets[4] + f)
        g = torch.zeros_like(self.jets)
        g[3] = g[4] = (2.0 * r["scale"] / self.n_total) * res
        return g

    def _run_sparse(self, all_reduce):
        lib, nn, p, s = self.lib, self.nn, self.pack, _stream()
        l1, l2 = nn.layer1, nn.layer2
        hs = int(l1.h.numel() == 1)
        st = self.stats
        rc = lib.spinn2d_jets_fwd_sparse(nn.kind, 2, self.N, self.Mf, self.Mx, 1, _ptr(self.x), _ptr(self.y),
                                         _ptr(l1.x), _ptr(l1.y), _ptr(l1.xf), _ptr(l1.yf), _ptr(l1.h), hs,
                                         _ptr(l2.weight), _ptr(l2.bias), _ptr(self.perm), self.cutoff,
                                         _ptr(self.jets), st[0:1].data_ptr(), _ptr(self.ws_sf),
                                         self.ws_sf.numel(), s)
        _cabi.check(rc, "spinn2d_jets_fwd_sparse")
        r = self.res
        rc = lib.spinn2d_poisson_residual(self.N, self.n_total, _ptr(self.x), _ptr(self.y), _ptr(self.jets),
                                          r["scale"], r["amp"], r["kx"], r["ky"], r["f0"], _ptr(self.gjets),
                                          _ptr(p.view("loss")), _ptr(self.ws_r), self.ws_r.numel(), s)
        _cabi.check(rc, "spinn2d_poisson_residual")
        rc = lib.spinn2d_jets_bwd_sparse(nn.kind, 2, self.upstream_mask, self.N, self.Mf, self.Mx, 1,
                                         _ptr(self.x), _ptr(self.y), _ptr(l1.x), _ptr(l1.y), _ptr(l1.xf),
                                         _ptr(l1.yf), _ptr(l1.h), hs, _ptr(l2.weight), _ptr(self.gjets),
                                         _ptr(self.perm), self.cutoff, _ptr(p.view("d_x")), _ptr(p.view("d_y")),
                                         _ptr(p.view("d_h")), _ptr(p.view("d_W")),
                                         _ptr(p.view("d_b")) if p.has("d_b") else None, st[1:2].data_ptr(),
                                         _ptr(self.ws_sb), self.ws_sb.numel(), s)
        _cabi.check(rc, "spinn2d_jets_bwd_sparse")
        if all_reduce:
            p.all_reduce()
        return p.flat

    def assign_grads(self):
        """Point the module's .grad at the (all-reduced) packed gradients."""
        l1, l2, p = self.nn.layer1, self.nn.layer2, self.pack
        l1.x.grad = p.view("d_x")
        l1.y.grad = p.view("d_y")
        l1.h.grad = p.view("d_h").reshape(l1.h.shape)
        l2.weight.grad = p.view("d_W").reshape(l2.weight.shape)
        if l2.bias is not None:
            l2.bias.grad = p.view("d_b")
        return p.view("loss")


class _ResidualStep:
    """Shared plumbing of the device-resident interior steps below: forward jets -> residual
    epilogue (loss + dLoss/dJets) -> fused backward into a GradPack, all through the C ABI with
    preallocated buffers."""

    LAUNCHES = 8        # pack_nodes, fwd, residual, add_partials, pack_nodes, bwd, (bias_partial,) finalize

    def _setup(self, nn, dim, n_points, n_jets, K):
        self.lib = _cabi.load()
        self.nn = nn
        l1, l2 = nn.layer1, nn.layer2
        free = l1.x if dim == 2 else l1.center
        fixed = l1.xf if dim == 2 else l1.fixed
        self.Mf, self.Mx = free.numel(), fixed.numel()
        M = self.Mf + self.Mx
        dev = free.device
        self.N, self.K = n_points, K
        self.pack = GradPack(dim, self.Mf, M, K, l1.h.numel(), l2.bias is not None, dev)
        self.jets = torch.empty((n_jets, n_points, K), dtype=torch.float32, device=dev)
        self.gjets = torch.zeros((n_jets, n_points, K), dtype=torch.float32, device=dev)
        fws = self.lib.spinn2d_fwd_workspace_bytes if dim == 2 else self.lib.spinn1d_fwd_workspace_bytes
        bws = self.lib.spinn2d_bwd_workspace_bytes if dim == 2 else self.lib.spinn1d_bwd_workspace_bytes
        self.ws_f = torch.empty(fws(n_points, M, K), dtype=torch.uint8, device=dev)
        self.ws_b = torch.empty(bws(n_points, M, K), dtype=torch.uint8, device=dev)
        self.ws_r = torch.empty(self.lib.spinn2d_poisson_residual_workspace_bytes(n_points), dtype=torch.uint8,
                                device=dev)

    def assign_grads(self):
        l1, l2, p = self.nn.layer1, self.nn.layer2, self.pack
        if p.dim == 2:
            l1.x.grad, l1.y.grad = p.view("d_x"), p.view("d_y")
        else:
            l1.center.grad = p.view("d_x")
        l1.h.grad = p.view("d_h").reshape(l1.h.shape)
        l2.weight.grad = p.view("d_W").reshape(l2.weight.shape)
        if l2.bias is not None:
            l2.bias.grad = p.view("d_b")
        return p.view("loss")


class Ode3InteriorStep(_ResidualStep):
    """BASELINE config 1 (code/ode3.py:12-16, code/ode_base.py:83-92): 1-D jets (u, u_x, u_xx) ->
    res = u_xx + f(x), loss = sum(res^2)/n_total -> backward with only the u_xx row present."""

    def __init__(self, nn, x, n_total=None, Kc=0.01):
        self.x = x.detach().contiguous()
        self._setup(nn, 1, self.x.numel(), 3, 1)
        self.n_total = float(self.N if n_total is None else n_total)
        self.Kc = float(Kc)

    def run(self, all_reduce=True):
        lib, nn, p, s = self.lib, self.nn, self.pack, _stream()
        l1, l2 = nn.layer1, nn.layer2
        hs = int(l1.h.numel() == 1)
        p.view("loss").zero_()
        rc = lib.spinn1d_jets_fwd(nn.kind, 2, self.N, self.Mf, self.Mx, 1, _ptr(self.x), _ptr(l1.center),
                                  _ptr(l1.fixed), _ptr(l1.h), hs, _ptr(l2.weight), _ptr(l2.bias), _ptr(self.jets),
                                  _ptr(self.ws_f), self.ws_f.numel(), s)
        _cabi.check(rc, "spinn1d_jets_fwd")
        rc = lib.spinn1d_ode3_residual(self.N, self.n_total, _ptr(self.x), _ptr(self.jets), self.Kc,
                                       _ptr(self.gjets), _ptr(p.view("loss")), _ptr(self.ws_r), self.ws_r.numel(), s)
        _cabi.check(rc, "spinn1d_ode3_residual")
        rc = lib.spinn1d_jets_bwd(nn.kind, 2, 0x4, self.N, self.Mf, self.Mx, 1, _ptr(self.x), _ptr(l1.center),
                                  _ptr(l1.fixed), _ptr(l1.h), hs, _ptr(l2.weight), _ptr(self.gjets),
                                  _ptr(p.view("d_x")), _ptr(p.view("d_h")), _ptr(p.view("d_W")),
                                  _ptr(p.view("d_b")) if p.has("d_b") else None, _ptr(self.ws_b),
                                  self.ws_b.numel(), s)
        _cabi.check(rc, "spinn1d_jets_bwd")
        if all_reduce:
            p.all_reduce()
        return p.flat


class CavityInteriorStep(_ResidualStep):
    """BASELINE config 4 (code/cavity.py:84-100): 2-D jets of the three outputs (u, v, p) ->
    continuity + momentum residuals, summed over the points -> backward with all 5 rows."""

    def __init__(self, nn, x, y, nu=0.01):
        if nn.layer2.weight.shape[0] != 3:
            raise ValueError("CavityInteriorStep needs the 3-output network (u, v, p)")
        self.x, self.y = x.detach().contiguous(), y.detach().contiguous()
        self._setup(nn, 2, self.x.numel(), 5, 3)
        self.nu = float(nu)

    def run(self, all_reduce=True):
        lib, nn, p, s = self.lib, self.nn, self.pack, _stream()
        l1, l2 = nn.layer1, nn.layer2
        hs = int(l1.h.numel() == 1)
        p.view("loss").zero_()
        rc = lib.spinn2d_jets_fwd(nn.kind, 2, self.N, self.Mf, self.Mx, 3, _ptr(self.x), _ptr(self.y), _ptr(l1.x),
                                  _ptr(l1.y), _ptr(l1.xf), _ptr(l1.yf), _ptr(l1.h), hs, _ptr(l2.weight),
                                  _ptr(l2.bias), _ptr(self.jets), _ptr(self.ws_f), self.ws_f.numel(), s)
        _cabi.check(rc, "spinn2d_jets_fwd")
        rc = lib.spinn2d_cavity_residual(self.N, self.nu, _ptr(self.jets), _ptr(self.gjets), _ptr(p.view("loss")),
                                         _ptr(self.ws_r), self.ws_r.numel(), s)
        _cabi.check(rc, "spinn2d_cavity_residual")
        rc = lib.spinn2d_jets_bwd(nn.kind, 2, 0, self.N, self.Mf, self.Mx, 3, _ptr(self.x), _ptr(self.y), _ptr(l1.x),
                                  _ptr(l1.y), _ptr(l1.xf), _ptr(l1.yf), _ptr(l1.h), hs, _ptr(l2.weight),
                                  _ptr(self.gjets), _ptr(p.view("d_x")), _ptr(p.view("d_y")), _ptr(p.view("d_h")),
                                  _ptr(p.view("d_W")), _ptr(p.view("d_b")) if p.has("d_b") else None,
                                  _ptr(self.ws_b), self.ws_b.numel(), s)
        _cabi.check(rc, "spinn2d_jets_bwd")
        if all_reduce:
            p.all_reduce()
        return p.flat
