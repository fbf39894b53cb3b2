"""Shared test helpers: golden loading and the parity metric."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

OP_CASES_2D = sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "op_*2d*.npz")))
OP_CASES_1D = sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "op_*1d*.npz")))


def load_case(name):
    d = np.load(os.path.join(GOLDEN, f"op_{name}.npz"))
    return {k: (d[k].item() if d[k].ndim == 0 else d[k]) for k in d.files}


def mixed_err(new, ref):
    """max over elements of |new-ref| / (|ref| + floor) with floor = 1e-2*max|ref|:
    the 'relative per element, with an absolute floor tied to the field's scale'
    metric of SURVEY.md 7.3 (strict per-element relative error is meaningless at
    the zero crossings of u_xx, dW, ... where fp32 itself is off by 1e-4..1e-2)."""
    new = np.asarray(new, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref)) if ref.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(new))) if new.size else 0.0
    return float(np.max(np.abs(new - ref) / (np.abs(ref) + 1e-2 * scale)))


def maxnorm_err(new, ref):
    new = np.asarray(new, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref)) if ref.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(new))) if new.size else 0.0
    return float(np.max(np.abs(new - ref)) / scale)


# --------------------------------------------------------------------------- #
# Oracle-backed stand-in for spinn_b200.ops.CudaBackend (CPU tests of host logic)
# --------------------------------------------------------------------------- #
class OracleBackend:
    """Implements the backend protocol of spinn_b200.ops with oracle/closed_form.py.
    TEST ONLY: lets the tower / plug-in surface / problem classes run in the
    GPU-less build container.  Injected with ops.set_backend()."""

    name = "oracle"

    def __init__(self):
        self.calls = {"forward": 0, "backward": 0}

    @staticmethod
    def _np(t):
        return None if t is None else t.detach().cpu().numpy().astype(np.float64)

    def forward(self, dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, b):
        import torch
        from oracle import closed_form as cf
        self.calls["forward"] += 1
        nj = {(2, 0): 1, (2, 1): 3, (2, 2): 5, (2, 3): 6, (1, 0): 1, (1, 1): 2, (1, 2): 3}[(dim, level)]
        n = self._np
        if dim == 2:
            J = cf.jets2d(kind, n(x), n(y), np.concatenate([n(xf), n(xfix)]),
                          np.concatenate([n(yf), n(yfix)]), n(h), n(W), n(b))
        else:
            J = cf.jets1d(kind, n(x), np.concatenate([n(xf), n(xfix)]), n(h), n(W), n(b))
        return torch.tensor(J[:nj], dtype=torch.float32, device=x.device)

    def backward(self, dim, kind, level, x, y, xf, yf, xfix, yfix, h, W, gj, want_bias, mask=0):
        import torch
        from oracle import closed_form as cf
        self.calls["backward"] += 1
        n = self._np
        g = n(gj)
        if mask:                                   # rows whose bit is clear are undefined: zero them
            for a in range(g.shape[0]):
                if not (mask >> a) & 1:
                    g[a] = 0.0
        full = 6 if dim == 2 else 3
        if g.shape[0] < full:
            g = np.concatenate([g, np.zeros((full - g.shape[0],) + g.shape[1:])], 0)
        if dim == 2:
            G = cf.grads2d(kind, n(x), n(y), np.concatenate([n(xf), n(xfix)]),
                           np.concatenate([n(yf), n(yfix)]), n(h), n(W), g, n_free=xf.numel())
        else:
            G = cf.grads1d(kind, n(x), np.concatenate([n(xf), n(xfix)]), n(h), n(W), g, n_free=xf.numel())
        t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=x.device)
        return (t(G["d_xc"]), t(G["d_yc"]) if dim == 2 else None, t(G["d_h"]), t(G["d_W"]),
                t(G["d_b"]) if want_bias else None)


    def forward_sparse(self, level, x, y, xf, yf, xfix, yfix, h, W, b, perm, cutoff, stats=None):
        assert perm is not None and sorted(perm.tolist()) == list(range(xf.numel() + xfix.numel()))
        self.calls["forward_sparse"] = self.calls.get("forward_sparse", 0) + 1
        return self.forward(2, 0, level, x, y, xf, yf, xfix, yfix, h, W, b)

    def backward_sparse(self, level, x, y, xf, yf, xfix, yfix, h, W, gj, want_bias, mask, perm, cutoff,
                        stats=None):
        self.calls["backward_sparse"] = self.calls.get("backward_sparse", 0) + 1
        return self.backward(2, 0, level, x, y, xf, yf, xfix, yfix, h, W, gj, want_bias, mask)
