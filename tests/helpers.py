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
