"""ctypes binding of libspinn_b200.so (C ABI declared in include/spinn_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this
module raises.  Build the library with ``python -m spinn_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspinn_b200.so")

_c_fp = ctypes.c_void_p      # device pointers travel as integers
_c_int = ctypes.c_int
_c_size = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/spinn_b200.h one to one
SIGNATURES = {
    "spinn_abi_version": (_c_int, []),
    "spinn_last_error": (ctypes.c_char_p, []),
    "spinn_num_jets": (_c_int, [_c_int, _c_int]),
    "spinn2d_fwd_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int]),
    "spinn2d_jets_fwd": (_c_int, [_c_int] * 6 + [_c_fp] * 7 + [_c_int] + [_c_fp] * 4 + [_c_size, _c_fp]),
    "spinn2d_bwd_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int]),
    "spinn2d_jets_bwd": (_c_int, [_c_int] * 7 + [_c_fp] * 7 + [_c_int] + [_c_fp] * 8 + [_c_size, _c_fp]),
    "spinn1d_fwd_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int]),
    "spinn1d_jets_fwd": (_c_int, [_c_int] * 6 + [_c_fp] * 4 + [_c_int] + [_c_fp] * 4 + [_c_size, _c_fp]),
    "spinn1d_bwd_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int]),
    "spinn1d_jets_bwd": (_c_int, [_c_int] * 7 + [_c_fp] * 4 + [_c_int] + [_c_fp] * 7 + [_c_size, _c_fp]),
    "spinn2d_sparse_fwd_workspace_bytes": (_c_size, [_c_int, _c_int]),
    "spinn2d_sparse_bwd_workspace_bytes": (_c_size, [_c_int, _c_int]),
    "spinn2d_jets_fwd_sparse": (_c_int, [_c_int] * 6 + [_c_fp] * 7 + [_c_int] + [_c_fp] * 3 + [ctypes.c_float]
                                + [_c_fp] * 3 + [_c_size, _c_fp]),
    "spinn2d_jets_bwd_sparse": (_c_int, [_c_int] * 7 + [_c_fp] * 7 + [_c_int] + [_c_fp] * 3 + [ctypes.c_float]
                                + [_c_fp] * 7 + [_c_size, _c_fp]),
    "spinn2d_poisson_residual_workspace_bytes": (_c_size, [_c_int]),
    "spinn2d_poisson_residual": (_c_int, [_c_int, ctypes.c_double] + [_c_fp] * 3 + [ctypes.c_float] * 5
                                 + [_c_fp] * 3 + [_c_size, _c_fp]),
    "spinn1d_ode3_residual": (_c_int, [_c_int, ctypes.c_double, _c_fp, _c_fp, ctypes.c_float, _c_fp, _c_fp, _c_fp,
                                       _c_size, _c_fp]),
    "spinn2d_cavity_residual": (_c_int, [_c_int, ctypes.c_float, _c_fp, _c_fp, _c_fp, _c_fp, _c_size, _c_fp]),
    "spinn2d_poisson_interior_step": (_c_int, [_c_int, _c_int, ctypes.c_double, _c_int, _c_int] + [_c_fp] * 7
                                      + [_c_int] + [_c_fp] * 2 + [ctypes.c_float] * 5 + [_c_fp] * 7
                                      + [_c_fp, _c_size, _c_fp, _c_size, _c_fp]),
    "spinn_set_tuning": (None, [ctypes.c_char_p]),
    "spinn_microbench": (_c_int, [_c_int] * 4 + [_c_fp, _c_fp, ctypes.POINTER(ctypes.c_double), _c_fp]),
}

_lib = None
ABI_VERSION = 2


class SpinnNativeError(RuntimeError):
    pass


def load():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpinnNativeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. "
            "Run `python -m spinn_b200.build` (there is no CPU/PyTorch fallback).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.spinn_abi_version() != ABI_VERSION:
        raise SpinnNativeError("libspinn_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().spinn_last_error().decode("utf-8", "replace")
        raise SpinnNativeError(f"{what} failed (code {rc}): {msg}")
