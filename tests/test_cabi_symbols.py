"""The C-ABI library loads and exports every symbol include/spinn_b200.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "spinn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spinn\w*)\s*\(", src)))


def test_header_symbols_are_exported():
    from spinn_b200 import _cabi, build
    build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_cabi.SIGNATURES), set(names) ^ set(_cabi.SIGNATURES)


def test_loader_signatures_and_version():
    from spinn_b200 import _cabi
    lib = _cabi.load()
    assert lib.spinn_abi_version() == _cabi.ABI_VERSION == 2
    assert lib.spinn_num_jets(2, 2) == 5 and lib.spinn_num_jets(2, 3) == 6 and lib.spinn_num_jets(1, 2) == 3
    assert lib.spinn_num_jets(1, 3) < 0
    # workspace sizing is pure host arithmetic (device query falls back to 148 SMs without a GPU)
    assert lib.spinn2d_fwd_workspace_bytes(1000, 4096, 1) >= 3 * 2048 * 16
    assert lib.spinn2d_bwd_workspace_bytes(4194304, 4096, 1) >= 4 * 2097152 * 16


def test_argument_validation_needs_no_gpu():
    from spinn_b200 import _cabi
    lib = _cabi.load()
    rc = lib.spinn2d_jets_fwd(5, 2, 10, 1, 0, 1, *([None] * 7), 0, *([None] * 4), 0, None)
    assert rc == -1 and b"kind" in lib.spinn_last_error()
    rc = lib.spinn1d_jets_fwd(0, 3, 10, 1, 0, 1, *([None] * 4), 0, *([None] * 4), 0, None)
    assert rc == -1 and b"level" in lib.spinn_last_error()
