"""The C ABI driven from plain C++ (tests/cabi_native/cabi_check.cpp: CUDA runtime + include/spinn_b200.h,
no Python, no torch in that process) -- what a non-Python host binding does.  The program is built by
__graft_entry__.build()."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cabi_native", "cabi_check")


def test_native_host_program_is_built_and_links_the_library():
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    out = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libspinn_b200.so" in out and "not found" not in out.split("libspinn_b200.so")[1].splitlines()[0]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(), ("4097", "300", "41"), ("1", "3", "0")])
def test_c_abi_from_a_plain_cpp_host(shape):
    r = subprocess.run([EXE, *shape], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cabi_check OK" in r.stdout, r.stdout + r.stderr
