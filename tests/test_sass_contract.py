"""Static checks of the shipped libspinn_b200.so (cuobjdump, no GPU needed): the per-pair instruction
counts that bench.py's roofline section and DESIGN.md section 4 quote are read off the SASS of the hot
inner loops; TMA bulk copies and packed f32x2 arithmetic are present; the hot kernels do not spill."""
import collections
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass():
    import sass_census
    from spinn_b200 import _cabi, build
    build.build()
    funcs = {}
    for name, body in sass_census.functions(_cabi.LIB_PATH):
        funcs[subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()] = body
    return sass_census, funcs


def _hot_loop(sass, prefix):
    census, funcs = sass
    hits = [k for k in funcs if k.startswith("void " + prefix)]
    assert len(hits) == 1, (prefix, hits)
    n_fp32, loop = census.census(funcs[hits[0]])
    ops = collections.Counter(op for _, op, _ in loop)
    ex2 = sum(v for k, v in ops.items() if k.startswith("MUFU.EX2"))
    mufu = sum(v for k, v in ops.items() if k.startswith("MUFU"))
    return n_fp32, ex2, mufu, ops, funcs[hits[0]]


# kernel instantiation (as launched by default at the BASELINE size) -> (EX2 per packed evaluation,
# FP32-pipe instructions per pair, MUFU per pair) -- the numbers in bench.KERNEL_COST
CASES = [
    ("g2k1_fwd_kernel<2, false, 256, 2, 4, false, false>", 2, ("gaussian", "fwd")),
    ("g2k1_fwd_kernel<2, false, 256, 2, 4, true, false>", 2, ("gaussian", "fwd")),
    ("g2k1_bwd_kernel<4, 256, 2, 256, 1>", 2, ("gaussian", "bwd_structured")),
    ("g2k1_bwd_kernel<2, 256, 2, 256, 0>", 2, ("gaussian", "bwd_general")),
    ("s2k1_fwd_kernel<2, false, 256, 2, 2, false, false>", 4, ("softplus", "fwd")),
    ("s2k1_bwd_kernel<2, 256, 2, 256, 24>", 4, ("softplus", "bwd_structured")),
    ("s2k1_bwd_kernel<2, 256, 2, 256, 31>", 4, ("softplus", "bwd_general")),
]


@pytest.mark.parametrize("prefix,ex2_per_eval,key", CASES)
def test_inner_loop_instruction_counts_match_the_roofline_table(sass, prefix, ex2_per_eval, key):
    import bench
    n_fp32, ex2, mufu, ops, _ = _hot_loop(sass, prefix)
    assert ex2 > 0 and ex2 % ex2_per_eval == 0
    evals = ex2 // ex2_per_eval                  # packed evaluations (two pairs each) in one loop trip
    cost = bench.KERNEL_COST[key]
    assert n_fp32 == cost["instr"] * evals, (prefix, n_fp32, evals, cost)
    assert mufu == 2 * cost["mufu"] * evals, (prefix, mufu, evals, cost)          # MUFU is per lane
    flop = 2 * (2 * ops["FFMA2"] + ops["FMUL2"] + ops["FADD2"]) / (2 * evals)      # executed, per pair
    assert flop == cost["flop_sass"] <= cost["flop"], (prefix, flop, cost)
    assert ops["FFMA2"] + ops["FMUL2"] + ops["FADD2"] == n_fp32                    # everything is packed


def test_hot_kernels_use_tma_and_do_not_spill(sass):
    for prefix in [c[0] for c in CASES]:
        _, _, _, ops, body = _hot_loop(sass, prefix)
        text = "\n".join(body)
        assert "UBLKCP" in text, prefix                                          # cp.async.bulk (TMA)
        assert "SYNCS" in text, prefix                                           # mbarrier
        assert ops["STL"] == 0 and ops["LDL"] == 0, prefix                       # hot loop free of local memory
        # (outside the loop only the epilogue's sinf() slow path may touch the stack)
        if ", true, false>" not in prefix:
            assert not re.search(r"\b(STL|LDL)\b", text), prefix
