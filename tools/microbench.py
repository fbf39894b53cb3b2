#!/usr/bin/env python
"""Run the instruction-throughput probes (spinn_microbench2) and print lane-ops/s and the
fraction of the nominal 148 SM x 128 lanes x clock rate."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinn_b200 import _cabi  # noqa: E402

NAMES = ["FFMA a=a*m+d (shared regs)", "FFMA a=b*c+a (3 distinct regs)", "FFMA2 a=a*m+d (shared)",
         "FFMA2 a=b*c+a (3 distinct pairs)", "FFMA2 a=b*bcast+a", "8 FMUL2 + 8 FADD2 (independent chains)",
         "FFMA a=a*m+imm",
         "half FFMA2 + half FFMA (distinct)", "FADD distinct", "FMUL distinct", "MUFU.EX2",
         "16 FFMA2 bcast (no MUFU)", "FFMA2 operands reused pairwise", "16 FFMA2 bcast + 2 MUFU.EX2",
         "16 FFMA2 bcast + 4 MUFU.EX2", "FMUL2 a=a*m (shared)", "FADD2 a=a+d (shared)",
         "FMUL2 a=a*b (2 distinct pairs)", "FADD2 a=a+b (2 distinct pairs)",
         "fwd mix 7 FFMA2:4 FMUL2:3 FADD2:2 EX2"]


def sass_census():
    """Static check that each probe's loop holds the instructions its name claims (cuobjdump)."""
    import collections
    import re
    import subprocess
    for which in range(len(NAMES)):
        out = subprocess.run(["cuobjdump", "-sass", "-fun", f"_Z12probe_kernelILi{which}EEviPKfPf", _cabi.LIB_PATH],
                             capture_output=True, text=True).stdout
        ops = collections.Counter()
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
            if m and m.group(1) in ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2", "MUFU"):
                ops[m.group(1)] += 1
        print(f"# SASS census probe {which:2d} ({NAMES[which]}): " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items())))


def main():
    if "--sass" in sys.argv:
        sass_census()
    lib = _cabi.load()
    dev = torch.device("cuda", 0)
    src = (1.0 + 1e-3 * torch.rand(64, device=dev)).float()
    sink = torch.zeros(4, device=dev)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    nominal = sms * 128 * 1.965e9
    for occ_blocks in (8, 4, 2):
        print(f"--- {occ_blocks} CTAs of 256 threads per SM")
        for which, name in enumerate(NAMES):
            ops = ctypes.c_double(0)
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record()
                rc = lib.spinn_microbench(which, 2048, sms * occ_blocks, 256, src.data_ptr(), sink.data_ptr(),
                                           ctypes.byref(ops), torch.cuda.current_stream().cuda_stream)
                b.record()
                torch.cuda.synchronize()
                assert rc == 0
                best = min(best, a.elapsed_time(b))
            rate = ops.value / (best * 1e-3)
            print(f"{which:2d} {name:38s} {rate / 1e12:7.2f} T lane-ops/s  {100 * rate / nominal:6.1f}% of 128/clk/SM")


if __name__ == "__main__":
    main()
