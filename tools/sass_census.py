#!/usr/bin/env python
"""Instruction census of the hottest loop of a kernel in libspinn_b200.so (cuobjdump -sass): finds the
backward branch whose body holds the most FP32-pipe instructions and counts the opcodes in it.

    python tools/sass_census.py <mangled-name-substring> [--dump]
"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from spinn_b200 import _cabi  # noqa: E402

FP32 = ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2")


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
        elif cur:
            body.append(line)
    if cur:
        yield cur, body


def census(body):
    ins = []
    for line in body:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    loops = []
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr:
                loops.append((int(t.group(1), 16), addr))
    best = None
    for lo, hi in loops:
        if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops):
            continue                                     # not innermost
        loop = [i for i in ins if lo <= i[0] <= hi]
        n = sum(1 for i in loop if i[1].split(".")[0] in FP32)
        if best is None or n > best[0]:
            best = (n, loop)
    return best


def main():
    pat = sys.argv[1]
    for name, body in functions(_cabi.LIB_PATH):
        if pat not in name:
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        best = census(body)
        if not best:
            continue
        n, loop = best
        c = collections.Counter(i[1].split(".")[0] + (".EX2" if ".EX2" in i[1] else ".RCP" if ".RCP" in i[1] else ".LG2" if ".LG2" in i[1] else "") if i[1].startswith("MUFU") else i[1].split(".")[0] for i in loop)
        print(f"## {dem[:110]}")
        print(f"   hottest loop: {len(loop)} instructions, {n} FP32-pipe: " + ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
        if "--dump" in sys.argv:
            for addr, op, rest in loop:
                print(f"      /*{addr:04x}*/ {op}{rest}")


if __name__ == "__main__":
    main()
