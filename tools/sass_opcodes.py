#!/usr/bin/env python3
"""Per-kernel counts of the SASS opcodes of libsdrgpu.so that do not exist before Hopper / Blackwell (packed f32x2
arithmetic, TMA bulk copies, mbarrier transactions) plus LDGSTS (cp.async).  usage: python tools/sass_opcodes.py > profiles/rNN_sass_blackwell_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = re.compile(r"^(FFMA2|FADD2|FMUL2|UBLKCP|SYNCS|LDGSTS|UTMA|UTC|LDTM|REDUX)")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sdrtrunk_b200", "libsdrgpu.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangled = subprocess.run(["cu++filt"], input=sass, capture_output=True, text=True).stdout or sass
    counts, name = collections.OrderedDict(), None
    for line in demangled.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            name = m.group(1)[:110]
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name and WANT.match(m.group(1)):
            counts[name][m.group(1)] += 1
    print("# SASS opcodes of libsdrgpu.so (cuobjdump -sass) that do not exist before Hopper / Blackwell, per kernel")
    print("# FFMA2 / FADD2 / FMUL2 = packed f32x2 fused multiply-add / add / multiply (sm_100: both rails correctly rounded);")
    print("# UBLKCP.S.G = cp.async.bulk global->shared (TMA engine, 1-D); SYNCS.* = mbarrier arrive / transaction arrive / try_wait;")
    print("# LDGSTS = cp.async (Ampere+)")
    for name, c in counts.items():
        if c:
            print(name)
            for op, n in sorted(c.items()):
                print("    %-40s %d" % (op, n))


if __name__ == "__main__":
    main()
