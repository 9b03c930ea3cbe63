#!/usr/bin/env python
"""SASS opcode evidence for the Blackwell-specific instructions of the library (runs without a GPU).

    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt

For every kernel in pybinding_b200/csrc/build/*.o: counts of the opcodes that show how operands move and where the math
runs -- UBLKCP (cp.async.bulk, the bulk-copy / TMA engine), SYNCS (mbarrier), LDGSTS (cp.async), DMMA (FP64 tensor pipe),
LDS / LDG / STG, DFMA / FFMA -- from `cuobjdump -sass`.
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UBLKCP", "SYNCS", "UTMALDG", "UTCMMA", "LDTM", "LDGSTS", "DMMA", "HMMA", "LDS", "LDG", "STG", "STS", "DFMA", "FFMA", "F2F", "SHFL", "ATOM", "RED", "BAR"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "pybinding_b200", "csrc", "build", "*.o")))
    only = sys.argv[1:]
    for obj in objs:
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels = collections.OrderedDict()
        name = None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                kernels[name] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and name:
                op = m.group(1)
                kernels[name][op] += 1
                for w in WATCH:
                    if op.startswith(w):
                        kernels[name]["~" + w] += 1
        print("== {}".format(os.path.basename(obj)))
        for k, c in kernels.items():
            short = re.sub(r"pbk::\(anonymous namespace\)::", "", k)
            short = re.sub(r"\(.*", "", short)
            if only and not any(o in short for o in only):
                continue
            total = sum(v for kk, v in c.items() if not kk.startswith("~"))
            watched = " ".join("{}={}".format(w, c["~" + w]) for w in WATCH if c["~" + w])
            print("  {:<72s} {:6d} instr  {}".format(short[:72], total, watched))


if __name__ == "__main__":
    main()
