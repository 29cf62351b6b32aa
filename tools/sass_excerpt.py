#!/usr/bin/env python
"""Instruction-mnemonic counts per kernel of the built libmppi_b200.so (cuobjdump -sass): the SASS evidence cited in
DESIGN.md.  usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTMALDG", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "FMNMX", "VIMNMX", "MUFU", "F2I", "LDS", "LDG",
        "STG", "SHFL", "BAR", "ACQBULK", "PREEXIT", "IMAD.WIDE", "LOP3"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ccv_mppi_path_tracker_b200", "libmppi_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input=sass, capture_output=True, text=True, check=True).stdout
    print("# SASS evidence for the shipped libmppi_b200.so (cuobjdump -sass, sm_100a): instruction-mnemonic counts per kernel.")
    print("# UTMALDG = TMA tensor load (cp.async.bulk.tensor), SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed FP32,")
    print("# FMNMX3 = three-input min/max, VIMNMX = integer min/max (.RELU clamp), MUFU = SFU, ELECT = elect.sync,")
    print("# ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch)\n")
    cur, counts, total = None, None, 0

    def flush():
        if cur is not None:
            body = ", ".join(f"{k} {counts[k]}" for k in KEYS if counts[k])
            print(f"{cur[:90]}\n    instructions {total}: {body}")

    for line in names.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            flush()
            cur, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            total += 1
            base = op.split(".")[0]
            for k in KEYS:
                if op == k or op.startswith(k + ".") or base == k:
                    counts[k] += 1
                    break
    flush()


if __name__ == "__main__":
    main()
