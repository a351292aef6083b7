#!/usr/bin/env python
"""profiles/k_ofdm.sass: instruction histogram of the 32K chain-mode OFDM kernel in the built library plus the
TMA / mbarrier instructions with two lines of context (the evidence for `cp.async.bulk` staging: UBLKCP, SYNCS).

usage: tools/sass_excerpt.py [library] [output]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUNC = "_ZN3t2k6k_ofdmILi14ELi512ELb1ELi0ELi2EEEvNS_8OfdmArgsE"


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gr-dvbt2ll_b200", "libdvbt2ll_cuda.so")
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "k_ofdm.sass")
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", FUNC, lib], capture_output=True, text=True).stdout
    lines = [ln for ln in txt.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln)]
    hist = collections.Counter()
    for ln in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            hist[m.group(1)] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs = "?"
    for i, ln in enumerate(res.splitlines()):
        if FUNC in ln:
            m = re.search(r"REG:(\d+) STACK:(\d+)", res.splitlines()[i + 1])
            regs = "%s registers, %s bytes of stack" % (m.group(1), m.group(2)) if m else "?"
    with open(out, "w") as f:
        f.write("# k_ofdm<14,512,C16,complex64,SPLIT=2> (32K chain-mode OFDM kernel), sm_100a SASS of gr-dvbt2ll_b200/libdvbt2ll_cuda.so (final round-2 build)\n")
        f.write("# regenerate: python tools/sass_excerpt.py   (cuobjdump -sass -fun %s)\n" % FUNC)
        f.write("# %d instructions, %s; histogram (mnemonic: count):\n" % (len(lines), regs))
        for k, v in hist.most_common():
            f.write("#   %-12s %d\n" % (k, v))
        f.write("#\n# TMA / mbarrier instructions with context (UBLKCP = cp.async.bulk: staging runs and the descriptor list; SYNCS = mbarrier):\n")
        keep = set()
        for i, ln in enumerate(lines):
            if re.search(r"UBLKCP|SYNCS|FENCE|ELECT", ln):
                keep.update(range(max(0, i - 2), min(len(lines), i + 3)))
        last = -2
        for i in sorted(keep):
            if i != last + 1:
                f.write("        ...\n")
            f.write(lines[i].rstrip() + "\n")
            last = i
    print(open(out).read()[:1500])


if __name__ == "__main__":
    main()
