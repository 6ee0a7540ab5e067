#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG / UBLKCP / UBLKRED, FP64 tensor -> DMMA) in the built library.

    python scripts/sass_summary.py > profiles/r2/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rosnet_b200", "lib", "librosnet_b200.so")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTCBAR.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UBLKCP", "UBLKRED", "DMMA", "HMMA", "UCGABAR", "SYNCS", "MEMBAR.ALL.GPU", "REDG", "RED.E"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::|rnb::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
    print("# cuobjdump -sass rosnet_b200/lib/librosnet_b200.so : instruction counts per kernel (sm_100a)")
    print("# nvcc: " + subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1])
    total = collections.Counter()
    for fn, c in per.items():
        if not c:
            continue
        total.update(c)
        print(f"{fn}\n    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    print("TOTAL\n    " + "  ".join(f"{k}={v}" for k, v in sorted(total.items())))


if __name__ == "__main__":
    sys.exit(main())
