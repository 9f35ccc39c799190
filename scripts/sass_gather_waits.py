#!/usr/bin/env python
"""Which instructions of the geometry warp's row loop wait on a scoreboard that global loads (tap gathers, depth-gradient
read) signal?  Decodes the control bits of the SASS (write barrier: bits 46-48, wait mask: bits 52-57 of the upper word).
A wait that sits between the gathers and the next row step defeats the software prefetch.
usage: scripts/sass_gather_waits.py object.o mangled_kernel_name"""
import re, subprocess, sys
from collections import Counter


def analyse(obj, fun):
    """-> (loop_lo, loop_hi, {scoreboard: loads}, [(address, wait_mask, text)] of the in-loop instructions that wait on one
    of those scoreboards)"""
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    lines = txt.split('\n'); i = 0; res = []
    while i < len(lines):
        m = re.match(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/', lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r'^\s+/\* (0x[0-9a-f]+) \*/', lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                res.append((int(m.group(1), 16), m.group(2), (hi >> 46) & 7, (hi >> 52) & 0x3f)); i += 2; continue
        i += 1
    # the geometry row loop: the innermost backward branch around the tap gathers
    pre = [a for a, t, _, _ in res if "LDG.E.CONSTANT" in t]
    loops = []
    for a, t, _, _ in res:
        if "BRA" in t:
            mm = re.search(r'0x([0-9a-f]+)', t)
            if mm and int(mm.group(1), 16) < a: loops.append((int(mm.group(1), 16), a))
    cands = [l for l in loops if sum(l[0] <= x <= l[1] for x in pre) >= 24]
    lo, hi = min(cands, key=lambda l: l[1] - l[0])
    isld = lambda t: re.search(r'\b(LDG|LD)\.E', t) is not None
    sbs = Counter(wb for a, t, wb, wm in res if lo <= a <= hi and isld(t) and wb != 7)
    waits = [(a, wm, t) for a, t, wb, wm in res if lo <= a <= hi and any((wm >> sb) & 1 for sb in sbs)]
    return lo, hi, dict(sbs), waits


if __name__ == "__main__":
    lo, hi, sbs, waits = analyse(sys.argv[1], sys.argv[2])
    print(f"geometry row loop {lo:#x}..{hi:#x}")
    print("scoreboards signalled by global loads in the loop:", sbs)
    for a, wm, t in waits:
        print(f"  {a:05x} wait={wm:06b} {t[:84]}")
