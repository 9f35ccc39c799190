#!/usr/bin/env python
"""Which instructions of the geometry warp's row loop wait on a scoreboard that global loads (tap gathers, depth-gradient
read) signal?  Decodes the control bits of the SASS (write barrier: bits 46-48, wait mask: bits 52-57 of the upper word).
A wait that sits between the gathers and the next row step defeats the software prefetch.
usage: scripts/sass_gather_waits.py object.o mangled_kernel_name"""
import re, subprocess, sys
from collections import Counter
txt = subprocess.run(["cuobjdump", "-sass", "-fun", sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
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
cands = [l for l in loops if any(l[0] <= x <= l[1] for x in pre)]
lo, hi = min(cands, key=lambda l: l[1] - l[0])
print(f"geometry row loop {lo:#x}..{hi:#x}")
isld = lambda t: re.search(r'\b(LDG|LD)\.E', t) is not None
sbs = Counter(wb for a, t, wb, wm in res if lo <= a <= hi and isld(t) and wb != 7)
print("scoreboards signalled by global loads in the loop:", dict(sbs))
for a, t, wb, wm in res:
    if not (lo <= a <= hi): continue
    tag = ""
    if isld(t): tag = f"   <- load, signals SB{wb}"
    hit = [sb for sb in sbs if (wm >> sb) & 1]
    if hit or tag:
        if tag and not hit and 'CONSTANT' in t: continue        # the gathers themselves: too many to list
        print(f"  {a:05x} wait={wm:06b} {t[:84]}{tag}")
