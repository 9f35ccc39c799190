#!/usr/bin/env python
"""Static SASS statistics of the strip kernel: the two row loops (the innermost backward branches that contain a
BAR.SYNC: SSIM warp first, geometry warp second) with their opcode histograms, spills (STL/LDL) inside them.
usage: scripts/sass_roles.py object.o mangled_kernel_name"""
import collections, re, subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in txt.split("\n"):
    m = re.match(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
bars = [a for a, t in ins if t.startswith("BAR.SYNC")]
loops = []
for a, t in ins:
    if "BRA" in t:
        mm = re.search(r'0x([0-9a-f]+)', t)
        if mm and int(mm.group(1), 16) < a: loops.append((int(mm.group(1), 16), a))
for b in bars:
    inner = [l for l in loops if l[0] <= b <= l[1]]
    if not inner: continue
    lo, hi = min(inner, key=lambda l: l[1] - l[0])
    if (hi - lo) // 16 < 100: continue
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            t = re.sub(r'^@!?U?P\d+\s+', '', t)
            c[t.split()[0].split('.')[0]] += 1
    print(f"row loop {lo:#x}..{hi:#x}: {sum(c.values())} instructions")
    print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(60)))
