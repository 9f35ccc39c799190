#!/usr/bin/env python
"""Static SASS statistics of one kernel: the largest loop (by backward branch) and its opcode histogram.
usage: scripts/sass_loop.py object.o mangled_kernel_name"""
import collections, re, subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in txt.split("\n"):
    m = re.match(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
loops = []
for a, t in ins:
    if "BRA" in t:
        mm = re.search(r'0x([0-9a-f]+)', t)
        if mm and int(mm.group(1), 16) < a: loops.append((a - int(mm.group(1), 16), int(mm.group(1), 16), a))
loops.sort(reverse=True)
for ln, lo, hi in loops[:6]: print(f"loop {lo:#x}..{hi:#x}: {ln // 16 + 1} instructions")
# innermost big loop = the t loop: pick the largest loop that contains no other big loop start.. simply the 3rd arg or default: largest < 90% of total
pick = None
for ln, lo, hi in loops:
    if pick is None or (ln < pick[0] and ln > 0.3 * loops[0][0]): pick = (ln, lo, hi)
ln, lo, hi = pick
c = collections.Counter()
for a, t in ins:
    if lo <= a <= hi:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        c[t.split()[0].split('.')[0]] += 1
print(f"histogram of loop {lo:#x}..{hi:#x} ({sum(c.values())} instructions):")
print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(60)))
