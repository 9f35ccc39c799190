#!/usr/bin/env python
"""Stall-sample breakdown per warp role of the strip kernel: splits the sampled SASS addresses at the second BAR.SYNC
row loop (SSIM warp code comes first, geometry warp code second).  usage: scripts/ncu_roles.py report.ncu-rep split_hex_offset"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; split = int(sys.argv[2], 16)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; recs = []
for r in rows:
    if len(r) > 5 and "# Samples" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: n = int(d["# Samples"])
        except (ValueError, KeyError): continue
        recs.append((int(d["Address"], 16), n, d))
base = min(a for a, _, _ in recs)
stall_cols = [c for c in hdr if c.startswith("stall_") and "not_issued" not in c.lower() and "(Not Issued)" not in c]
for name, sel in (("SSIM warp code", lambda a: a - base < split), ("geometry warp code", lambda a: a - base >= split)):
    c = collections.Counter(); tot = 0; inst = 0
    for a, n, d in recs:
        if not sel(a): continue
        tot += n
        try: inst += int(d.get("Instructions Executed", "0") or 0)
        except ValueError: pass
        for col in stall_cols:
            v = d.get(col)
            if v not in (None, "", "0"): c[col[6:]] += int(v)
    s = sum(c.values()) or 1
    print(f"{name}: {tot} samples, {inst} warp-instructions executed")
    print("   " + "  ".join(f"{k}:{100*v/s:.1f}%" for k, v in c.most_common(9)))
