#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report (sampling), with stall reason split and CUDA source line.
usage: scripts/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; out = []
for r in rows:
    if len(r) > 5 and r[0] in ("Address", "Line No") or (len(r) > 5 and "Source" in r[:3] and "# Samples" in r):
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: n = int(d["# Samples"])
        except (ValueError, KeyError): continue
        out.append((n, d))
tot = sum(n for n, _ in out) or 1
print("total samples", tot)
stall_cols = [c for c in hdr if c.startswith("stall_")]
for n, d in sorted(out, key=lambda x: -x[0])[:top]:
    reasons = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols if d.get(c) not in (None, "", "0")), reverse=True)[:3]
    print(f"{100*n/tot:5.2f}%  {d.get('Address','')[-6:]:>6}  {d.get('Source','')[:70]:70s}  " + " ".join(f"{k}:{v}" for v, k in reasons))
