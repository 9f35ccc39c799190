#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction counts by CUDA source line.
usage: scripts/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0]); fname = None; hdr = None
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); ssi = hdr.index("# Samples"); continue
    if hdr and len(r) > ci:
        try: n = int(r[ci]); s = int(r[ssi]); ln = int(r[0])
        except ValueError: continue
        agg[(fname, ln, r[1].strip()[:110])][0] += n; agg[(fname, ln, r[1].strip()[:110])][1] += s
tot = sum(v[0] for v in agg.values()); tots = max(1, sum(v[1] for v in agg.values()))
print("total warp-inst", tot, "samples", tots)
for (f, l, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:5.1f}% smp{100*s/tots:5.1f}% {f}:{l}: {src}")
