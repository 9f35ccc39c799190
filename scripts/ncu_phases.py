#!/usr/bin/env python
"""Aggregate ncu per-instruction counts of the fused tile kernel by PHASE (source-line ranges found by markers)."""
import collections, csv, io, re, subprocess, sys, os
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(root, "self-supervised-normal-depth-estimation_b200", "csrc")
def ranges(fname, markers):
    lines = open(os.path.join(csrc, fname)).read().split("\n")
    pts = []
    for name, pat in markers:
        for i, l in enumerate(lines):
            if pat in l:
                pts.append((i + 1, name)); break
        else:
            raise SystemExit("marker not found: " + pat)
    pts.sort()
    return pts
tile = ranges("snde_tile.cuh", [("reduce", "void warp_sum16"), ("B.row_sums", "struct RowSums"), ("C.co_row_sums", "struct CoSums"),
                                ("decode", "The tile program."), ("0.load", "---- phase 0"), ("A", "for (int k = 0; k < n_others"),
                                ("B", "---- phase B"), ("C", "---- phase C"), ("G", "---- phase G"), ("S", "if (do_smooth)"),
                                ("W.write", "---- write the depth gradient")])
math_ = ranges("snde_math.cuh", [("misc", "namespace snde {"), ("A.reproject", "struct Reproj {"), ("G.reproject_bwd(ieee)", "SNDE_HD float reproject_bwd("),
                                 ("A.taps", "struct Taps {"), ("ref.ssim", "struct Ssim {"), ("B.ssim_sums", "struct SsimS {"),
                                 ("B.ssim_bwd", "SNDE_HD void ssim_sums_bwd"), ("G.reproject", "SNDE_HD void reproject_nodiv"),
                                 ("A.blend", "SNDE_HD float tap_blend_fma"), ("S.edge", "SNDE_HD float edge_weight3"), ("pose", "SNDE_HD void mm3")])
def phase(f, ln):
    tab = tile if f == "snde_tile.cuh" else math_ if f == "snde_math.cuh" else None
    if tab is None: return "other:" + f
    name = "pre"
    for start, nm in tab:
        if ln >= start: name = nm
    return name
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg = collections.Counter(); smp = collections.Counter(); fname = None; hdr = None
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); ssi = hdr.index("# Samples"); continue
    if hdr and len(r) > ci:
        try: n = int(r[ci]); s = int(r[ssi]); ln = int(r[0])
        except ValueError: continue
        ph = phase(fname, ln); agg[ph] += n; smp[ph] += s
tot = sum(agg.values()); ts = max(1, sum(smp.values()))
print(f"total warp-inst {tot}  (x32 / 9052160 px = {tot*32/9052160:.0f} thread-inst per warped px at config2)")
groups = collections.Counter(); gs = collections.Counter()
for k, v in agg.items(): groups[k.split(".")[0]] += v; gs[k.split(".")[0]] += smp[k]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print(f"  {100*v/tot:5.1f}% inst  {100*smp[k]/ts:5.1f}% samples  {k}")
print("by phase:")
for k, v in sorted(groups.items(), key=lambda kv: -kv[1]): print(f"  {100*v/tot:5.1f}% inst  {100*gs[k]/ts:5.1f}% samples  {k}   ~{v*32/9052160:.0f} inst/px")
