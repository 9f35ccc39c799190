#!/usr/bin/env python
"""Aggregate ncu per-instruction counts of the fused strip kernel by PHASE (source-line ranges found by markers).
usage: scripts/ncu_phases.py report.ncu-rep [warped_px]"""
import collections, csv, io, os, subprocess, sys
rep = sys.argv[1]
PX = float(sys.argv[2]) if len(sys.argv) > 2 else 9052160.0
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
strip = ranges("snde_strip.cuh", [("warp.exchange", "// warp primitives"), ("B.hsum_row", "SNDE_STRIP_FN void hsum_row"), ("BC.add2(rolling sums)", "SNDE_STRIP_FN F2 add2"),
                                  ("setup.lane", "struct StripLane {"), ("0.load_owner_row", "SNDE_STRIP_FN void load_owner_row"), ("A", "SNDE_STRIP_FN void phase_a"),
                                  ("step.head", "SNDE_STRIP_FN void step("), ("B.hsum", "horizontal sums of row t -"), ("B", "phase B (row"),
                                  ("C.hsum", "horizontal sums of the coefficient"), ("C", "phase C (row"), ("G", "rows t-2: phase G"),
                                  ("S", "smoothness of row t-2"), ("W.write", "GRAD && doC && in[0]"), ("setup.task", "void strip_task("),
                                  ("publish", "publish this pass")])
math_ = ranges("snde_math.cuh", [("misc", "namespace snde {"), ("A.reproject", "struct Reproj {"), ("G.reproject_bwd(ieee)", "SNDE_HD float reproject_bwd("),
                                 ("A.taps", "struct Taps {"), ("ref.ssim", "struct Ssim {"), ("B.ssim_sums", "struct SsimS {"),
                                 ("B.ssim_bwd", "SNDE_HD void ssim_sums_bwd"), ("AG.reproject_nodiv", "SNDE_HD void reproject_nodiv(const float *Kinv, const float *P, float xt, float yt, float d, Reproj &o) {"),
                                 ("G.reproject_bwd", "SNDE_HD float reproject_bwd_fast"),
                                 ("A.blend", "SNDE_HD float tap_blend_fma"), ("S.edge", "SNDE_HD float edge_weight3"), ("pose", "SNDE_HD void mm3")])
def phase(f, ln):
    tab = strip if f == "snde_strip.cuh" else math_ if f == "snde_math.cuh" else None
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
print(f"total warp-inst {tot}  (x32 / {PX:.0f} px = {tot*32/PX:.0f} thread-inst per warped px)")
groups = collections.Counter(); gs = collections.Counter()
for k, v in agg.items(): groups[k.split(".")[0]] += v; gs[k.split(".")[0]] += smp[k]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]): print(f"  {100*v/tot:5.1f}% inst  {100*smp[k]/ts:5.1f}% samples  {k}   ~{v*32/PX:.0f} inst/px")
print("by phase:")
for k, v in sorted(groups.items(), key=lambda kv: -kv[1]): print(f"  {100*v/tot:5.1f}% inst  {100*gs[k]/ts:5.1f}% samples  {k}   ~{v*32/PX:.0f} inst/px")
