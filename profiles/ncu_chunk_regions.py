#!/usr/bin/env python
"""Per-phase warp-instruction counts of k_traj_chunk per tile-step from an .ncu-rep.
usage: python profiles/ncu_chunk_regions.py rep ntile_steps [ntop]"""
import csv, io, subprocess, sys
rep, ntile = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = cur = fname = None
agg = {}
ti = ts = 0
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 9: continue
    if r[0] != "": cur = (fname, int(r[0]), r[1].strip()[:100]); agg.setdefault(cur, [0, 0, 0])
    else:
        try: sa, ins, thr = int(r[4]), int(r[7]), int(r[8])
        except ValueError: continue
        agg[cur][0] += sa; agg[cur][1] += ins; agg[cur][2] += thr; ti += ins; ts += sa
print("warp instructions per tile-step: %.1f   (total %.3g, stall samples %d)" % (ti / ntile, ti, ts))
lines = open("beep_b200/csrc/traj_chunk.cuh").read().split("\n")
def find(pat):
    for i, l in enumerate(lines):
        if pat in l: return i + 1
    raise SystemExit("marker not found: " + pat)
marks = [("helpers", "#define CHUNK_NWARP"), ("1a Iv gather", "/* 1a."), ("1b age matvec", "/* 1b."), ("2 wait+pop obs", "/* 2. the chunk"),
         ("3 scan", "/* 3. scan"), ("3 flush/draw", "/* draws on the pooled list"), ("4a obs", "/* 4a."), ("4b dinf+store", "/* 4b."),
         ("kernel+sched", "template <int NT> struct ChunkCfg"), ("end", "/* int32 populations of a batch")]
pos = [(n, find(p)) for n, p in marks]
def show(name, s):
    print(f"{name:16s} inst/ts {s[1]/ntile:7.1f} ({100*s[1]/ti:4.1f}%)  samples {100*s[0]/ts:4.1f}%  lanes {s[2]/max(s[1],1):4.1f}")
for k in range(len(pos) - 1):
    s = [0, 0, 0]
    for (ff, l, _), v in agg.items():
        if ff == "traj_chunk.cuh" and pos[k][1] <= l < pos[k + 1][1]:
            for q in range(3): s[q] += v[q]
    show(pos[k][0], s)
files = sorted({f for f, _, _ in agg})
for f in files:
    if f == "traj_chunk.cuh": continue
    s = [0, 0, 0]
    for (ff, l, _), v in agg.items():
        if ff == f:
            for q in range(3): s[q] += v[q]
    if s[1] / ti > 0.003: show(f, s)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
print("top lines by instructions:")
for (f, l, s_), (sa, i, th) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"  {f}:{l:<4d} inst/ts {i/ntile:6.1f} lanes {th/max(i,1):4.1f} samp {100*sa/ts:5.1f}%  {s_[:95]}")
