#!/usr/bin/env python
"""Per-region instruction counts of k_traj per tile-step from an .ncu-rep.
usage: python profiles/ncu_regions.py rep ntile_steps"""
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
print("instructions per tile-step: %.1f" % (ti / ntile))
lines = open("beep_b200/csrc/kernels.cuh").read().split("\n")
def find(pat, start=0):
    for i in range(start, len(lines)):
        if pat in lines[i]: return i + 1
    return None
# regions of the fast tile step (tile_step_fast) and of the kernel body
f0 = find("__device__ __forceinline__ unsigned tile_step_fast(")
marks = [("generic", "__device__ __forceinline__ unsigned tile_step("), ("fast:setup", None), ("fast:FOI", "/* 1. state.cc:547-578"), ("fast:scan", "/* 2. scan:"),
         ("fast:draw", "/* 3. draws on the dense"), ("fast:trmask", "/* transitions with events, warp-wide"), ("fast:obs", "/* 4a. obsmodel"),
         ("fast:update", "/* 4b. state.cc:460-485"), ("kernel", "/* tile index -> (particle, area)"), ("end", "/* K4  observation model")]
pos = []
for name, pat in marks:
    if name == "generic": ln = find(pat)
    elif name == "fast:setup": ln = f0
    elif name.startswith("fast:"): ln = find(pat, f0)
    else: ln = find(pat, f0)
    pos.append((name, ln))
pos = [("helpers", find("/* K1-K3  trajectory kernel"))] + pos
def reg(f, a, b):
    s = [0, 0, 0]
    for (ff, l, _), (sa, i, th) in agg.items():
        if ff == f and a <= l < b: s[0] += sa; s[1] += i; s[2] += th
    return s
for k in range(len(pos) - 1):
    s = reg("kernels.cuh", pos[k][1], pos[k + 1][1])
    print(f"{pos[k][0]:12s} inst/ts {s[1]/ntile:7.1f} ({100*s[1]/ti:4.1f}%)  samples {100*s[0]/ts:4.1f}%  lanes {s[2]/max(s[1],1):4.1f}")
for f in ("beep_rng.h", "sync.h", "sm_30_intrinsics.hpp"):
    s = [0, 0, 0]
    for (ff, l, _), (sa, i, th) in agg.items():
        if ff == f: s[0] += sa; s[1] += i; s[2] += th
    print(f"{f:12s} inst/ts {s[1]/ntile:7.1f} ({100*s[1]/ti:4.1f}%)  samples {100*s[0]/ts:4.1f}%  lanes {s[2]/max(s[1],1):4.1f}")
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
print("top lines by instructions:")
for (f, l, s_), (sa, i, th) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"  {f}:{l:<4d} inst/ts {i/ntile:6.1f} lanes {th/max(i,1):4.1f} samp {100*sa/ts:5.1f}%  {s_[:95]}")
