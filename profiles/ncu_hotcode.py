#!/usr/bin/env python
"""Hot code footprint of a kernel from an .ncu-rep: SASS instructions by how often they execute.
usage: python profiles/ncu_hotcode.py rep   (the L1.5 instruction cache holds 32 KB = 2048 instructions)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
ex = []
for r in rows:
    if r and r[0] in ("Address", "Line No", "#"): hdr = r; continue
    if hdr is None: continue
    try:
        i_exec = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
    except ValueError:
        i_exec = None
    if i_exec is None or len(r) <= i_exec: continue
    try: ex.append((int(r[i_exec]), r[1] if len(r) > 1 else ""))
    except ValueError: pass
if not ex:
    print("header:", hdr); sys.exit(1)
tot = sum(e for e, _ in ex)
n = len(ex)
print(f"SASS instructions {n} ({n*16/1024:.0f} KB), executed at least once {sum(1 for e,_ in ex if e>0)}")
mx = max(e for e, _ in ex)
for frac in (0.5, 0.1, 0.01, 0.001):
    k = sum(1 for e, _ in ex if e >= frac * mx)
    print(f"  executed >= {frac:g} x the hottest: {k} instructions ({k*16/1024:.1f} KB)")
srt = sorted((e for e, _ in ex), reverse=True)
acc = 0
for k, e in enumerate(srt):
    acc += e
    if acc >= 0.9 * tot:
        print(f"  90% of the dynamic instructions come from {k+1} static instructions ({(k+1)*16/1024:.1f} KB)"); break
acc = 0
for k, e in enumerate(srt):
    acc += e
    if acc >= 0.99 * tot:
        print(f"  99% of the dynamic instructions come from {k+1} static instructions ({(k+1)*16/1024:.1f} KB)"); break
