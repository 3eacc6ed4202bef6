#!/usr/bin/env python
"""Static size of the HOT code of a kernel by source line, from an .ncu-rep: for every source line the number of SASS
instructions that executed at least 1 % as often as the hottest one (the footprint that competes for the 32 KB L1.5
instruction cache).  usage: python profiles/ncu_hot_lines.py rep [ntop]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
ins = []
for r in rows:
    if r and r[0] in ("Address", "#"): hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        ins.append((r[hdr.index("Address")], int(r[hdr.index("Instructions Executed")])))
    except (ValueError, IndexError):
        pass
mx = max(e for _, e in ins)
hot = {a for a, e in ins if e >= 0.01 * mx}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, line, text = None, None, ""
cnt = collections.Counter(); txt = {}
seen = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != "":
        line, text = (fname, r[0]), r[1]
        txt[line] = text
        continue
    if len(r) > 2 and r[2] in hot and r[2] not in seen:
        seen.add(r[2]); cnt[line] += 1
tot = sum(cnt.values())
print(f"hot instructions {len(hot)} ({len(hot) * 16 / 1024:.1f} KB); attributed {tot}")
byfile = collections.Counter()
for (f, l), c in cnt.items(): byfile[f] += c
for f, c in byfile.most_common(): print(f"  {f:28s} {c:5d} instr {c * 16 / 1024:5.1f} KB")
for (f, l), c in cnt.most_common(ntop): print(f"  {f}:{l:5s} {c:4d}  {txt[(f, l)][:110]}")
