#!/usr/bin/env python
"""Per-source-line stall samples by reason from an .ncu-rep (source page).
usage: python profiles/ncu_stalls.py rep [reason ...]   (default: long_sb short_sb wait no_inst barrier)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
reasons = sys.argv[2:] or ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_barrier", "stall_branch_resolving"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = cur = fname = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; idx = {k: i for i, k in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "": cur = (fname, int(r[0]), r[1].strip()[:90]); continue
    for k in reasons + ["# Samples"]:
        try: v = int(r[idx[k]])
        except ValueError: v = 0
        agg[cur][k] += v; tot[k] += v
print("total samples", tot["# Samples"], {k: "%.1f%%" % (100 * tot[k] / max(tot["# Samples"], 1)) for k in reasons})
for k in reasons:
    print("==", k)
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1][k])[:14]:
        if c[k] == 0: break
        print("  %5.1f%%  %s:%d  %s" % (100 * c[k] / max(tot[k], 1), key[0], key[1], key[2]))
