#!/usr/bin/env python
"""Static size of the hot path: SASS instructions executed at least `frac` times per tile-step.
usage: python profiles/ncu_hotsize.py rep ntile_steps [frac=0.05]"""
import csv, io, subprocess, sys, collections
rep, ntile = sys.argv[1], float(sys.argv[2])
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = cur = fname = None
hot = collections.Counter(); tot = collections.Counter(); dyn = collections.Counter()
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; idx = {k: i for i, k in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "": cur = (fname, int(r[0]), r[1].strip()[:80]); continue
    try: ex = int(r[idx["Instructions Executed"]])
    except ValueError: continue
    tot[cur] += 1
    if ex >= frac * ntile: hot[cur] += 1; dyn[cur] += ex
print("static SASS: %d total, %d hot (executed >= %.2f x per tile-step) = %.1f KB" % (sum(tot.values()), sum(hot.values()), frac, sum(hot.values()) * 16 / 1024))
byf = collections.Counter()
for (f, l, s), c in hot.items(): byf[f] += c
print(dict(byf))
for key, c in sorted(hot.items(), key=lambda kv: -kv[1])[:40]:
    print("  %4d static  %7.1f dyn/ts  %s:%d  %s" % (c, dyn[key] / ntile, key[0], key[1], key[2]))
