#!/usr/bin/env python
"""Per-stall-reason totals of one kernel, restricted to a SASS line range: ncu_stalls.py rep regex [lo hi]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hs[0]; end = hs[1] - 1 if len(hs) > 1 else len(rows)
h = rows[start]
body = [r for r in rows[start + 1:end] if len(r) == len(h)]
cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = {h[i]: 0 for i in cols}
for r in body[lo:hi]:
    for i in cols:
        try: tot[h[i]] += int(r[i])
        except ValueError: pass
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print("%-26s %8d %6.2f%%" % (k, v, 100.0 * v / max(s, 1)))
