#!/usr/bin/env python
"""Top stall sites of one kernel in an .ncu-rep (SASS view):  python tools/ncu_hot.py rep.ncu-rep <kernel regex> [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blk = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # which matching launch / view (SASS and source views alternate)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hi[blk]; end = hi[blk + 1] - 1 if len(hi) > blk + 1 else len(rows)
h = rows[start]
si = h.index("Warp Stall Sampling (All Samples)"); ii = h.index("Instructions Executed"); src = h.index("Source")
body = [r for r in rows[start + 1:end] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body); toti = sum(int(r[ii]) for r in body)
print("kernel", rows[start - 1][1][:80], "samples", tot, "warp-instructions", toti, "SASS lines", len(body))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:topn]
for i in sorted(idx):
    r = body[i]
    print("%5d %6.2f%% exec=%9s  %s" % (i, 100.0 * int(r[si]) / max(tot, 1), r[ii], r[src].strip()[:110]))
# bucketed view: share of samples and of executed warp-instructions per 100 SASS lines, and by opcode
print("--- buckets of 100 SASS lines: samples% / instr%")
for b in range(0, len(body), 100):
    s = sum(int(r[si]) for r in body[b:b + 100]); n = sum(int(r[ii]) for r in body[b:b + 100])
    if s * 200 > tot or n * 200 > toti:
        print("%5d-%5d  %6.2f%%  %6.2f%%" % (b, b + 99, 100.0 * s / tot, 100.0 * n / toti))
ops = {}
for r in body:
    toks = r[src].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    a = ops.setdefault(op, [0, 0]); a[0] += int(r[si]); a[1] += int(r[ii])
print("--- by opcode: samples% / instr%")
for op, (s, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%-12s %6.2f%%  %6.2f%%" % (op, 100.0 * s / tot, 100.0 * n / toti))
