#!/usr/bin/env python
"""profiles/<prefix>_stream_ncu.md from the `ncu --set full` capture of tools/micro_once.py (composite fwd / bwd at
S = 64, 192, 512 and sample_pdf 63 / 128): measured DRAM bytes against the algorithmic bytes of SURVEY.md 8d.

    python tools/ncu_stream_md.py gpurun_out/r2_stream.ncu-rep profiles/r2 "<title>"
"""
import csv, subprocess, sys
rep, prefix, title = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units = rows[0], dict(zip(rows[0], rows[1]))


def val(d, k):
    v = float(d[k].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units.get(k, ""), 1.0)


# launch order of tools/micro_once.py: per S in (64, 192, 512): fwd x2, bwd x2; then sample_pdf x2.  Launch 1 of each pair is kept.
plan = []
for S in (64, 192, 512):
    n = (1 << 25) // S
    plan += [("composite fwd", S, n, n * (S * 80 + 92)), None, ("composite bwd", S, n, n * (S * 152 + 108)), None]
plan += [("sample_pdf 63 bins / 128 samples", 64, 1 << 20, (1 << 20) * 4 * (63 + 62 + 256)), None]
# the capture holds launches in order: fwd, fwd, bwd, bwd, ... ; keep the SECOND of each pair
recs = [dict(zip(head, r)) for r in rows[2:]]
lines = []
i = 0
for S in (64, 192, 512):
    n = (1 << 25) // S
    for name, alg in (("composite fwd", n * (S * 80 + 92)), ("composite bwd", n * (S * 152 + 108))):
        d = recs[i + 1]; i += 2
        lines.append((name, S, n, alg, d))
d = recs[i + 1]
lines.append(("sample_pdf (63 bins, 128 samples)", 64, 1 << 20, (1 << 20) * 4 * (63 + 62 + 256), d))
stall = lambda d, k: float(d.get("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k, "0").replace(",", ""))
with open(prefix + "_stream_ncu.md", "w") as f:
    f.write("# ncu --set full, stand-alone streaming kernels, %s\n\n" % title)
    f.write("`ncu --set full --clock-control none --import-source on -k regex:'composite_(fwd|bwd)|sample_pdf' python tools/micro_once.py` "
            "(second launch of every kernel; 32 Mi samples per compositing slab, 1 Mi rays for sample_pdf).  Algorithmic bytes: SURVEY.md 8d "
            "(80 / 152 B per sample + per-ray terms; 1 524 B per ray).  Times under ncu are serialised single launches; the CUDA-event "
            "figures of the same kernels are in the `micro` record of the bench lines.\n\n")
    f.write("| kernel | S | rays | ms | dram read | dram write | algorithmic | traffic / algorithmic | DRAM GB/s | kernel | l1tex % | issue % | warps active % | regs | dominant stalls (warps per issue) |\n")
    f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|---|\n")
    for name, S, n, alg, d in lines:
        t = val(d, "gpu__time_duration.sum")
        rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
        st = sorted(((stall(d, k), k) for k in ("long_scoreboard", "short_scoreboard", "mio_throttle", "math_pipe_throttle", "wait", "not_selected", "lg_throttle", "barrier")), reverse=True)[:3]
        f.write("| %s | %d | %d | %.4f | %.3f GB | %.3f GB | %.3f GB | %.2f | %.0f | `%s` | %.1f | %.1f | %.1f | %s | %s |\n" % (
            name, S, n, t * 1e3, rd / 1e9, wr / 1e9, alg / 1e9, (rd + wr) / alg, (rd + wr) / t / 1e9,
            d["Kernel Name"].split("(")[0].replace("void ", "").replace("ibln::", "")[:44],
            float(d["l1tex__throughput.avg.pct_of_peak_sustained_active"]), float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]), d["launch__registers_per_thread"],
            ", ".join("%s %.1f" % (k, v) for v, k in st)))
print(open(prefix + "_stream_ncu.md").read())
