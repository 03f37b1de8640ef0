#!/usr/bin/env python
"""profiles/*.md from gpurun_out captures:  ncu_report_md.py <full .ncu-rep> <launch-list csv> <out prefix> <title>"""
import csv, subprocess, sys, collections
rep, launches, prefix, title = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head = rows[0]
cols = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tc active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__cluster_size", "cluster")]
units = dict(zip(head, rows[1]))
with open(prefix + "_ncu_full.md", "w") as f:
    f.write("# ncu --set full, MLP kernels of one training step, %s\n\n" % title)
    f.write("`ncu --set full --clock-control none --import-source on -k regex:mlp_(fwd|dgrad|wgrad)_kernel -s 30 -c 10 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --skip phases,strong,render,micro`\n\n")
    f.write("The ten MLP launches of ONE training step (N_rand = 4096): coarse main (stash), coarse eps-normal (sigma-only), coarse reflected, "
            "fine main (stash), fine eps-normal, fine reflected, then the backward of the fine and of the coarse network (dgrad, wgrad). "
            "Per-launch times under ncu are cold-cache and serialised.\n\n")
    f.write("| # | kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|---|" + "---:|" * len(cols) + "\n")
    for i, row in enumerate(rows[2:]):
        d = dict(zip(head, row))
        name = d["Kernel Name"].replace("ibln::mlp::", "").split("(")[0]
        vals = []
        for k, _ in cols:
            v = d.get(k, "")
            u = units.get(k, "")
            vals.append(("%s %s" % (v, u)).strip() if u in ("Gbyte", "Mbyte", "Kbyte", "byte") else v)
        f.write("| %d | `%s` | " % (i, name) + " | ".join(vals) + " |\n")
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, mi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) > mi:
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1; a[1] += float(r[mi].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
with open(prefix + "_launches_summary.md", "w") as f:
    f.write("# ncu launch list, %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --skip phases,strong,render,micro`\n\n" % title)
    f.write("All launches of the run (3 warm-up + 1 timed + 2+1 end-to-end + 2 breakdown steps = 9 training steps); per-launch times are cold-cache and serialised, so compare SHARES.\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:90], n, ms, 100 * ms / tot))
    mlp = sum(ms for k, (n, ms) in agg.items() if "mlp_fwd_kernel" in k)
    f.write("\nTotal %.2f ms over %d launches. `mlp_fwd_kernel` (3 instantiations) share: %.1f%%.\n" % (tot, sum(a[0] for a in agg.values()), 100 * mlp / tot))
# per-launch DRAM traffic of the dominant kernel (bench.py reads this file for roofline.traffic)
import json
def _bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
out2 = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out2.splitlines()))
hh, uu = rr[0], dict(zip(rr[0], rr[1]))
vals = []
for row in rr[2:]:
    d = dict(zip(hh, row))
    if "mlp_fwd_kernel" in d["Kernel Name"]:
        vals.append(_bytes(d["dram__bytes_read.sum"], uu["dram__bytes_read.sum"]) + _bytes(d["dram__bytes_write.sum"], uu["dram__bytes_write.sum"]))
json.dump({"kernel": "mlp_fwd_kernel", "launches": len(vals), "dram_bytes_per_launch": vals, "dram_bytes_per_launch_mean": sum(vals) / max(len(vals), 1),
           "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, the %d mlp_fwd launches of one training step; %s" % (len(vals), title)},
          open(prefix + "_mlp_fwd_traffic.json", "w"), indent=1)
