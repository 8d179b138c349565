"""Build profiles/r1_summary.md from the ncu launch list and the full-set reports in gpurun_out/."""
import csv, collections, re, subprocess, sys, os
out = []
lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    v = float(row["Metric Value"].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out.append("# Round 1 profile summary (B200, `python bench.py --steps 1 --warmup 1 --no-cpu-baseline`)\n")
out.append("Source: `profiles/r1_launches.csv` (ncu `--metrics gpu__time_duration.sum --clock-control none`; 2 steps = warm-up + 1).\n"
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n")
out.append("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
fam = collections.OrderedDict()
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot > 0.002:
        out.append("| `%s` | %d | %.2f | %.1f %% | %.3f |" % (k, n, t, 100 * t / tot, t / n))
    f = re.sub(r"<.*", "", k); fam[f] = fam.get(f, 0) + t
out.append("\nBy kernel family: " + ", ".join("`%s` %.1f %%" % (k, 100 * v / tot) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:8]) + "\n")
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
out.append("## Full-set captures (`ncu --set full --clock-control none --import-source on`)\n")
for tag in ("grad", "eri", "scf"):
    rep = "gpurun_out/prof_r1_%s.ncu-rep" % tag
    if not os.path.isfile(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open("profiles/r1_ncu_full_%s.raw.csv" % tag, "w").write(raw)
    r = list(csv.reader(raw.split("\n")))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        if len(row) < len(hdr):
            continue
        name = re.sub(r"\(const.*|\(int.*", "", row[hdr.index("Kernel Name")]).replace("void ", "")
        out.append("### `%s`\n" % name)
        vals = {}
        for w in want:
            if w in hdr:
                vals[w] = (row[hdr.index(w)], units[hdr.index(w)])
                out.append("* %s = %s %s" % (w, row[hdr.index(w)], units[hdr.index(w)]))
        try:
            fl = (2 * float(vals[want[9]][0].replace(",", "")) + float(vals[want[10]][0].replace(",", "")) + float(vals[want[11]][0].replace(",", "")))
            ms = float(vals[want[0]][0].replace(",", ""))
            u = vals[want[0]][1]
            sec = ms * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0, "s": 1.0}[u]
            out.append("* executed FP64 flop (2*DFMA + DMUL + DADD) = %.3e -> %.2f TFLOP/s executed" % (fl, fl / sec * 1e-12))
        except Exception as e:
            pass
        out.append("")
open("profiles/r1_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])
