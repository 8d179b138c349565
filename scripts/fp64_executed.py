"""profiles/fp64_executed.json from an ncu metrics pass over the ERI tile kernels and the J/K kernel of one bench step.

On the GPU box (after the same command has exited 0 without ncu):
    ncu --clock-control none --csv --log-file gpurun_out/fp64.csv -k regex:'k_eri_(grad_)?(tiles|tiles_compact|chunks|sparse)|k_jk_(cols|stream)' -c 40 --metrics \
      smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
      smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,\
      smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,\
      dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras [--spacing 24]
Here:  python scripts/fp64_executed.py gpurun_out/fp64.csv <primitive quartets evaluated per pass (dq_stats.n_prim_quartets)> \
           [profiles/fp64_executed_dense.json] [profiles/r2_fp64_executed_dense.csv]
"""
import collections
import csv
import json
import shutil
import sys

src, npq = sys.argv[1], float(sys.argv[2])          # src: one csv, or several joined with ',' (later files add kernel families)
out_json = sys.argv[3] if len(sys.argv) > 3 else "profiles/fp64_executed.json"
out_csv = sys.argv[4] if len(sys.argv) > 4 else "profiles/r1_fp64_executed.csv"
per = collections.OrderedDict()
seen_fams = set()
for fi, one in enumerate(src.split(",")):
    rows = [r for r in csv.reader(l for l in open(one) if l.startswith('"')) if len(r) > 14 and r[0] != "ID"]
    fams_here = set(r[4].split("(")[0].replace("void ", "").split("<")[0] for r in rows)
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        if name.split("<")[0] in seen_fams:
            continue                              # a family counts from the first file that has it
        per.setdefault((fi, r[0], name), {})[r[12]] = float(r[14].replace(",", ""))
    seen_fams |= fams_here
fam = {}
ALIAS = {"k_eri_tiles_def": "k_eri_tiles", "k_eri_grad_deferred": "k_eri_grad_chunks"}      # value pass: k_eri_tiles (L < 2) + k_eri_tiles_def; gradient pass: k_eri_grad_chunks + k_eri_grad_deferred
_UNUSED = {}      # the value pass launches k_eri_tiles (L < 2) and k_eri_tiles_def (L >= 2): one family
for (_, _, name), m in per.items():
    f = name.split("<")[0].replace("dq::", "")
    f = ALIAS.get(f, f)
    a = fam.setdefault(f, {"launches": 0, "flop": 0.0, "inst": 0.0, "ns": 0.0, "pipe_ns": 0.0, "dram": 0.0, "winst": 0.0, "tinst": 0.0})
    wi = m.get("smsp__inst_executed.sum", 0.0)
    a["winst"] += wi
    a["tinst"] += wi * m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0)
    dfma = m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
    dmul = m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
    dadd = m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
    ns = m["gpu__time_duration.sum"]
    a["launches"] += 1
    a["flop"] += 2 * dfma + dmul + dadd
    a["inst"] += dfma + dmul + dadd
    a["ns"] += ns
    a["pipe_ns"] += ns * m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0) / 100.0
    a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
out = {"source": "%s (ncu metrics pass of one bench.py step)" % out_csv,
       "primitive_quartets_evaluated_per_pass": npq}
for f, a in fam.items():
    e = {"launches": a["launches"], "ms_under_ncu": a["ns"] * 1e-6, "executed_flop": a["flop"],
         "executed_fp64_thread_instructions": a["inst"], "pipe_fp64_active_frac": a["pipe_ns"] / a["ns"],
         "executed_tflops_under_ncu": a["flop"] / a["ns"] * 1e-3}
    if a["winst"] > 0:
        e["lanes_per_instruction"] = a["tinst"] / a["winst"]
        e["warp_instructions"] = a["winst"]
    if f in ("k_eri_tiles", "k_eri_tiles_b", "k_eri_tiles_compact", "k_eri_grad_chunks", "k_eri_grad_sparse"):
        e["flop_per_primitive_quartet"] = a["flop"] / npq
        e["fp64_instructions_per_primitive_quartet"] = a["inst"] / npq
        e["dram_bytes_per_step"] = a["dram"]
    else:
        e["dram_bytes_per_launch"] = a["dram"] / a["launches"]
    out[f] = e
json.dump(out, open(out_json, "w"), indent=1)
with open(out_csv, "w") as f:
    for one in src.split(","):
        f.write(open(one).read())
print(json.dumps(out, indent=1))
