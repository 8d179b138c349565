#!/bin/bash
# stall-reason metrics (per issue) of the tile kernels on the dense default workload -> gpurun_out/r2_stalls_$1.csv
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-x}; shift
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --profile-range"
S=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread
timeout 900 env "$@" ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_stalls_$tag.csv -k regex:'k_eri_' --metrics $S $CMD > gpurun_out/r2_stalls_$tag.log 2>&1; echo "rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r2_stalls_$tag.csv") if l.startswith('"')) if len(r)>14 and r[0]!="ID"]
per=collections.OrderedDict()
for r in rows: per.setdefault((r[0], r[4].split("(")[0].replace("void ","")),{})[r[12]]=float(r[14].replace(",",""))
for (i,n),m in per.items():
    s={k.split("stalled_")[1].split("_per")[0]:v for k,v in m.items() if "stalled" in k}
    print("%-34s %7.2f ms regs %3d lanes %4.1f issue %4.1f pipe %4.1f | "%(n[:34], m["gpu__time_duration.sum"]*1e-6, m["launch__registers_per_thread"], m["smsp__thread_inst_executed_per_inst_executed.ratio"], m["smsp__issue_active.avg.pct_of_peak_sustained_active"], m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]) + " ".join("%s %.2f"%(k[:9],v) for k,v in sorted(s.items(), key=lambda x:-x[1])[:6]))
PY
