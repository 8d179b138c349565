#!/bin/bash
# round-2 GPU call 5: A/B runs on the dense workload + ncu metric pass restricted to the timed step (--profile-range)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B > gpurun_out/r2_b5_$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_b5_$name.log") if l.startswith("{")][-1])
    print("$name", "E=%.10f"%d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "jk GB/s %.0f"%d["roofline"]["jk_kernel"]["achieved"])
except Exception as e: print("$name", "failed", e)
PY
}
run default A=1
run plain DQ_ERI_KERNEL=plain
run float DQ_ACCUM=float
run ldg DQ_JK_LDG=1
run sort0 DQ_SORT=0
run nozs DQ_NO_ZERO_SKIP=1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
CMD="python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline --profile-range"
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_fp64_dense.csv -k regex:'k_eri_|k_jk_' --metrics $M $CMD > gpurun_out/r2_ncu_m.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_m.log | cut -c1-300
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_fp64_dense_plain.csv -k regex:'k_eri_tiles' --metrics $M env DQ_ERI_KERNEL=plain $CMD > gpurun_out/r2_ncu_m2.log 2>&1
ls -la gpurun_out
