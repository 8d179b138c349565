#!/bin/bash
# round-2 GPU call 2: failing tests verbosely, A/B benches, ncu passes on the dense geometry (CSV pages only come back)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "h2o_mixed or hcn or full_size or octamer or identical_bits or error_statuses or config3 or nuclear or hessian" > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
B="python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2_b2_default.log 2>&1
timeout 300 env DQ_ACCUM=float $B > gpurun_out/r2_b2_float.log 2>&1
timeout 300 env DQ_JK_LDG=1 $B > gpurun_out/r2_b2_ldg.log 2>&1
timeout 300 env DQ_SORT=0 $B > gpurun_out/r2_b2_sort0.log 2>&1
timeout 300 env DQ_SORT=2 $B > gpurun_out/r2_b2_sort2.log 2>&1
for f in default float ldg sort0 sort2; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_b2_$f.log") if l.startswith("{")][-1])
    print("$f", "E=%.10f"%d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "jk GB/s %.0f"%d["roofline"]["jk_kernel"]["achieved"])
except Exception as e: print("$f", "failed", e)
PY
done
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
CMD="python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r2_plain.log 2>&1 && \
timeout 900 ncu --clock-control none --csv --log-file gpurun_out/r2_fp64_dense.csv -k regex:'k_eri_(grad_)?(tiles|chunks)|k_jk_stream' -c 40 --metrics $M $CMD > gpurun_out/r2_ncu_m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eri_grad_chunks' -s 1 -c 2 -o /tmp/r2_grad_dense $CMD > gpurun_out/r2_ncu_g.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eri_tiles' -s 1 -c 2 -o /tmp/r2_eri_dense $CMD > gpurun_out/r2_ncu_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_jk_stream' -s 2 -c 1 -o /tmp/r2_jk $CMD > gpurun_out/r2_ncu_j.log 2>&1
for n in r2_grad_dense r2_eri_dense r2_jk; do
  ncu -i /tmp/$n.ncu-rep --page raw --csv > gpurun_out/$n.raw.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page source --csv > gpurun_out/$n.source.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page details --csv > gpurun_out/$n.details.csv 2>/dev/null
  ls -la /tmp/$n.ncu-rep
done
du -sh gpurun_out; ls -la gpurun_out | tail -25
