#!/bin/bash
# round-2 GPU call 8: profiles of the current code on the dense default workload.
#   1. plain run (must exit 0 before anything runs under ncu)
#   2. launch list of the timed step (gpu__time_duration per launch)
#   3. FP64 instruction counts / pipe / lanes per kernel family
#   4. ncu --set full of one launch each of the gradient kernel, the value kernel and the J/K kernel -> csv pages
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --profile-range"
timeout 300 $CMD > gpurun_out/r2_p8_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_p8_plain.log; exit 1; }
tail -1 gpurun_out/r2_p8_plain.log | cut -c1-300
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_launches.csv --metrics gpu__time_duration.sum $CMD > gpurun_out/r2_p8_l.log 2>&1; echo "launch list rc=$?"
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_fp64_dense.csv -k regex:'k_eri_|k_jk_' --metrics $M $CMD > gpurun_out/r2_p8_m.log 2>&1; echo "metrics rc=$?"
for spec in "grad:k_eri_grad_chunks:1:2" "gdef:k_eri_grad_deferred:1:1" "eri:k_eri_tiles:1:2" "jk:k_jk_stream:3:1"; do
  IFS=: read tag rx skip cnt <<< "$spec"
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o /tmp/r2_$tag $CMD > gpurun_out/r2_p8_$tag.log 2>&1; echo "$tag full rc=$?"
  ncu -i /tmp/r2_$tag.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$tag.raw.csv 2>/dev/null
  ncu -i /tmp/r2_$tag.ncu-rep --page source --csv > gpurun_out/r2_ncu_full_$tag.source.csv 2>/dev/null
  ncu -i /tmp/r2_$tag.ncu-rep --page details --csv > gpurun_out/r2_ncu_full_$tag.details.csv 2>/dev/null
done
du -sh gpurun_out; ls -la gpurun_out | grep r2_ncu_full
