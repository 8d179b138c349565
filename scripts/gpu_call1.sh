#!/bin/bash
# round-2 GPU call 1: full GPU test suite, dense bench, A/B runs, ncu passes on the dense geometry
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -15 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_bench1.log 2>&1; echo "bench rc=$?"
timeout 300 env DQ_ACCUM=float python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1_float.log 2>&1
timeout 300 env DQ_JK_LDG=1 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1_ldg.log 2>&1
timeout 300 env DQ_SORT=0 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1_sort0.log 2>&1
timeout 300 env DQ_SORT=2 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1_sort2.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio
CMD="python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r2_plain.log 2>&1 && \
timeout 900 ncu --clock-control none --csv --log-file gpurun_out/r2_fp64_dense.csv -k regex:'k_eri_(grad_)?(tiles|chunks)|k_jk_stream' -c 40 --metrics $M $CMD > gpurun_out/r2_ncu_m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eri_grad_chunks' -c 7 -o gpurun_out/r2_grad_dense $CMD > gpurun_out/r2_ncu_g.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eri_tiles' -c 7 -o gpurun_out/r2_eri_dense $CMD > gpurun_out/r2_ncu_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_jk_stream' -s 2 -c 2 -o gpurun_out/r2_jk $CMD > gpurun_out/r2_ncu_j.log 2>&1
ls -la gpurun_out | tail -12
cat gpurun_out/r2_bench1.log | tail -3 | cut -c1-3000
