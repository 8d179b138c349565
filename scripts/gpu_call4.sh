#!/bin/bash
# round-2 GPU call 4 (re-entry): state of the suite and the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.log 2>&1
nproc >> gpurun_out/r2_smi.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 ) > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -40 gpurun_out/r2_pytest4.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench4.log 2>&1; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench4.log | cut -c1-6000
