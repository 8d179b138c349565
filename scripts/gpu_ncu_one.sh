#!/bin/bash
# ncu --set full of one kernel: bash scripts/gpu_ncu_one.sh <kernel regex> <skip> <tag> [ENV=val ...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rx=$1; skip=$2; tag=$3; shift 3
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --profile-range"
timeout 900 env "$@" ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/one_$tag $CMD > gpurun_out/r2_one_$tag.log 2>&1; echo "rc=$?"
ncu -i /tmp/one_$tag.ncu-rep --page raw --csv > gpurun_out/r2_one_$tag.raw.csv 2>/dev/null
ncu -i /tmp/one_$tag.ncu-rep --page source --csv > gpurun_out/r2_one_$tag.source.csv 2>/dev/null
ls -la gpurun_out/r2_one_$tag.*
