#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( DQ_GRAD_DEFER=1 timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/r2_pytest_defer.log 2>&1; echo "pytest(defer) rc=$?"; tail -12 gpurun_out/r2_pytest_defer.log | cut -c1-200
bash scripts/gpu_env_ab.sh DQ_GRAD_DEFER=1
