#!/bin/bash
# round-2 GPU call 6: suite + default bench with the private-strip J/K kernel, plain value kernel, per-class asymptotic thresholds
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 ) > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -8 gpurun_out/r2_pytest6.log
B="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B > gpurun_out/r2_b6_$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_b6_$name.log") if l.startswith("{")][-1])
    print("$name", "E=%.10f"%d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "jk GB/s %.0f"%d["roofline"]["jk_kernel"]["achieved"])
except Exception as e: print("$name", "failed", e)
PY
}
run default A=1
run jkshared DQ_JK_SHARED=1
run float DQ_ACCUM=float
( time timeout 900 python bench.py ) > gpurun_out/r2_bench6.log 2>&1; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench6.log | cut -c1-1500
