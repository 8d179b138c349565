#!/bin/bash
# round-2 GPU call 7: chunk-size A/B of the gradient pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B > gpurun_out/r2_b7_$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_b7_$name.log") if l.startswith("{")][-1])
    print("$name", "E=%.10f"%d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "jk GB/s %.0f"%d["roofline"]["jk_kernel"]["achieved"])
except Exception as e: print("$name", "failed", e)
PY
}
run default A=1
run chunk6 DQ_GRAD_CHUNK=6
run chunk24 DQ_GRAD_CHUNK=24
run chunk48 DQ_GRAD_CHUNK=48
run chunk96 DQ_GRAD_CHUNK=96
run default2 A=1
