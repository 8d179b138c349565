#!/bin/bash
# round-2 GPU call 9: quick suite + bench after a kernel change
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest9.log
B="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B > gpurun_out/r2_b9_$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_b9_$name.log") if l.startswith("{")][-1])
    print("$name", "E=%.10f"%d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "jk GB/s %.0f"%d["roofline"]["jk_kernel"]["achieved"])
except Exception as e: print("$name", "failed", e)
PY
}
run default A=1
run default2 A=1
for e in "$@"; do run "$(echo $e | tr '=' '_')" $e; done
