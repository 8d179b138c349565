#!/bin/bash
# A/B of environment switches on the short bench: bash scripts/gpu_env_ab.sh VAR=val [VAR=val ...]   (one run per argument + a default run)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B > gpurun_out/r2_ab_$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_ab_$name.log") if l.startswith("{")][-1])
    print("$name", {k: round(v,2) for k,v in d["phases_ms"].items() if k in ("ms_eri","ms_eri_grad","ms_scf","ms_reverse","ms_jk","ms_total")})
except Exception as e: print("$name", "failed", e)
PY
}
run default A=1
for e in "$@"; do run "$(echo $e | tr '=' '_')" $e; done
