#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --max-scf 8 --core-guess"
for i in 1 2; do timeout 300 $B > gpurun_out/r2_exp_$i.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_exp_$i.log") if l.startswith("{")][-1])
print("run $i", d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items() if k in ("ms_eri","ms_eri_grad","ms_scf","ms_jk","ms_total")})
PY
done
