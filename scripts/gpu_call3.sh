#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "integrals or config3_hcn or compacting or octamer or direct" > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -12 gpurun_out/r2_pytest3.log
B="python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline"
run() { name=$1; shift; timeout 300 env "$@" $B > gpurun_out/r2_b3_$name.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_b3_$name.log") if l.startswith("{")][-1])
    print("$name", "E=%.10f"%d["config"]["energy_hartree"], d["config"]["scf_iterations"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "jk GB/s %.0f"%d["roofline"]["jk_kernel"]["achieved"])
except Exception as e: print("$name", "failed", e)
PY
}
run default A=1
run plain DQ_ERI_KERNEL=plain
run default2 A=1
run plain2 DQ_ERI_KERNEL=plain
run sort0 DQ_SORT=0
run sort0plain DQ_SORT=0 DQ_ERI_KERNEL=plain
