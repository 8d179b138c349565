#!/bin/bash
# multi-GPU check: bench under torchrun at N = $1 (ours + the reference arm), and the 2-GPU NCCL test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_scale_n$N.log 2>&1; echo "ours N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_scale_n$N.log") if l.startswith("{")][-1])
    print("N=$N value %.4g e2e %.4g ms/step %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v,2) for k,v in d["phases_ms"].items()})
    print("  jk", d["roofline"]["jk_kernel"]["achieved"], "config5", d.get("config5",{}).get("ms"), "direct", d.get("integral_direct",{}).get("ms_per_step"), "sparse24", d.get("sparse_24bohr_core_guess",{}).get("ms_per_step"))
except Exception as e: print("failed", e)
PY
tail -3 gpurun_out/r2_scale_n$N.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2_scale_ref_n$N.log 2>&1; echo "reference N=$N rc=$?"
grep '^{' gpurun_out/r2_scale_ref_n$N.log | cut -c1-400
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu or two_rank" > gpurun_out/r2_pytest_multi.log 2>&1; tail -2 gpurun_out/r2_pytest_multi.log; fi
