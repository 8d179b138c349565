#!/bin/bash
# final validation as the driver runs it: smoke, the GPU suite, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_final_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r2_final_bench.log 2>&1; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2_final_ref.log 2>&1; echo "reference rc=$?"
grep '^{' gpurun_out/r2_final_ref.log | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_final_bench.log") if l.startswith("{")][-1])
print("value %.4g e2e %.4g ms/step %.1f launches %d clocks %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
print({k: round(v,2) for k,v in d["phases_ms"].items()})
r=d["roofline"]; print(r["kernel"], "achieved %.2f frac %.3f pipe %s lanes %s contract frac %.2f"%(r["achieved"], r["frac"], r.get("fp64_pipe_active_ncu"), r.get("active_lanes_per_warp_instruction_ncu"), r["contract_count"]["frac"]))
o=r["other_tile_kernel"]; print(o["kernel"], "achieved %.2f frac %.3f"%(o["achieved"], o["frac"]))
j=r["jk_kernel"]; print("jk %.0f GB/s frac %.3f ms/build %.3f traffic %s"%(j["achieved"], j["frac"], j["ms_per_build"], j["traffic"]))
print("direct %.0f ms (jk/build %.0f) ; sparse24 %.1f ms ; config5 %.2f ms dE %.1e ; cpu %s"%(d["integral_direct"]["ms_per_step"], d["integral_direct"]["ms_jk_per_build"], d["sparse_24bohr_core_guess"]["ms_per_step"], d["config5"]["ms"], d["config5"]["max_abs_dE_vs_oracle_on_8_molecules"], d["cpu_baseline"]["value"]))
print("E", d["config"]["energy_hartree"], d["config"]["scf_iterations"], d["config"]["converged"])
PY
tail -3 gpurun_out/r2_final_bench.log | grep real
