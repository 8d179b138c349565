#!/usr/bin/env python
"""bench.py — RHF energy + full basis-parameter gradient of (H2O)_32 / STO-3G on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--waters 32]

One STEP = one complete pass of the hot path over the synthetic cluster: normalisation, S/T/V, the ERI tile pass,
the Roothaan SCF from the core guess to |dE| < 1e-8, its exact reverse sweep, the ERI-gradient tile pass and the
gradient assembly -> E and dE/d(ln alpha), dE/d(coef), dE/d(centres) (2016 numbers).  Metric (BASELINE.json):
ERI+gradient primitive quartets per second = primitive quartets of the system / step time.

  value : device time only (CUDA events inside the library, inputs already uploaded)
  e2e   : wall clock through the public API (Energy.rhf_grad -> C ABI) with HOST buffers: host planning, H2D of the
          system arrays and D2H of E + gradient inside the timed region
  --impl reference : the CPU port of the reference's algorithm (oracle/rhf_oracle.cpp, forward-mode gradient exactly as
          Task.py:25-51 does with algopy) on all host threads, on a bounded sample (see cpu_baseline.sample)
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "eri_plus_gradient_primitive_quartets_per_s"
UNIT = "primitive quartets/s"
FLOP_VALUE, FLOP_GRAD = 375.0, 1500.0   # algorithmic flop per primitive quartet
REF_SAMPLE_WATERS = 3                   # --impl reference: (H2O)_3 per step, a few seconds on the host cores
CPU_BASELINE_WATERS = 4                 # cpu_baseline of the default run: (H2O)_4, ~10-30 s of CPU work


def executed_counts():
    """ncu-counted FP64 work of the two tile kernels on THIS workload (profiles/fp64_executed.json, written by
    scripts/fp64_executed.py from an ncu metrics pass of this same command): executed flop per primitive quartet and the
    DRAM traffic per launch set.  None when the file is absent (the roofline then falls back to the algorithmic count)."""
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_executed.json")) as f:
            return json.load(f)
    except Exception:
        return None


def expanded_counts():
    """The same counts measured before any exact-zero skipping existed (every primitive quartet expanded): the kernels'
    dense-work efficiency, for comparison with the lane-divergent zero-skipping run."""
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_executed_expanded.json")) as f:
            return json.load(f)
    except Exception:
        return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gbps", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "MEASURED_PEAKS.json %s (of measured)" % k
    except Exception:
        pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent on this box)"


SPACING = 24.0      # bohr; --spacing overrides (kernel studies on denser, non-converging clusters)


def workload(n_waters):
    from diffiqult_b200 import System_mol, synthetic
    from diffiqult_b200.Basis import basis_set_3G_STO
    mol = synthetic.water_cluster(n_waters, spacing=SPACING)
    return System_mol(mol, basis_set_3G_STO, 10 * n_waters, "w%d" % n_waters)


def prim_quartets(s):
    k = np.asarray(s.list_contr, dtype=np.int64)
    kk = np.outer(k, k)[np.triu_indices(len(k))]          # primitives per unique function pair
    tot = int(kk.sum())
    return (tot * tot + int((kk * kk).sum())) // 2       # sum over unique pair-of-pairs of KK_x * KK_y


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, read in-process through NVML (the data source of
    nvidia-smi; forking nvidia-smi every 200 ms from a process that owns a CUDA context stalls its launches)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_sm, self.stop_flag = index, [], set(), None, False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: say so instead of inventing numbers
            self.reasons.add("unavailable: %s" % type(e).__name__)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "NVML in-process, 100 ms period"}


def cpu_reference_step(sample_waters, threads=None):
    """One bounded step of the CPU port: energy + forward-mode gradient (argnum 0,1,2) of (H2O)_sample."""
    from oracle import cpu_oracle as orc
    if threads:
        orc.lib().orc_set_num_threads(C.c_int(threads))
    s = workload(sample_waters)
    a = orc.SysArrays(s)
    t0 = time.perf_counter()
    e = orc.rhf(a, max_scf=100)
    for n in (0, 1, 2):
        g = orc.rhf_grad(a, n, max_scf=100)
        assert g["status"] == 0
    dt = time.perf_counter() - t0
    return prim_quartets(s) / dt, dt, e["E"], orc.lib().orc_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_waters
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, e, cores = cpu_reference_step(sample)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals])) * 1e3
    s_full = workload(args.waters)
    sample_txt = ("(H2O)_%d / STO-3G cut from the same cluster generator: RHF energy + forward-mode gradient for argnum 0,1,2 "
                  "(the reference's algorithm, Task.py:25-51), C++/OpenMP port oracle/rhf_oracle.cpp; the reference itself is "
                  "single-threaded Python 2 + algopy and cannot run (H2O)_%d at all" % (sample, args.waters))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "(H2O)_%d STO-3G RHF energy + full parameter gradient (argnum 0,1,2)" % args.waters,
                       "n_basis": int(s_full.nbasis), "n_primitives": int(len(s_full.alpha)),
                       "primitive_quartets": prim_quartets(s_full)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffiqult_b200 import Energy, _capi, distributed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: diffiqult_b200 has no CPU path")
    torch.cuda.set_device(local)
    allreduce = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        allreduce = distributed.make_allreduce()
    s = workload(args.waters)
    nq = prim_quartets(s)
    lnalpha = np.log(s.alpha)
    stream = distributed.current_stream_handle()
    kw = dict(max_scf=args.max_scf, device=local, stream=stream, world_size=world, rank=rank, allreduce=allreduce, jk_mode=args.jk_mode,
              screen_threshold=args.screen)

    def step():
        return Energy.rhf_grad(lnalpha, s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak = C.c_double(0.0)
    _capi.check(_capi.lib().dq_fp64_peak(C.c_int32(local), C.byref(peak)))
    for _ in range(args.warmup):
        r = step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    stats = []
    for _ in range(args.steps):
        r = step()
        stats.append(r["stats"])
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    dev_ms = float(np.sum([st["ms_total"] - st["ms_setup"] for st in stats]))
    red = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    wall, dev_ms = float(red[0]), float(red[1])
    # For the record, outside the timed region: the same step with EVERY primitive quartet expanded - no skipping of
    # underflowed Gaussian products, dense-weight gradient kernel (the library reads these switches on every call).  Same
    # energy and gradient; this is what the default run saves by not multiplying by exact zeros.
    expanded = None
    if not args.no_expanded and args.screen <= 0:
        os.environ["DQ_NO_ZERO_SKIP"] = "1"
        os.environ["DQ_GRAD_KERNEL"] = "dense"
        try:
            step()
            barrier()
            t1 = time.perf_counter()
            rx = step()
            barrier()
            ex_wall = time.perf_counter() - t1
        finally:
            del os.environ["DQ_NO_ZERO_SKIP"], os.environ["DQ_GRAD_KERNEL"]
        redx = torch.tensor([ex_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(redx, op=dist.ReduceOp.MAX)
        expanded = {"ms_per_step": float(redx[0]) * 1e3, "energy_hartree": rx["E"],
                    "max_abs_gradient_difference_to_default_run": float(max(np.abs(rx["grad"][n] - r["grad"][n]).max() for n in range(3))),
                    "phases_ms": {k: rx["stats"][k] for k in ("ms_eri", "ms_scf", "ms_reverse", "ms_eri_grad")},
                    "what": "one step with DQ_NO_ZERO_SKIP=1 DQ_GRAD_KERNEL=dense (every primitive quartet of both passes expanded "
                            "wherever its weight is not exactly zero), measured after the timed region"}
    if rank == 0:
        st = stats[-1]
        K = args.steps
        ms_grad_kernel = float(np.mean([x["ms_eri_grad"] for x in stats]))
        ms_eri_kernel = float(np.mean([x["ms_eri"] for x in stats]))
        builds = max(st["fock_builds"], 1)
        ms_jk_build = float(np.mean([x["ms_jk"] for x in stats])) / builds
        npq = st["n_prim_quartets"]                      # primitive quartets this rank evaluates per ERI pass
        alg_grad = npq * FLOP_GRAD / (ms_grad_kernel * 1e-3) * 1e-12
        alg_eri = npq * FLOP_VALUE / (ms_eri_kernel * 1e-3) * 1e-12
        # the ncu counts are data dependent: they hold for the default workload only
        ex = executed_counts() if (args.spacing == 24.0 and args.waters == 32 and args.screen <= 0) else None
        gk = "k_eri_grad_sparse" if st["grad_kernel_sparse"] else "k_eri_grad_chunks"
        vk = "k_eri_tiles_compact" if st["eri_kernel_compact"] else "k_eri_tiles"
        if ex and (gk not in ex or vk not in ex):
            ex = None
        if ex:
            achieved = npq * ex[gk]["flop_per_primitive_quartet"] / (ms_grad_kernel * 1e-3) * 1e-12
            ach_eri = npq * ex[vk]["flop_per_primitive_quartet"] / (ms_eri_kernel * 1e-3) * 1e-12
            basis = ("EXECUTED FP64 flop per primitive quartet on this workload (ncu smsp__sass_thread_inst_executed_op_"
                     "{2*dfma,dmul,dadd}_pred_on summed over the 7 class launches, %s) x primitive quartets / live CUDA-event "
                     "time of those launches" % ex.get("source", "profiles/fp64_executed.json"))
        else:
            achieved, ach_eri = alg_grad, alg_eri
            basis = "ALGORITHMIC flop (SURVEY.md 8d) - profiles/fp64_executed.json absent"
        hbm_peak, hbm_src = measured_hbm_peak()
        jk_bytes = st["n_tiles"] * 2048.0                # 8 B per stored contracted ERI, every tile read once per build
        jk_gbs = jk_bytes / (ms_jk_build * 1e-3) * 1e-9 if ms_jk_build > 0 else 0.0
        step_ms = dev_ms / K
        peak_src = ("measured live: dq_fp64_peak FMA loop (MEASURED_PEAKS.json has no FP64 entry); the loop's FMAs have one "
                    "register operand - a DFMA with three distinct register operands, the normal case in the integral "
                    "kernels, issues at 2/3 of this rate on B200 (scripts/microbench/fp64_pipe.cu, DESIGN.md section 3)")
        alg_note = ("SURVEY.md 8(d) count of the reference's formulas (2.7 Boys calls x ~100 flop dominate); this geometry is "
                    "almost entirely in the t >= 50 regime where the reference's continued fraction equals its closed form to the "
                    "last bit, so far fewer flop are executed: frac > 1 here is not a hardware statement, `frac` is")

        def tile_kernel(name, ms, ach, alg, alg_flop, key):
            return {"kernel": "%s<*> (7 class launches per step, %.0f %% of the step)" % (name, 100 * ms / step_ms),
                    "bound": "fp64", "bound_note": "FP64 CUDA-core pipe: neither HBM nor the tensor cores bound this kernel",
                    "ms_per_step": ms, "achieved": ach, "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value,
                    "traffic": ex[key].get("dram_bytes_per_step") if ex else None, "achieved_basis": basis, "peak_source": peak_src,
                    "fp64_pipe_active_ncu": ex[key].get("pipe_fp64_active_frac") if ex else None,
                    "algorithmic": {"flop_per_primitive_quartet": alg_flop, "achieved": alg, "frac": alg / peak.value, "note": alg_note}}
        exd = expanded_counts()

        def dense_ref(key):
            if not exd or key not in exd:
                return None
            e = exd[key]
            return {"executed_tflops": e["executed_tflops_under_ncu"], "frac": e["executed_tflops_under_ncu"] / peak.value,
                    "fp64_pipe_active_ncu": e["pipe_fp64_active_frac"], "ms": e["ms_under_ncu"],
                    "what": "the same kernel family with EVERY primitive quartet expanded (before exact-zero skipping; "
                            "profiles/fp64_executed_expanded.json): its dense-work efficiency.  With zero skipping the step is "
                            "faster but lanes whose primitive pair vanished idle next to lanes whose pair did not, so the "
                            "executed-flop rate and `frac` above are lower than this"}
        k_val = tile_kernel(vk, ms_eri_kernel, ach_eri, alg_eri, FLOP_VALUE, vk)
        k_val["all_quartets_expanded"] = dense_ref("k_eri_tiles")
        k_grad = tile_kernel(gk, ms_grad_kernel, achieved, alg_grad, FLOP_GRAD, gk)
        k_grad["all_quartets_expanded"] = dense_ref("k_eri_grad_chunks")
        # the dominant kernel family of the step leads; the other tile kernel and the HBM-bound J/K kernel follow
        roofline = dict(k_val if ms_eri_kernel >= ms_grad_kernel else k_grad)
        roofline["other_tile_kernel"] = k_grad if ms_eri_kernel >= ms_grad_kernel else k_val
        roofline["jk_kernel"] = {"kernel": "k_jk_cols", "bound": "hbm", "achieved": jk_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": jk_gbs / hbm_peak, "peak_source": hbm_src, "ms_per_build": ms_jk_build,
                                 "builds_per_step": builds, "algorithmic_bytes_per_build": jk_bytes,
                                 "traffic": ex.get("k_jk_cols", {}).get("dram_bytes_per_launch") if ex else None}
        h2d = int(lnalpha.nbytes + np.asarray(s.coef).nbytes + s.xyz.nbytes + 4 * (2 * s.nbasis + s.natoms) + s.atom.nbytes)
        d2h = int(8 * (1 + 2 * len(s.alpha) + 3 * s.nbasis) + 8 * (s.nbasis * s.nbasis + s.nbasis))
        line = {
            "metric": METRIC, "value": nq / (dev_ms / K * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": wall / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "(H2O)_%d STO-3G RHF energy + full parameter gradient (argnum 0,1,2)" % args.waters,
                       "grid_spacing_bohr": args.spacing,
                       "n_basis": int(s.nbasis), "n_primitives": int(len(s.alpha)), "primitive_quartets": nq,
                       "scf_iterations": st["scf_iterations"], "converged": r["status"] == 0, "energy_hartree": r["E"],
                       "primitive_pairs_underflowed": "%.1f %% of the primitive-pair prefactors exp(-ab|AB|^2/(a+b)) underflow to exactly 0 "
                       "in double precision (monomers 24+ bohr apart); primitive quartets containing one are exact zeros, value "
                       "and derivatives, and are not expanded (bit-identical results; DQ_NO_ZERO_SKIP=1 expands them)" % (
                           100.0 * st["n_prim_pairs_underflowed"] / max(st["n_prim_pairs"], 1)),
                       "gradient_quartets_with_exactly_zero_weight": "%.1f %% of this rank's contracted quartets: D and the adjoint "
                       "Fock matrices are exactly zero between uncoupled far-apart monomers (the eigensolver keeps exact zeros), so "
                       "these quartets add exactly 0 to the gradient and their primitive loops are skipped; the ERI pass itself "
                       "evaluates everything" % (100.0 * st["n_grad_quartets_zero_weight"] / max(st["n_quartets"], 1)),
                       "jk_mode": ("integral-direct (tiles recomputed in every Fock build)" if st["jk_direct"] else
                                   "stored ERI tiles (%.2f GB per rank)" % (st["n_tiles"] * 2048 / 1e9)),
                       "l2": "ERI tile store larger than L2; every step recomputes everything from host inputs",
                       "screening": ("off: all %d primitive quartets evaluated, as in the reference" % nq) if args.screen <= 0 else
                                    ("opt-in threshold %g: %d of %d tiles skipped" % (args.screen, st["n_tiles_skipped"], st["n_tiles"])),
                       "parallelism": "tiles round-robin over %d ranks; allreduce of J|K per Fock build and of the gradient" % world},
            "e2e": {"value": nq / (wall / K), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(np.sum([x["launches"] for x in stats])),
            "all_quartets_expanded": expanded,
            "clocks": sampler.summary(),
            "roofline": roofline,
            "phases_ms": {k: float(np.mean([x[k] for x in stats])) for k in
                          ("ms_setup", "ms_one_electron", "ms_eri", "ms_scf", "ms_reverse", "ms_eri_grad", "ms_grad_other", "ms_jk", "ms_total")},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, e, cores = cpu_reference_step(CPU_BASELINE_WATERS)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "(H2O)_%d STO-3G (same cluster generator) energy + forward-mode gradient argnum 0,1,2: "
                                              "C++/OpenMP port of the reference algorithm (oracle/rhf_oracle.cpp), %.1f s of CPU "
                                              "work on %d threads" % (CPU_BASELINE_WATERS, dt, cores)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--waters", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-expanded", action="store_true", help="skip the extra all-quartets-expanded step after the timed region")
    ap.add_argument("--ref-waters", type=int, default=REF_SAMPLE_WATERS, help="--impl reference: monomers in the bounded CPU sample")
    ap.add_argument("--jk-mode", default="auto", choices=["auto", "stored", "direct"])
    ap.add_argument("--spacing", type=float, default=24.0, help="grid spacing of the water cluster in bohr (default 24: the "
                    "densest grid on which the reference's plain Roothaan iteration converges; smaller values are kernel studies)")
    ap.add_argument("--max-scf", type=int, default=100)
    ap.add_argument("--screen", type=float, default=0.0, help="opt-in tile screening threshold (default 0: evaluate everything, as the reference)")
    args = ap.parse_args()
    global SPACING
    SPACING = args.spacing
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
