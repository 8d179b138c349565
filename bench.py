#!/usr/bin/env python
"""bench.py — RHF energy + full basis-parameter gradient of (H2O)_32 / STO-3G on N B200s (BASELINE config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--waters 32] [--spacing 5.67]

Workload: SURVEY.md 8(d) — 32 rotated, jittered water monomers on a 4x4x2 grid with 5.67 bohr (3.0 A) spacing,
seed 20260101, N = 224 basis functions, 672 primitives, 2.57e10 unique primitive quartets, 2016 gradient components.
The reference's plain Roothaan iteration does not converge this cluster from the core guess; it does from the
reference's own `readguess` (Energy.py:148-157) given the superposition of the 32 monomer SCFs (8 iterations,
E = -2398.52725184 Eh, the CPU oracle's numbers).  The guess matrix is an INPUT: it is built once before the timed region
(32 small SCFs on the GPU, `guess_ms`) and handed to every step as a host array, like the reference's .npy file.

One STEP = one complete pass of the hot path: normalisation, S/T/V, the ERI tile pass, the SCF to |dE| < 1e-8, its exact
reverse sweep, the ERI-gradient tile pass and the gradient assembly -> E and dE/d(ln alpha), dE/d(coef), dE/d(centres).
Metric (BASELINE.json): ERI+gradient primitive quartets per second = primitive quartets of the system / step time.  On
this dense geometry no Gaussian product underflows and no two-particle weight is exactly zero: every counted quartet
is evaluated, in both passes.

  value : device time only (CUDA events inside the library, inputs already uploaded)
  e2e   : wall clock through the public API (Energy.rhf_grad -> C ABI) with HOST buffers: host planning, H2D of the
          system arrays and the guess, D2H of E + gradient inside the timed region
  --impl reference : the CPU port of the reference's algorithm (oracle/rhf_oracle.cpp, forward-mode gradient exactly as
          Task.py:25-51 does with algopy) on all host threads, on a bounded sample (see cpu_baseline.sample)
After the timed region the default run appends, as extra keys: the round-1 workload (same monomers 24 bohr apart, core
guess, exact-zero skipping live), one integral-direct step, and BASELINE config 5 (1024 perturbed CH4, fused batch path).
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
JK_KERNEL = "k_jk_stream"               # the stored-tile J/K contraction kernel (HBM bound; TMA-fed ring)
DENSE_SPACING = 5.67
SPACING = DENSE_SPACING


def executed_counts(spacing):
    """ncu-counted FP64 work of the tile kernels on THIS workload (profiles/fp64_executed*.json, written by
    scripts/fp64_executed.py from an ncu metrics pass of this same command).  The counts are data dependent, so there is
    one file per geometry.  None when absent (the roofline then uses the algorithmic count)."""
    name = "fp64_executed_dense.json" if spacing == DENSE_SPACING else ("fp64_executed.json" if spacing == 24.0 else None)
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
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


def cluster(n_waters, spacing=None):
    from diffiqult_b200 import synthetic
    return synthetic.water_cluster(n_waters, spacing=SPACING if spacing is None else spacing)


def workload(n_waters, spacing=None):
    from diffiqult_b200 import System_mol
    from diffiqult_b200.Basis import basis_set_3G_STO
    return System_mol(cluster(n_waters, spacing), basis_set_3G_STO, 10 * n_waters, "w%d" % n_waters)


def use_guess(spacing):
    """The dense grids need the readguess start; from ~20 bohr on the reference converges from its core guess."""
    return spacing < 20.0


def prim_quartets(s):
    k = np.asarray(s.list_contr, dtype=np.int64)
    kk = np.outer(k, k)[np.triu_indices(len(k))]          # primitives per unique function pair
    tot = int(kk.sum())
    return (tot * tot + int((kk * kk).sum())) // 2       # sum over unique pair-of-pairs of KK_x * KK_y


def evaluated_prim_quartets(s):
    """Unique primitive quartets that are really expanded: both Gaussian-product prefactors nonzero in double precision
    (exp(-a b |AB|^2 / (a+b)) underflows to exactly 0 beyond ~745; such quartets are exact zeros in the reference too and are
    left out of the kernels' loops).  Same count as prim_quartets() with the surviving primitive pairs of each function pair."""
    al = np.asarray(s.alpha, dtype=np.float64)
    xyz = np.asarray(s.xyz, dtype=np.float64)
    k = np.asarray(s.list_contr, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(k)])
    n = len(k)
    r2 = ((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1)
    alive = np.zeros((n, n), dtype=np.int64)
    kmax = int(k.max())
    a = np.zeros((n, kmax))
    m = np.zeros((n, kmax), dtype=bool)
    for i in range(n):
        a[i, :k[i]] = al[off[i]:off[i + 1]]
        m[i, :k[i]] = True
    with np.errstate(under="ignore"):
        for p in range(kmax):
            for q in range(kmax):
                ap, bq = a[:, p][:, None], a[:, q][None, :]
                ok = m[:, p][:, None] & m[:, q][None, :]
                ex = np.exp(-np.where(ok, ap * bq / np.where(ok, ap + bq, 1.0), 0.0) * r2)
                alive += (ok & (ex != 0.0)).astype(np.int64)
    kk = alive[np.triu_indices(n)]
    tot = int(kk.sum())
    return (tot * tot + int((kk * kk).sum())) // 2


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


def oracle_threads():
    """All host cores, set explicitly: torchrun exports OMP_NUM_THREADS=1, which would serialise the OpenMP port."""
    from oracle import cpu_oracle as orc
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    orc.lib().orc_set_num_threads(C.c_int(n))
    return n


def cpu_reference_step(sample_waters, spacing=None):
    """One bounded step of the CPU port: energy + forward-mode gradient (argnum 0,1,2) of (H2O)_sample cut from the same
    cluster generator at the same spacing, started the same way (monomer readguess on the dense grid)."""
    from oracle import cpu_oracle as orc
    from diffiqult_b200 import synthetic
    from diffiqult_b200.Basis import basis_set_3G_STO
    spacing = SPACING if spacing is None else spacing
    cores = oracle_threads()
    s = workload(sample_waters, spacing)
    a = orc.SysArrays(s)
    t0 = time.perf_counter()
    if use_guess(spacing):
        def solve(sm):
            o = orc.rhf(orc.SysArrays(sm), max_scf=100)
            assert o["status"] == 0
            return o["C"]
        orc.set_guess(synthetic.monomer_guess(cluster(sample_waters, spacing), basis_set_3G_STO, solve=solve))
    try:
        e = orc.rhf(a, max_scf=100)
        assert e["status"] == 0, "the CPU sample did not converge"
        for n in (0, 1, 2):
            g = orc.rhf_grad(a, n, max_scf=100)
            assert g["status"] == 0
    finally:
        orc.set_guess(None)
    dt = time.perf_counter() - t0
    return evaluated_prim_quartets(s) / dt, dt, e["E"], cores


def workload_name(waters, spacing):
    return "(H2O)_%d STO-3G RHF energy + full parameter gradient (argnum 0,1,2), %.2f bohr grid, %s" % (
        waters, spacing, "monomer readguess" if use_guess(spacing) else "core guess")


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
    sample_txt = ("(H2O)_%d / STO-3G cut from the same cluster generator (%.2f bohr grid, %s): RHF energy + forward-mode gradient "
                  "for argnum 0,1,2 (the reference's algorithm, Task.py:25-51), C++/OpenMP port oracle/rhf_oracle.cpp on %d "
                  "threads, %.1f s per step; rate = primitive quartets of the sample / time.  The reference itself is "
                  "single-threaded Python 2 + algopy and cannot run (H2O)_%d at all" % (
                      sample, SPACING, "monomer readguess" if use_guess(SPACING) else "core guess", cores, ms * 1e-3, args.waters))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.waters, SPACING), "grid_spacing_bohr": SPACING,
                       "n_basis": int(s_full.nbasis), "n_primitives": int(len(s_full.alpha)),
                       "primitive_quartets": evaluated_prim_quartets(s_full), "primitive_quartets_nominal": prim_quartets(s_full)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config5(torch, dist, world, rank, local, n_mol=1024, reps=3):
    """BASELINE config 5: 1024 perturbed CH4 / STO-3G, energy + 81 gradient components each, through the fused
    batch path; molecules shard contiguously over ranks (no data-path collective), rows gathered with one all_gather."""
    from diffiqult_b200 import System_mol, synthetic, batch
    from diffiqult_b200.Basis import basis_set_3G_STO
    syss = [System_mol(m, basis_set_3G_STO, 10, "ch4") for m in synthetic.perturbed_methanes(n_mol)]
    mine = [syss[i] for i in batch.shard(n_mol, world, rank)]

    def once():
        E, st, g, stats = batch.rhf_grad_batch_fused(mine, max_scf=50, device=local)
        rows = np.concatenate([E[:, None]] + [np.stack([x[n] for x in g]) for n in range(3)], axis=1)     # [mine][1 + 81]
        if world > 1:
            per = (n_mol + world - 1) // world
            buf = torch.zeros((per, rows.shape[1]), dtype=torch.float64, device="cuda")
            buf[:rows.shape[0]] = torch.from_numpy(rows).cuda()
            out = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(out, buf)
            rows = torch.cat(out)[:n_mol].cpu().numpy()
        return rows, st, stats

    once()
    ts = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows, st, stats = once()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t[0])
    out = {"workload": "%d perturbed CH4 / STO-3G (seed 20260102), RHF energy + gradient argnum 0,1,2, fused batch path "
                       "(dq_rhf_grad_batch), %d molecules per rank" % (n_mol, len(mine)),
           "ms": sec * 1e3, "molecules_per_s": n_mol / sec, "n_gpus": world, "converged_this_rank": int((st == 0).sum()),
           "launches_this_rank": int(stats["launches"]),
           "timing": "wall clock around the public call with host buffers (H2D of the geometries, D2H of E + gradients, "
                     "all_gather of the rows), median of %d, max over ranks" % reps}
    if rank == 0:
        from oracle import cpu_oracle as orc
        cores = oracle_threads()
        pick = list(range(8))
        t0 = time.perf_counter()
        dE, dG = 0.0, 0.0
        for i in pick:
            a = orc.SysArrays(syss[i])
            o = orc.rhf(a, max_scf=50)
            dE = max(dE, abs(o["E"] - rows[i, 0]))
            col = 1
            for n in (0, 1, 2):
                og = orc.rhf_grad(a, n, max_scf=50)
                if og["status"] == 0:
                    dG = max(dG, float(np.abs(og["grad"] - rows[i, col:col + og["grad"].size]).max()))
                col += og["grad"].size
        dt = time.perf_counter() - t0
        out["max_abs_dE_vs_oracle_on_8_molecules"] = dE
        out["max_abs_dgrad_vs_oracle_on_8_molecules"] = dG
        out["cpu_baseline"] = {"value": len(pick) / dt, "unit": "molecules/s", "cores": cores, "kind": "port",
                               "sample": "the first 8 of the 1024 molecules: energy + forward-mode gradient argnum 0,1,2 with "
                                         "oracle/rhf_oracle.cpp, %.1f s" % dt}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffiqult_b200 import Energy, _capi, distributed, synthetic
    from diffiqult_b200.Basis import basis_set_3G_STO
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
    nq_nominal = prim_quartets(s)
    nq = evaluated_prim_quartets(s)          # what the metric counts: primitive quartets that are really expanded
    lnalpha = np.log(s.alpha)
    stream = distributed.current_stream_handle()
    base_kw = dict(device=local, stream=stream, world_size=world, rank=rank, allreduce=allreduce)
    kw = dict(base_kw, max_scf=args.max_scf, jk_mode=args.jk_mode, screen_threshold=args.screen)

    # the readguess input: 32 monomer SCFs on this GPU, once, before the timed region (every rank builds the same matrix)
    guess, guess_ms = None, 0.0
    if use_guess(args.spacing) and not args.core_guess:
        synthetic.monomer_guess(cluster(1), basis_set_3G_STO, device=local)       # warm the library up
        t0 = time.perf_counter()
        guess = synthetic.monomer_guess(cluster(args.waters), basis_set_3G_STO, device=local)
        guess_ms = (time.perf_counter() - t0) * 1e3

    def step(**over):
        k = dict(kw, **over)
        return Energy.rhf_grad(lnalpha, s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, guess_C=guess, **k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak = C.c_double(0.0)
    _capi.check(_capi.lib().dq_fp64_peak(C.c_int32(local), C.byref(peak)))
    # device spin-up: the first process on a freshly started box ran its first ~20 s of FP64 work up to 1.6x slower
    # (measured: ERI pass 330 / 234 / 208 ms in three consecutive processes) - a few seconds of the FP64 microbenchmark before
    # the W warm-up steps bring the GPU to its steady state; not part of any timed or warm-up step
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < args.spinup:
        _capi.check(_capi.lib().dq_fp64_peak(C.c_int32(local), C.byref(peak)))
    for _ in range(args.warmup):
        r = step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    if args.profile_range:            # ncu --profile-from-start off: only the timed steps are captured (not the guess SCFs)
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    stats = []
    for _ in range(args.steps):
        r = step()
        stats.append(r["stats"])
    barrier()
    wall = time.perf_counter() - t0
    if args.profile_range:
        torch.cuda.profiler.stop()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    dev_ms = float(np.sum([st["ms_total"] - st["ms_setup"] for st in stats]))
    red = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    wall, dev_ms = float(red[0]), float(red[1])

    def timed_step(fn, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        t1 = time.perf_counter()
        out = fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0]) * 1e3

    extras = {}
    if not args.no_extras and args.screen <= 0:
        # (1) integral-direct Fock builds (north_star 3): one step, tiles recomputed inside every J/K contraction
        if args.jk_mode != "direct":
            rd, ms = timed_step(lambda: step(jk_mode="direct"), warm=0)
            extras["integral_direct"] = {
                "ms_per_step": ms, "energy_hartree": rd["E"], "fock_builds": rd["stats"]["fock_builds"],
                "ms_jk_per_build": rd["stats"]["ms_jk"] / max(rd["stats"]["fock_builds"], 1),
                "ms_eri_pass_stored_mode": float(np.mean([x["ms_eri"] for x in stats])),
                "max_abs_gradient_difference_to_stored_run": float(max(np.abs(rd["grad"][n] - r["grad"][n]).max() for n in range(3))),
                "what": "the same step with jk_mode=direct: no tile store, every Fock build (forward and reverse sweep) "
                        "re-evaluates all ERI tiles and contracts them in registers (k_jk_direct); measured once after the "
                        "timed region"}
        # (2) the round-1 workload: the same monomers 24 bohr apart, core guess; exact-zero skipping is live there
        if args.spacing == DENSE_SPACING and args.waters == 32:
            s24 = workload(args.waters, 24.0)
            ln24 = np.log(s24.alpha)

            def step24():
                return Energy.rhf_grad(ln24, s24.coef, s24.xyz, s24.l, s24.charges, s24.atom, s24.list_contr, s24.ne, **kw)
            r24, ms = timed_step(step24, warm=2)
            st24 = r24["stats"]
            extras["sparse_24bohr_core_guess"] = {
                "ms_per_step": ms, "energy_hartree": r24["E"], "scf_iterations": st24["scf_iterations"], "converged": r24["status"] == 0,
                "nominal_primitive_quartets_per_s": nq_nominal / (ms * 1e-3),
                "evaluated_primitive_quartets_per_s": evaluated_prim_quartets(s24) / (ms * 1e-3),
                "primitive_pairs_underflowed_frac": st24["n_prim_pairs_underflowed"] / max(st24["n_prim_pairs"], 1),
                "gradient_quartets_with_exactly_zero_weight_frac": st24["n_grad_quartets_zero_weight"] / max(st24["n_quartets"], 1),
                "phases_ms": {k: st24[k] for k in ("ms_eri", "ms_scf", "ms_reverse", "ms_eri_grad")},
                "what": "round 1's headline workload (core guess converges only on this wide grid): 62 % of the primitive "
                        "pairs underflow to exactly 0 and 80 % of the gradient quartets carry exactly zero weight, so most "
                        "NOMINAL quartets are never expanded - a wall time, not a kernel throughput"}
        # (3) BASELINE config 5
        if not args.no_config5:
            extras["config5"] = config5(torch, dist, world, rank, local)
    if rank == 0:
        st = stats[-1]
        K = args.steps
        ms_grad_kernel = float(np.mean([x["ms_eri_grad"] for x in stats]))
        ms_eri_kernel = float(np.mean([x["ms_eri"] for x in stats]))
        builds = max(st["fock_builds"], 1)
        ms_jk_build = float(np.mean([x["ms_jk"] for x in stats])) / builds
        npq = st["n_prim_quartets"]                      # primitive quartets this rank evaluates per ERI pass
        ex = executed_counts(args.spacing) if (args.waters == 32 and args.screen <= 0) else None
        gk = "k_eri_grad_sparse" if st["grad_kernel_sparse"] else "k_eri_grad_chunks"      # (incl. k_eri_grad_deferred, the second kernel of the L >= 2 classes)
        vk = "k_eri_tiles_compact" if st["eri_kernel_compact"] else "k_eri_tiles"      # (incl. k_eri_tiles_def, the L >= 2 classes)
        if st["jk_direct"]:
            vk = "k_jk_direct"
        if ex and (gk not in ex or vk not in ex):
            ex = None
        hbm_peak, hbm_src = measured_hbm_peak()
        jk_bytes = st["n_tiles"] * 2048.0                # 8 B per stored contracted ERI, every tile read once per build
        jk_gbs = jk_bytes / (ms_jk_build * 1e-3) * 1e-9 if ms_jk_build > 0 else 0.0
        step_ms = dev_ms / K
        peak_src = ("measured live: dq_fp64_peak FMA loop (MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 FMA x 2 x "
                    "1.965 GHz = 37.2)")

        def tile_kernel(name, ms, alg_flop, key):
            """roofline entry of one tile-kernel family.  `achieved` counts the FP64 flop the kernel EXECUTES on this workload
            (ncu, profiles/fp64_executed*.json); SURVEY.md 8(d)'s contract count is reported beside it - it is an upper
            bound on the necessary work, not a count of it, and would put the gradient kernel above peak."""
            alg = npq * alg_flop / (ms * 1e-3) * 1e-12
            d = {"kernel": "%s<*> (7 class launches per step, %.0f %% of the step)" % (name, 100 * ms / step_ms),
                 "bound": "fp64", "bound_note": "FP64 CUDA-core pipe: neither HBM nor the tensor cores bound this kernel",
                 "ms_per_step": ms, "peak": peak.value, "unit": "TFLOP/s", "peak_source": peak_src, "traffic": None}
            contract = {"flop_per_primitive_quartet": alg_flop, "achieved": alg, "frac": alg / peak.value,
                        "note": "SURVEY.md 8(d): the reference's formulas counted op by op with ~100 flop per Boys call (value), and "
                                "the 4x reverse-mode BOUND for value + partials (gradient).  79 % of this cluster's quartets have "
                                "t >= 36-50, where the reference's continued fraction equals its closed form (5 flop), and the "
                                "hand-written adjoint costs < 2x the value: the kernels execute less than the contract counts"}
            if ex:
                e = ex[key]
                exe = npq * e["flop_per_primitive_quartet"] / (ms * 1e-3) * 1e-12
                d.update({"achieved": exe, "frac": exe / peak.value,
                          "achieved_basis": "EXECUTED FP64 flop per primitive quartet on THIS workload (%.1f = 2*DFMA + DMUL + DADD thread "
                                            "instructions counted by ncu over the class launches, %s) x %d primitive quartets of this "
                                            "rank / live CUDA-event time of the class launches" % (
                                                e["flop_per_primitive_quartet"], ex.get("source"), npq),
                          "fp64_pipe_active_ncu": e.get("pipe_fp64_active_frac"),
                          "active_lanes_per_warp_instruction_ncu": e.get("lanes_per_instruction"),
                          "traffic": e.get("dram_bytes_per_step"), "contract_count": contract})
            else:
                d.update({"achieved": alg, "frac": alg / peak.value, "contract_count": contract,
                          "achieved_basis": "SURVEY.md 8(d) contract count (no ncu counts committed for this geometry / kernel)"})
            return d
        k_val = tile_kernel(vk, ms_eri_kernel, FLOP_VALUE, vk)
        k_grad = tile_kernel(gk, ms_grad_kernel, FLOP_GRAD, gk)
        # the dominant kernel family of the step leads; the other tile kernel and the HBM-bound J/K kernel follow
        roofline = dict(k_val if ms_eri_kernel >= ms_grad_kernel else k_grad)
        roofline["other_tile_kernel"] = k_grad if ms_eri_kernel >= ms_grad_kernel else k_val
        if not st["jk_direct"]:
            roofline["jk_kernel"] = {"kernel": JK_KERNEL, "bound": "hbm", "achieved": jk_gbs, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": jk_gbs / hbm_peak, "peak_source": hbm_src, "ms_per_build": ms_jk_build,
                                     "builds_per_step": builds, "algorithmic_bytes_per_build": jk_bytes,
                                     "traffic": ex.get(JK_KERNEL, {}).get("dram_bytes_per_launch") if ex else None}
        h2d = int(lnalpha.nbytes + np.asarray(s.coef).nbytes + s.xyz.nbytes + 4 * (2 * s.nbasis + s.natoms) + s.atom.nbytes +
                  (guess.nbytes if guess is not None else 0))
        d2h = int(8 * (1 + 2 * len(s.alpha) + 3 * s.nbasis) + 8 * (s.nbasis * s.nbasis + s.nbasis))
        uf = st["n_prim_pairs_underflowed"] / max(st["n_prim_pairs"], 1)
        zw = st["n_grad_quartets_zero_weight"] / max(st["n_quartets"], 1)
        line = {
            "metric": METRIC, "value": nq / (dev_ms / K * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": wall / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args.waters, args.spacing), "grid_spacing_bohr": args.spacing,
                       "n_basis": int(s.nbasis), "n_primitives": int(len(s.alpha)),
                       "primitive_quartets": nq, "primitive_quartets_nominal": nq_nominal,
                       "value_counts": "EVALUATED unique primitive quartets (both Gaussian-product prefactors nonzero in double "
                                       "precision); the nominal count also holds the %.1f %% that are exact zeros in the reference as "
                                       "well and are not expanded" % (100.0 * (1.0 - nq / nq_nominal)),
                       "primitive_quartet_slots_walked_per_pass_this_rank": int(npq),
                       "guess": ("readguess = superposition of the %d monomer SCFs (Energy.py:148-157), built once before the "
                                 "timed region in %.1f ms, uploaded every step" % (args.waters, guess_ms)) if guess is not None
                                else "core Hamiltonian (Energy.py:159)",
                       "scf_iterations": st["scf_iterations"], "converged": r["status"] == 0, "energy_hartree": r["E"],
                       "primitive_pairs_underflowed_frac": uf, "gradient_quartets_with_exactly_zero_weight_frac": zw,
                       "evaluated": "every counted primitive quartet is expanded in both passes" if zw == 0 else
                                    "the gradient pass additionally skips quartets of exactly zero weight (fraction above)",
                       "accumulate": "fixed point (order-independent, bit-reproducible)" if not os.environ.get("DQ_ACCUM") else os.environ["DQ_ACCUM"],
                       "jk_mode": ("integral-direct (tiles recomputed in every Fock build)" if st["jk_direct"] else
                                   "stored ERI tiles (%.2f GB per rank)" % (st["n_tiles"] * 2048 / 1e9)),
                       "l2": "ERI tile store larger than L2; every step recomputes everything from host inputs",
                       "screening": ("off: no threshold decides what is evaluated (only exact zeros are left out), as in the reference") if args.screen <= 0 else
                                    ("opt-in threshold %g: %d of %d tiles skipped" % (args.screen, st["n_tiles_skipped"], st["n_tiles"])),
                       "parallelism": "tiles round-robin over %d ranks; allreduce of J|K per Fock build and of the gradient" % world},
            "e2e": {"value": nq / (wall / K), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(np.sum([x["launches"] for x in stats])),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "phases_ms": {k: float(np.mean([x[k] for x in stats])) for k in
                          ("ms_setup", "ms_one_electron", "ms_eri", "ms_scf", "ms_reverse", "ms_eri_grad", "ms_grad_other", "ms_jk", "ms_total")},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            v, dt, e, cores = cpu_reference_step(CPU_BASELINE_WATERS)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "(H2O)_%d STO-3G cut from the same cluster generator (%.2f bohr grid, %s): energy + "
                                              "forward-mode gradient argnum 0,1,2, C++/OpenMP port of the reference algorithm "
                                              "(oracle/rhf_oracle.cpp), %.1f s of CPU work on %d threads" % (
                                                  CPU_BASELINE_WATERS, args.spacing,
                                                  "monomer readguess" if use_guess(args.spacing) else "core guess", dt, cores)}
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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys measured after the timed region "
                    "(integral-direct step, the 24-bohr core-guess workload, config 5)")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    ap.add_argument("--core-guess", action="store_true", help="start from the core Hamiltonian even on a dense grid (does not converge)")
    ap.add_argument("--ref-waters", type=int, default=REF_SAMPLE_WATERS, help="--impl reference: monomers in the bounded CPU sample")
    ap.add_argument("--jk-mode", default="auto", choices=["auto", "stored", "direct"])
    ap.add_argument("--spacing", type=float, default=DENSE_SPACING, help="grid spacing of the water cluster in bohr (default 5.67 = "
                    "SURVEY.md 8(d), started from the monomer readguess; 24: the grid on which the core guess converges)")
    ap.add_argument("--max-scf", type=int, default=100)
    ap.add_argument("--spinup", type=float, default=5.0, help="seconds of FP64 microbenchmark before the warm-up steps (device spin-up)")
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
