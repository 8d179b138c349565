"""BASELINE config 5: N perturbed CH4 / STO-3G, RHF energy + full parameter gradient (81 numbers per molecule).

    python scripts/bench_batch.py [N=1024] [--threads 8] [--reps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P scripts/bench_batch.py 1024

Fused device path (csrc/dq_batch.cu) vs the per-molecule threaded driver; under torchrun every rank takes a contiguous
slice of the batch (batch.shard: pure replicas, no data-path collective) and the rows are gathered with all_gather.
Prints one JSON line (rank 0).  Timing: wall clock around the public call with host buffers (inputs are host arrays, the
results come back to the host), max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffiqult_b200 import System_mol, synthetic, batch
from diffiqult_b200.Basis import basis_set_3G_STO

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="?", type=int, default=1024)
ap.add_argument("--threads", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--skip-threaded", action="store_true")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
syss = [System_mol(m, basis_set_3G_STO, 10, "ch4") for m in synthetic.perturbed_methanes(args.n)]
mine = [syss[i] for i in batch.shard(args.n, world, rank)]
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def gather(E, grads):
    rows = np.concatenate([E[:, None]] + [np.stack([g[n] for g in grads]) for n in range(3)], axis=1)     # [mine][1 + 81]
    if world == 1:
        return rows
    per = (args.n + world - 1) // world
    buf = torch.zeros((per, rows.shape[1]), dtype=torch.float64, device="cuda")
    buf[:rows.shape[0]] = torch.from_numpy(rows).cuda()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat(out)[:args.n].cpu().numpy()


def timed(fn):
    fn()
    ts = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        ts.append(dt)
    return r, float(np.median(ts))


def fused():
    E, st, g, stats = batch.rhf_grad_batch_fused(mine, max_scf=50, device=local)
    return gather(E, g), st, stats


(rows, st, stats), t_fused = timed(fused)
line = {"workload": "%d perturbed CH4 STO-3G, RHF energy + gradient (argnum 0,1,2)" % args.n, "n_gpus": world,
        "fused_s": t_fused, "fused_ms_per_molecule": t_fused / args.n * 1e3, "molecules_per_s": args.n / t_fused,
        "converged_this_rank": int((st == 0).sum()), "launches_this_rank": stats["launches"],
        "device_ms": {k: stats[k] for k in ("ms_eri", "ms_scf", "ms_eri_grad", "ms_total")},
        "primitive_quartets_per_s": 2 * 0 + stats["n_prim_quartets"] * world / t_fused, "E0": float(rows[0, 0]),
        "checksum": float(np.abs(rows).sum())}
if world == 1 and not args.skip_threaded:
    sub = syss[:min(args.n, 128)]

    def threaded():
        return batch.rhf_grad_batch(sub, max_scf=50, threads=args.threads, mode="threads")
    (E2, st2, g2), t_thr = timed(threaded)
    line["threaded_ms_per_molecule"] = t_thr / len(sub) * 1e3
    line["threaded_sample"] = len(sub)
    line["max_dE_fused_vs_threaded"] = float(np.abs(E2 - rows[:len(sub), 0]).max())
    line["max_dgrad_fused_vs_threaded"] = float(max(np.abs(np.concatenate(g2[i]) - rows[i, 1:]).max() for i in range(len(sub))))
if rank == 0:
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
