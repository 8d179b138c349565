#!/usr/bin/env python
"""Real multi-GPU parity check: launch with torchrun, one rank per GPU (NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_check.py [--waters 4]

Every rank runs the sharded energy + gradient of CH2O / STO-3G, the water dimer and (H2O)_waters; rank 0 then repeats
them unsharded on its own GPU and compares: energies 1e-10, gradients 1e-8 (the north-star tolerances), equal SCF
iteration counts, the tile counts of the ranks adding up to the unsharded count.  Prints one JSON line on rank 0 and
exits non-zero on any mismatch.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--waters", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from diffiqult_b200 import Energy, System_mol, distributed, synthetic
    from diffiqult_b200.Basis import basis_set_3G_STO
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cb = distributed.make_allreduce()
    stream = distributed.current_stream_handle()
    cases = {"ch2o": System_mol(synthetic.CH2O, basis_set_3G_STO, 16, "ch2o"),
             "w2": System_mol(synthetic.water_cluster(2, spacing=24.0), basis_set_3G_STO, 20, "w2"),
             "w%d" % args.waters: System_mol(synthetic.water_cluster(args.waters, spacing=24.0), basis_set_3G_STO, 10 * args.waters, "wn")}
    report, ok = {}, True
    for name, s in cases.items():
        a = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
        r = Energy.rhf_grad(*a, max_scf=100, device=local, stream=stream, world_size=world, rank=rank, allreduce=cb)
        tiles = torch.tensor([r["stats"]["n_tiles"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(tiles)
        # every rank must hold the same replicated result
        chk = torch.tensor([r["E"]] + [float(np.abs(g).sum()) for g in r["grad"]], dtype=torch.float64, device="cuda")
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = float((hi - lo).abs().max())
        if rank == 0:
            one = Energy.rhf_grad(*a, max_scf=100, device=local, stream=stream)
            dE = abs(one["E"] - r["E"])
            dg = max(float(np.abs(one["grad"][n] - r["grad"][n]).max()) for n in range(3))
            good = (r["status"] == 0 and one["niter"] == r["niter"] and dE < 1e-10 and dg < 1e-8 and spread < 1e-9 and
                    int(tiles[0]) == one["stats"]["n_tiles"])
            ok = ok and good
            report[name] = {"E": r["E"], "dE": dE, "dgrad": dg, "niter": r["niter"], "rank_spread": spread,
                            "tiles_sum": int(tiles[0]), "tiles_single": one["stats"]["n_tiles"], "ok": bool(good)}
    dist.barrier()
    if rank == 0:
        print(json.dumps({"world": world, "allreduce_calls": cb.calls["n"], "cases": report, "ok": bool(ok)}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
