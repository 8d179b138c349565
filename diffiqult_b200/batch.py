"""Batches of independent molecules (BASELINE.json config 5: 1024 perturbed CH4, SURVEY.md 8e "pure replicas").

Every molecule is an independent RHF energy + gradient; there is no exchange between them.  Round-1 implementation:
a pool of host threads, each driving its own CUDA stream (`cudaStreamPerThread`) through the C ABI (ctypes releases the
GIL during the call, the library keeps its per-call state thread-local), so the latency-bound small-molecule SCFs of
different molecules overlap on one GPU; across GPUs the molecules are dealt to ranks/devices in contiguous slices.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import Energy

STREAM_PER_THREAD = 2   # cudaStreamPerThread handle


def rhf_grad_batch(systems, max_scf=30, device=0, threads=8, with_grad=True):
    """systems: list of System_mol.  Returns (E[n], status[n], grads[n] or None) in input order."""
    fn = Energy.rhf_grad if with_grad else Energy.rhf

    def one(s):
        return fn(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=max_scf,
                  device=device, stream=STREAM_PER_THREAD)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, systems))
    E = np.array([r["E"] for r in res])
    status = np.array([r["status"] for r in res])
    grads = [r["grad"] for r in res] if with_grad else None
    return E, status, grads


def shard(n_items, world_size, rank):
    """Contiguous slice of a batch owned by `rank` (1024 molecules -> 128 per GPU)."""
    per = (n_items + world_size - 1) // world_size
    return range(min(rank * per, n_items), min((rank + 1) * per, n_items))
