"""Batches of independent molecules (BASELINE.json config 5: 1024 perturbed CH4, SURVEY.md 8e "pure replicas").

Every molecule is an independent RHF energy + gradient; there is no exchange between them.

* molecules of ONE structure (same atoms, basis layout, electron count -- the perturbed-geometry batch): the fused device
  path `dq_rhf_grad_batch` (csrc/dq_batch.cu): one thread per (contracted quartet, molecule) for the integral passes and
  one CTA per molecule for the whole SCF and its reverse sweep -- a dozen launches for the whole batch;
* mixed structures: a pool of host threads, each driving its own CUDA stream (`cudaStreamPerThread`) through the
  per-molecule C ABI (ctypes releases the GIL during the call, the library keeps its per-call state thread-local).
Across GPUs the molecules are dealt to ranks/devices in contiguous slices (`shard`).
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import Energy, _capi

STREAM_PER_THREAD = 2   # cudaStreamPerThread handle


def same_structure(systems):
    s0 = systems[0]
    key0 = (tuple(int(z) for z in s0.charges), tuple(tuple(int(v) for v in t) for t in s0.l), tuple(int(k) for k in s0.list_contr),
            int(s0.ne))
    for s in systems[1:]:
        if (tuple(int(z) for z in s.charges), tuple(tuple(int(v) for v in t) for t in s.l), tuple(int(k) for k in s.list_contr),
                int(s.ne)) != key0:
            return False
    return True


def rhf_grad_batch_fused(systems, max_scf=30, device=0, with_grad=True, log=True, stream=None, accumulate="fixed"):
    """All molecules in one set of launches (they must share one structure).  Returns (E[n], status[n], grads, stats) with
    grads[i] = [d/d ln(alpha), d/d coef, d/d xyz] as Energy.rhf_grad returns them."""
    s0, n = systems[0], len(systems)
    a0 = _capi.SystemArrays(np.log(s0.alpha) if log else s0.alpha, s0.coef, s0.xyz, s0.l, s0.charges, s0.atom, s0.list_contr, s0.ne, log)
    alpha = np.ascontiguousarray([np.log(s.alpha) if log else s.alpha for s in systems], dtype=np.float64)
    coef = np.ascontiguousarray([np.ravel(s.coef) for s in systems], dtype=np.float64)
    xyz = np.ascontiguousarray([np.ravel(s.xyz) for s in systems], dtype=np.float64)
    atoms = np.ascontiguousarray([np.ravel(s.atom) for s in systems], dtype=np.float64)
    E, status, niter = np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    ga, gc, gx = np.zeros((n, a0.nprim)), np.zeros((n, a0.nprim)), np.zeros((n, 3 * a0.nbasis))
    opt = _capi.options(max_scf=max_scf, device=device, stream=stream, accumulate=accumulate)
    _capi.check(_capi.lib().dq_rhf_grad_batch(C.byref(a0.struct), C.c_int32(n), atoms.ctypes, alpha.ctypes, coef.ctypes, xyz.ctypes,
                                              C.byref(opt), C.c_int32(1 if with_grad else 0), E.ctypes, status.ctypes, niter.ctypes,
                                              ga.ctypes, gc.ctypes, gx.ctypes))
    grads = [[ga[i], gc[i], gx[i]] for i in range(n)] if with_grad else None
    st = _capi.stats()
    st["niter"] = niter
    return E, status, grads, st


def rhf_grad_batch(systems, max_scf=30, device=0, threads=8, with_grad=True, mode="auto"):
    """systems: list of System_mol.  Returns (E[n], status[n], grads[n] or None) in input order.
    mode: "fused" (one structure, nbasis <= 40), "threads" (any mix), "auto" (fused when possible)."""
    if mode == "fused" or (mode == "auto" and len(systems) > 1 and same_structure(systems) and systems[0].nbasis <= 40 and
                           max(int(k) for k in systems[0].list_contr) ** 2 <= 40):
        E, status, grads, _ = rhf_grad_batch_fused(systems, max_scf=max_scf, device=device, with_grad=with_grad)
        return E, status, grads
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
