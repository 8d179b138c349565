"""ctypes binding of libdiffiqult_b200.so (include/diffiqult_b200.h).

The library holds every kernel of the hot path; there is NO Python/NumPy fallback: if the library is missing
or no CUDA device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libdiffiqult_b200.so")

ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)   # (dev ptr, count, cudaStream_t, user)


class dq_system(C.Structure):
    _fields_ = [("natoms", C.c_int32), ("charges", C.c_void_p), ("atoms", C.c_void_p),
                ("nbasis", C.c_int32), ("contr", C.c_void_p), ("ltype", C.c_void_p),
                ("nprim", C.c_int32), ("alpha", C.c_void_p), ("coef", C.c_void_p), ("xyz", C.c_void_p),
                ("ne", C.c_int32), ("log_alpha", C.c_int32)]


class dq_options(C.Structure):
    _fields_ = [("max_scf", C.c_int32), ("e_tol", C.c_double), ("device", C.c_int32), ("stream", C.c_void_p),
                ("world_size", C.c_int32), ("rank", C.c_int32), ("allreduce", ALLREDUCE_FN),
                ("allreduce_user", C.c_void_p), ("jk_mode", C.c_int32), ("direct_batch_bytes", C.c_int64),
                ("screen_threshold", C.c_double), ("guess_C", C.c_void_p), ("accumulate", C.c_int32)]


class dq_stats(C.Structure):
    _fields_ = [("ms_setup", C.c_double), ("ms_one_electron", C.c_double), ("ms_eri", C.c_double),
                ("ms_scf", C.c_double), ("ms_reverse", C.c_double), ("ms_eri_grad", C.c_double),
                ("ms_grad_other", C.c_double), ("ms_total", C.c_double), ("launches", C.c_int64), ("n_tiles", C.c_int64),
                ("n_quartets", C.c_int64), ("n_prim_quartets", C.c_int64), ("scf_iterations", C.c_int32),
                ("fock_builds", C.c_int32), ("grad_terms", C.c_int32), ("jk_direct", C.c_int32), ("n_tiles_skipped", C.c_int64),
                ("ms_jk", C.c_double), ("n_grad_quartets_zero_weight", C.c_int64), ("n_prim_pairs", C.c_int64),
                ("n_prim_pairs_underflowed", C.c_int64), ("grad_kernel_sparse", C.c_int32), ("eri_kernel_compact", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["dq_plan", "dq_normalization", "dq_one_electron", "dq_eri_packed", "dq_fock", "dq_rhf", "dq_rhf_grad", "dq_rhf_grad_ex", "dq_rhf_grad_batch", "dq_eigh",
           "dq_boys", "dq_fp64_peak", "dq_get_stats", "dq_last_error", "dq_device_count", "dq_release_cache"]

_lib = None


class DiffiqultB200Error(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DiffiqultB200Error(
                "%s is missing: build it with `python -m diffiqult_b200.build` (there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.dq_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(_lib, name)
    return _lib


def check(status, allow=(0,)):
    if status not in allow:
        raise DiffiqultB200Error("diffiqult_b200 status %d: %s" % (status, lib().dq_last_error().decode()))
    return status


def stats():
    s = dq_stats()
    lib().dq_get_stats(C.byref(s))
    return s.as_dict()


def plan(sysarrays, world_size=1, rank=0):
    """Host-only work decomposition of one rank: dict(blocks, block_pairs, tiles, quartets, prim_quartets, tiles_total)."""
    out = (C.c_int64 * 6)()
    check(lib().dq_plan(C.byref(sysarrays.struct), C.c_int32(world_size), C.c_int32(rank), out))
    return dict(zip(("blocks", "block_pairs", "tiles", "quartets", "prim_quartets", "tiles_total"), [int(x) for x in out]))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ltypes(l):
    """reference l tuples -> 0 (s) / 1..3 (p along x,y,z): la.index(max(la)) in Integrals.py:92,166,336."""
    out = []
    for t in l:
        t = tuple(int(x) for x in t)
        if max(t) == 0:
            out.append(0)
        elif max(t) == 1 and sum(t) == 1:
            out.append(1 + t.index(1))
        else:
            raise NotImplementedError("only s and p Gaussians are implemented (as in the reference): l=%r" % (t,))
    return _i(out)


class SystemArrays(object):
    """Keeps the contiguous arrays alive and exposes the dq_system struct over them."""

    def __init__(self, alpha, coef, xyz, l, charges, atoms, contr, ne, log_alpha):
        self.alpha, self.coef, self.xyz = _f(alpha).ravel(), _f(coef).ravel(), _f(xyz).reshape(-1)
        self.ltype = l if (isinstance(l, np.ndarray) and l.dtype == np.int32 and l.ndim == 1) else ltypes(l)
        self.contr, self.charges, self.atoms = _i(contr).ravel(), _i(charges).ravel(), _f(atoms).reshape(-1)
        self.nbasis, self.nprim, self.natoms = int(self.contr.size), int(self.alpha.size), int(self.charges.size)
        if self.xyz.size != 3 * self.nbasis or self.ltype.size != self.nbasis:
            raise ValueError("xyz / l do not match the number of basis functions")
        if self.coef.size != self.nprim or int(self.contr.sum()) != self.nprim:
            raise ValueError("alpha / coef / contr_list do not match")
        self.struct = dq_system(self.natoms, self.charges.ctypes.data, self.atoms.ctypes.data, self.nbasis,
                                self.contr.ctypes.data, self.ltype.ctypes.data, self.nprim, self.alpha.ctypes.data,
                                self.coef.ctypes.data, self.xyz.ctypes.data, int(ne), 1 if log_alpha else 0)


_noop_allreduce = ALLREDUCE_FN()


JK_MODES = {"auto": 0, "stored": 1, "direct": 2}
ACCUMULATE = {"default": 0, "fixed": 1, "float": 2}


def options(max_scf=30, e_tol=0.0, device=0, stream=None, world_size=1, rank=0, allreduce=None, jk_mode="auto",
            direct_batch_bytes=0, screen_threshold=0.0, guess_C=None, accumulate="fixed"):
    return dq_options(int(max_scf), float(e_tol), int(device), C.c_void_p(stream) if stream else None, int(world_size),
                      int(rank), allreduce if allreduce is not None else _noop_allreduce, None,
                      JK_MODES[jk_mode] if isinstance(jk_mode, str) else int(jk_mode), int(direct_batch_bytes),
                      float(screen_threshold), guess_C.ctypes.data if guess_C is not None else None,
                      ACCUMULATE[accumulate] if isinstance(accumulate, str) else int(accumulate))
