"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes front-end of oracle/rhf_oracle.cpp.

Builds oracle/_build/librhf_oracle.so with g++ on first use.  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "rhf_oracle.cpp")
_LIB = os.path.join(_HERE, "_build", "librhf_oracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    if not force and os.path.isfile(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _LIB, _SRC])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        os.environ.setdefault("OMP_STACKSIZE", "64M")
        _lib = C.CDLL(build())
        _lib.orc_eri_index.restype = C.c_long
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ltypes(l):
    return _i([0 if max(t) == 0 else 1 + list(t).index(max(t)) for t in l])


class SysArrays(object):
    """Flat arrays of a System_mol-like object (attributes alpha, coef, xyz, l, list_contr, charges, atom, ne)."""

    def __init__(self, s):
        self.alpha, self.coef, self.xyz = _f(s.alpha), _f(s.coef), _f(s.xyz)
        self.ltype, self.contr = ltypes(s.l), _i(s.list_contr)
        self.charges, self.atoms = _i(s.charges), _f(s.atom)
        self.ne, self.nbasis, self.natoms, self.nprim = int(s.ne), len(s.list_contr), len(s.charges), len(s.alpha)


def boys(t, m):
    t = _f(np.atleast_1d(t))
    F, dF = np.zeros_like(t), np.zeros_like(t)
    rc = lib().orc_boys(C.c_int(t.size), t.ctypes, C.c_int(m), F.ctypes, dF.ctypes)
    assert rc == 0
    return F, dF


def normalization(a):
    out = np.zeros(a.nprim)
    lib().orc_normalization(C.c_int(a.nbasis), a.contr.ctypes, a.ltype.ctypes, a.alpha.ctypes, a.coef.ctypes, out.ctypes)
    return out


def one_electron(a, coefn):
    n = a.nbasis
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    rc = lib().orc_one_electron(C.c_int(a.natoms), a.charges.ctypes, a.atoms.ctypes, C.c_int(n), a.contr.ctypes,
                                a.ltype.ctypes, a.alpha.ctypes, _f(coefn).ctypes, a.xyz.ctypes, S.ctypes, T.ctypes, V.ctypes)
    assert rc == 0
    return S, T, V


def eri_size(n):
    m = n * (n + 1) // 2
    return m * (m + 1) // 2


def eri(a, coefn):
    out = np.zeros(eri_size(a.nbasis))
    rc = lib().orc_eri(C.c_int(a.nbasis), a.contr.ctypes, a.ltype.ctypes, a.alpha.ctypes, _f(coefn).ctypes, a.xyz.ctypes,
                       out.ctypes)
    assert rc == 0
    return out


def eri_range(a, coefn, q0, q1):
    out = np.zeros(q1 - q0)
    lib().orc_eri_range(C.c_int(a.nbasis), a.contr.ctypes, a.ltype.ctypes, a.alpha.ctypes, _f(coefn).ctypes,
                        a.xyz.ctypes, C.c_long(q0), C.c_long(q1), out.ctypes)
    return out


def fock(H, eri_vec, D):
    n = H.shape[0]
    F = np.zeros((n, n))
    lib().orc_fock(C.c_int(n), _f(H).ctypes, _f(eri_vec).ctypes, _f(D).ctypes, F.ctypes)
    return F


def eri_index(i, j, k, l, n):
    return lib().orc_eri_index(C.c_int(i), C.c_int(j), C.c_int(k), C.c_int(l), C.c_int(n))


def set_guess(C0):
    """readguess (Energy.py:148-157): AO x MO matrix used to build the first Fock matrix, or None."""
    global _guess_keep
    _guess_keep = None if C0 is None else _f(C0)
    lib().orc_set_guess(C.c_void_p(None) if C0 is None else _guess_keep.ctypes)


_guess_keep = None


def rhf(a, max_scf=100):
    """-> dict(E, status (0 converged / 1 not: the reference's 99999), niter, esteps, C, eps)."""
    n = a.nbasis
    E, nit = C.c_double(0), C.c_int(0)
    es, Cm, eps = np.zeros(1024), np.zeros((n, n)), np.zeros(n)
    rc = lib().orc_rhf(C.c_int(a.natoms), a.charges.ctypes, a.atoms.ctypes, C.c_int(n), a.contr.ctypes, a.ltype.ctypes,
                       C.c_int(a.ne), _f(np.log(a.alpha)).ctypes, a.coef.ctypes, a.xyz.ctypes, C.c_int(max_scf),
                       C.byref(E), C.byref(nit), es.ctypes, Cm.ctypes, eps.ctypes)
    return dict(E=E.value, status=rc, niter=nit.value, esteps=es[:nit.value].copy(), C=Cm, eps=eps)


def rhf_grad(a, argnum, max_scf=100):
    """Forward-mode dE/d(argument group) through the truncated SCF, as Task.py:25-51 does with algopy."""
    P = 3 * a.nbasis if argnum == 2 else a.nprim
    g = np.zeros(P)
    E, nit = C.c_double(0), C.c_int(0)
    rc = lib().orc_rhf_grad(C.c_int(a.natoms), a.charges.ctypes, a.atoms.ctypes, C.c_int(a.nbasis), a.contr.ctypes,
                            a.ltype.ctypes, C.c_int(a.ne), _f(np.log(a.alpha)).ctypes, a.coef.ctypes, a.xyz.ctypes,
                            C.c_int(max_scf), C.c_int(argnum), C.byref(E), C.byref(nit), g.ctypes)
    return dict(E=E.value, status=rc, niter=nit.value, grad=g)
