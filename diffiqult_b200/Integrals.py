"""Integral entry points with the reference's names and argument order (diffiqult/Integrals.py), each a
single call into the CUDA library.  `dtype` is accepted for signature compatibility and ignored: parameter
derivatives are produced by the fused gradient path (Energy.rhfenergy_grad), not by operator overloading."""
import ctypes as C

import numpy as np

from . import _capi
from .Tools import eri_vector_size


def _arrays(alpha, coef, xyz, l, contr_list, charge=(), atoms=()):
    return _capi.SystemArrays(alpha, coef, xyz, l, charge, atoms, contr_list, 0, False)


def normalization(alpha, c, l, list_contr, dtype=None, device=0):
    """Integrals.py:477-501: raw contraction coefficients -> normalised ones."""
    n = len(list_contr)
    a = _arrays(alpha, c, np.zeros((n, 3)), l, list_contr)
    out = np.zeros(a.nprim)
    opt = _capi.options(device=device)
    _capi.check(_capi.lib().dq_normalization(C.byref(a.struct), C.byref(opt), out.ctypes))
    return out


def _one_electron(alpha, coef, xyz, l, nbasis, contr_list, charge, atoms, device):
    a = _arrays(alpha, coef, xyz, l, contr_list, charge, atoms)
    S, T, V = (np.zeros((a.nbasis, a.nbasis)) for _ in range(3))
    opt = _capi.options(device=device)
    _capi.check(_capi.lib().dq_one_electron(C.byref(a.struct), C.byref(opt), S.ctypes, T.ctypes, V.ctypes))
    return S, T, V


def overlapmatrix(alpha, coef, xyz, l, nbasis, contr_list, dtype=None, device=0):
    """Integrals.py:10-20 (coef already normalised)."""
    return _one_electron(alpha, coef, xyz, l, nbasis, contr_list, (), (), device)[0]


def kineticmatrix(alpha, coef, xyz, l, nbasis, contr_list, dtype=None, device=0):
    """Integrals.py:135-145."""
    return _one_electron(alpha, coef, xyz, l, nbasis, contr_list, (), (), device)[1]


def nuclearmatrix(alpha, coef, xyz, l, nbasis, charge, atoms, numatoms, contr_list, dtype=None, device=0):
    """Integrals.py:51-62."""
    return _one_electron(alpha, coef, xyz, l, nbasis, contr_list, charge, atoms, device)[2]


def one_electron_matrices(alpha, coef, xyz, l, nbasis, charge, atoms, contr_list, device=0):
    """S, T, V in one launch (what rhfenergy uses)."""
    return _one_electron(alpha, coef, xyz, l, nbasis, contr_list, charge, atoms, device)


def erivector(alpha, coef, xyz, l, nbasis, contr_list, dtype=None, device=0):
    """Integrals.py:190-216: packed two-electron repulsion integrals in the reference's order."""
    a = _arrays(alpha, coef, xyz, l, contr_list)
    out = np.zeros(eri_vector_size(a.nbasis))
    opt = _capi.options(device=device)
    _capi.check(_capi.lib().dq_eri_packed(C.byref(a.struct), C.byref(opt), out.ctypes))
    return out
