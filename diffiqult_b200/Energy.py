"""rhfenergy: the inner drop-in boundary (reference: diffiqult/Energy.py:75-324).

The 19-slot positional signature built by Tasks._energy_args (Task.py:134-144) is kept, so callers written
for the reference (Tasks, the BFGS wrapper) work unchanged; the body is one call into the CUDA library."""
import ctypes as C

import numpy as np

from . import _capi

SCF_FAILED = 99999  # Energy.py:322


def _check_max_scf(max_scf):
    # Energy.py:164 `for i in range(max_scf)`: with max_scf < 1 the reference runs no iteration and then reads the undefined
    # E_elec (NameError).  There is no energy to return: reject it here as well.
    if int(max_scf) < 1:
        raise ValueError("max_scf must be >= 1 (the reference's SCF loop needs at least one iteration)")


def _sys(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, log):
    return _capi.SystemArrays(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, log)


def fockmatrix(alpha, coef, xyz, l, charges, atoms, contr_list, D, device=0, jk_mode="auto", direct_batch_bytes=0):
    """Energy.py:29-38 F = Hcore + sum_kl D_kl [2 (ij|kl) - (ik|jl)], integrals rebuilt on the device
    (coef normalised).  Returns (F, Hcore)."""
    a = _capi.SystemArrays(alpha, coef, xyz, l, charges, atoms, contr_list, 0, False)
    n = a.nbasis
    D = np.ascontiguousarray(D, dtype=np.float64)
    F, H = np.zeros((n, n)), np.zeros((n, n))
    opt = _capi.options(device=device, jk_mode=jk_mode, direct_batch_bytes=direct_batch_bytes)
    _capi.check(_capi.lib().dq_fock(C.byref(a.struct), C.byref(opt), D.ctypes, F.ctypes, H.ctypes))
    return F, H


def rhf(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, max_scf=30, log=True, device=0, stream=None,
        world_size=1, rank=0, allreduce=None, jk_mode="auto", direct_batch_bytes=0, screen_threshold=0.0, guess_C=None,
        accumulate="fixed"):
    """Full result of one RHF single point: dict(E, status, niter, esteps, C, eps, stats)."""
    _check_max_scf(max_scf)
    a = _sys(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, log)
    n = a.nbasis
    E, nit = C.c_double(0.0), C.c_int32(0)
    es, Cm, eps = np.zeros(max(int(max_scf), 1)), np.zeros((n, n)), np.zeros(n)
    guess = None if guess_C is None else np.ascontiguousarray(guess_C, dtype=np.float64)
    opt = _capi.options(max_scf=max_scf, device=device, stream=stream, world_size=world_size, rank=rank, allreduce=allreduce,
                        jk_mode=jk_mode, direct_batch_bytes=direct_batch_bytes, screen_threshold=screen_threshold, guess_C=guess,
                        accumulate=accumulate)
    st = _capi.check(_capi.lib().dq_rhf(C.byref(a.struct), C.byref(opt), C.byref(E), C.byref(nit), es.ctypes, Cm.ctypes,
                                        eps.ctypes), allow=(0, 1))
    return dict(E=E.value, status=st, niter=nit.value, esteps=es[:nit.value].copy(), C=Cm, eps=eps, stats=_capi.stats())


def rhf_grad(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, max_scf=30, log=True, device=0, stream=None,
             world_size=1, rank=0, allreduce=None, jk_mode="auto", direct_batch_bytes=0, screen_threshold=0.0, guess_C=None,
             accumulate="fixed", atoms_grad=False):
    """Energy and its gradient w.r.t. (alpha as passed, raw coefficients, Gaussian centres) in one fused pass:
    dict(E, status, niter, grad=[g_alpha, g_coef, g_xyz], stats).  atoms_grad=True adds grad_atoms[natoms*3], the gradient
    w.r.t. the nuclear coordinates with the Gaussians held fixed (the reference's branch Energy.py:125-130).  accumulate="fixed" (default): order-independent
    fixed-point sums, two calls on the same inputs return identical bits; "float": floating-point atomics."""
    _check_max_scf(max_scf)
    a = _sys(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, log)
    E, nit = C.c_double(0.0), C.c_int32(0)
    ga, gc, gx = np.zeros(a.nprim), np.zeros(a.nprim), np.zeros(3 * a.nbasis)
    guess = None if guess_C is None else np.ascontiguousarray(guess_C, dtype=np.float64)
    opt = _capi.options(max_scf=max_scf, device=device, stream=stream, world_size=world_size, rank=rank, allreduce=allreduce,
                        jk_mode=jk_mode, direct_batch_bytes=direct_batch_bytes, screen_threshold=screen_threshold, guess_C=guess,
                        accumulate=accumulate)
    gn = np.zeros(3 * a.natoms) if atoms_grad else None
    st = _capi.check(_capi.lib().dq_rhf_grad_ex(C.byref(a.struct), C.byref(opt), C.byref(E), C.byref(nit), ga.ctypes, gc.ctypes,
                                                gx.ctypes, gn.ctypes if atoms_grad else None), allow=(0, 1))
    out = dict(E=E.value, status=st, niter=nit.value, grad=[ga, gc, gx], stats=_capi.stats())
    if atoms_grad:
        out["grad_atoms"] = gn
    return out


def rhfenergy(alpha_old, coef2, xyz, l, charges, xyz_atom, natoms, nbasis, contr_list, ne, max_scf, max_d, log, eigen,
              printguess, readguess, name, write, dtype, device=0):
    """Energy.py:75: returns E_elec + E_nuc, or 99999 when the SCF did not converge in max_scf iterations.

    eigen=False (density purification) is a dead branch of the reference's public API
    and is rejected; readguess (a .npy of C written by printguess) starts the SCF from that density (Energy.py:148-157)."""
    if not eigen:
        raise NotImplementedError("eigen=False (canonical purification) is unreachable through Tasks and not implemented")
    guess = np.load(readguess) if readguess is not None else None          # Energy.py:148-149
    r = rhf(alpha_old, coef2, xyz, l, charges, xyz_atom, contr_list, ne, max_scf=max_scf, log=log, device=device, guess_C=guess)
    if printguess is not None:
        np.save(printguess, r["C"])  # Energy.py:197-198
    if r["status"] != 0:
        print('E_tot: ' + str(r["E"]) + '\n')
        print('SCF DID NOT CONVERGED')
        return SCF_FAILED
    if write:   # Energy.py:311-316: name.molden is written only for a converged SCF
        from . import Integrals, Molden
        alpha = np.exp(np.asarray(alpha_old, dtype=np.float64)) if log else np.asarray(alpha_old, dtype=np.float64)
        coefn = Integrals.normalization(alpha, coef2, l, contr_list, device=device)
        Molden.write_molden(str(name) + '.molden', r, alpha, coefn, np.asarray(xyz, dtype=np.float64), l, charges,
                            np.asarray(xyz_atom, dtype=np.float64), contr_list, ne, device=device)
    return r["E"]
