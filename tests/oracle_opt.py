"""TEST INFRASTRUCTURE: the reference's optimisation host loop (diffiqult_b200/Optimize.py = Optimize.py:716-898 restated)
driven by the CPU ORACLE's energy and forward-mode gradients instead of the CUDA library - what the reference itself would
compute for `Tasks.runtask('Opt', ...)` (Task.py:362-392).  Used by tests/golden/make_opt_golden.py and the GPU tests."""
import numpy as np

from diffiqult_b200 import Optimize
from diffiqult_b200.Task import Tasks
from oracle import cpu_oracle as orc


class _Sys(object):
    pass


def _arrays(full):
    """19-slot rhfenergy argument list (Task.py:134-144) -> oracle arrays"""
    s = _Sys()
    s.alpha = np.exp(np.asarray(full[0], dtype=float)) if full[12] else np.asarray(full[0], dtype=float)
    s.coef, s.xyz = np.asarray(full[1], dtype=float), np.asarray(full[2], dtype=float)
    s.l, s.charges, s.atom, s.list_contr, s.ne = full[3], full[4], full[5], full[8], full[9]
    return orc.SysArrays(s), int(full[10])


def energy(*full):
    a, max_scf = _arrays(full)
    o = orc.rhf(a, max_scf=max_scf)
    return o["E"] if o["status"] == 0 else Optimize.SCF_FAILED      # Energy.py:322


def gradient(narg):
    def g(*full):
        a, max_scf = _arrays(full)
        o = orc.rhf_grad(a, narg, max_scf=max_scf)
        if o["status"] != 0:
            raise FloatingPointError("oracle gradient status %d" % o["status"])
        return o["grad"]
    return g


def run_opt(system, argnum, max_scf, maxiter, gtol=None):
    """The reference's 'Opt' (Task.py:281-296, 362-392) with the oracle underneath -> OptimizeResult (return_all)."""
    t = Tasks(system, '/tmp/dq_oracle_opt')
    args = t._energy_args(max_scf=max_scf, max_d=max_scf, log=True, printguess=None, readguess=None, name='x', write=False)
    var = [args[i] for i in argnum]
    rest = [a for i, a in enumerate(args) if i not in argnum]
    if gtol is None:
        gtol = 1e-05 if 0 in argnum else 1e-07                      # Task.py:288-291
    return Optimize.minimize_bfgs(energy, var, args=tuple(rest), argnum=list(argnum), jac=[gradient(n) for n in argnum],
                                  gtol=gtol, maxiter=maxiter, return_all=True)
