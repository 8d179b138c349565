"""BFGS over several argument groups with the reference's line search (host loop).

Restates `diffiqult/Optimize.py:716-898` (minimize_bfgs, a modified copy of SciPy's BFGS), `:268-330`
(wrap_function / wrap_functions) and `diffiqult/Linesearch.py:36-199` (line_search_wolfe1 on MINPACK's DCSRCH, here
SciPy's Python port `scipy.optimize._dcsrch.DCSRCH`; the Fortran `minpack2.dcsrch` the reference imports no longer
exists), falling back to `line_search_wolfe2` (`Linesearch.py:201-572`, SciPy's own copy is used).
Every objective / gradient evaluation is one call into the CUDA library.

Reference behaviour kept so that optimisation TRAJECTORIES match the shipped example outputs
(examples/*/*.out: tests/test_gpu_parity.py::test_bfgs_reproduces_reference_examples):
  * Linesearch.py:87-89 evaluates phi(s) = f(xk + s*pk[0]) -- the FIRST component of the search direction, broadcast
    over all parameters -- while its slope uses the full pk (Linesearch.py:91-97).  `reference_linesearch=True` (default)
    reproduces that; False gives the textbook f(xk + s*pk).
  * the energy-change test compares against the constant 100.0 and never fires (Optimize.py:776,825);
  * a trial point whose SCF fails (99999) aborts the DCSRCH search (Linesearch.py:181-184), then wolfe2 is tried.
Fixed: wrap_function split the flat vector at cumulative PRODUCTS of the shapes (Optimize.py:281-287), which only
works by accident for one or two groups; cumulative sums are used, so argnum=[0,1,2] works too.
"""
import warnings

import numpy as np
from scipy.optimize import OptimizeResult
try:    # private SciPy modules (>= 1.12 ships the Python DCSRCH port; pinned in requirements.txt): fail with a clear message
    from scipy.optimize._dcsrch import DCSRCH
    from scipy.optimize._linesearch import line_search_wolfe2, LineSearchWarning
except ImportError as _e:      # pragma: no cover
    raise ImportError("diffiqult_b200.Optimize needs scipy.optimize._dcsrch.DCSRCH and scipy.optimize._linesearch "
                      "(SciPy >= 1.12, see requirements.txt); found SciPy without them: %s" % _e)

SCF_FAILED = 99999

_status_message = {'success': 'Optimization terminated successfully.',
                   'maxiter': 'Maximum number of iterations has been exceeded.',
                   'pr_loss': 'Desired error not necessarily achieved due to precision loss.'}


class _LineSearchError(RuntimeError):
    pass


class _ScfFailed(Exception):
    pass


def _splitter(shapes):
    sizes = [int(np.prod(s)) for s in shapes]
    cuts = np.cumsum(sizes)[:-1]
    return lambda x: [np.copy(np.reshape(p, s)) for p, s in zip(np.split(x, cuts), shapes)]


def wrap_function(function, args, argnum, shapes):
    """Optimize.py:268-303: f(flat x) -> function(*full argument list with the optimised groups re-inserted)."""
    ncalls = [0]
    split = _splitter(shapes)
    args = list(args)

    def wrapper(x):
        ncalls[0] += 1
        parts, full, rest = split(x), [], iter(args)
        for i in range(len(argnum) + len(args)):
            full.append(parts[argnum.index(i)] if i in argnum else next(rest))
        return function(*full)
    return ncalls, wrapper


def wrap_functions(fprimes, args, argnum, shapes):
    """Optimize.py:305-330: concatenated gradient of all groups."""
    ncalls = [0]
    fns = [wrap_function(fp, args, argnum, shapes)[1] for fp in fprimes]

    def gradients(x):
        ncalls[0] += 1
        return np.concatenate([np.ravel(np.asarray(f(x), dtype=np.float64)) for f in fns])
    return ncalls, gradients


def line_search_wolfe1(f, fprime, xk, pk, gfk, old_fval, old_old_fval, c1=1e-4, c2=0.9, amax=50, amin=1e-8, xtol=1e-14,
                       reference_linesearch=True):
    """Linesearch.py:36-199 (line_search_wolfe1 + scalar_search_wolfe1)."""
    gval = [gfk]
    fc, gc = [0], [0]
    direction = pk[0] if reference_linesearch else pk      # sic: Linesearch.py:89

    def phi(s):
        fc[0] += 1
        v = f(xk + s * direction)
        if v == SCF_FAILED:
            raise _ScfFailed()
        return v

    def derphi(s):
        gval[0] = fprime(xk + s * pk)
        gc[0] += 1
        return np.dot(gval[0], pk)

    derphi0 = np.dot(gfk, pk)
    phi0 = old_fval
    if old_old_fval is not None and derphi0 != 0:
        alpha1 = min(1.0, 1.01 * 2 * (phi0 - old_old_fval) / derphi0)
        if alpha1 < 0:
            alpha1 = 1.0
    else:
        alpha1 = 1.0
    try:
        stp, phi1, phi0, task = DCSRCH(phi, derphi, c1, c2, xtol, amin, amax)(alpha1, phi0=phi0, derphi0=derphi0, maxiter=100)
    except _ScfFailed:
        stp, phi1 = None, SCF_FAILED
    return stp, fc[0], gc[0], phi1, phi0, gval[0]


def _line_search_wolfe12(f, fprime, xk, pk, gfk, old_fval, old_old_fval, xtol=1e-14, reference_linesearch=True):
    """Optimize.py:598-625."""
    ret = line_search_wolfe1(f, fprime, xk, pk, gfk, old_fval, old_old_fval, xtol=xtol,
                             reference_linesearch=reference_linesearch)
    if ret[0] is None:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', LineSearchWarning)
            ret = line_search_wolfe2(f, fprime, xk, pk, gfk, old_fval, old_old_fval, amax=50)
    if ret[0] is None:
        raise _LineSearchError()
    return ret


def minimize_bfgs(fun, x0, args=(), argnum=None, jac=None, callback=None, name=None, gtol=1e-1, etol=1e-5, norm=np.inf,
                  maxiter=30, disp=False, return_all=False, xtol_linew=1e-14, verbose=0, print_out=False,
                  reference_linesearch=True, **unknown_options):
    """Optimize.py:716-898.  x0: list of arrays (one per entry of argnum); jac: list of gradient callables."""
    if unknown_options:
        warnings.warn("Unknown solver options: %s" % ", ".join(map(str, unknown_options)), RuntimeWarning)
    shapes = [np.shape(i) for i in x0]
    x0 = np.concatenate([np.asarray(i, dtype=np.float64).flatten() for i in x0]).flatten()
    if maxiter is None:
        maxiter = len(x0) * 200
    func_calls, f = wrap_function(fun, args, list(argnum), shapes)
    grad_calls, myfprime = wrap_functions(jac, args, list(argnum), shapes)

    val = 100.0                         # sic: never updated (Optimize.py:776)
    k, N = 0, len(x0)
    I = np.eye(N, dtype=int)
    Hk, xk = I, x0
    allvecs = [x0]
    warnflag = 0
    old_fval = f(xk)
    gfk = myfprime(x0)
    old_old_fval = None
    gnorm = np.linalg.norm(gfk, ord=np.inf)
    while (gnorm > gtol) and (k < maxiter):
        if verbose:
            print('STEP: ', k, old_fval)
        pk = -np.dot(Hk, gfk)
        try:
            alpha_k, fc, gc, old_fval, old_old_fval, gfkp1 = _line_search_wolfe12(
                f, myfprime, xk, pk, gfk, old_fval, old_old_fval, xtol=xtol_linew, reference_linesearch=reference_linesearch)
        except _LineSearchError:
            warnflag = 2
            break
        xkp1 = xk + alpha_k * pk
        allvecs.append(xkp1)
        sk = xkp1 - xk
        xk = xkp1
        if gfkp1 is None:
            gfkp1 = myfprime(xkp1)
        yk = gfkp1 - gfk
        gfk = gfkp1
        if callback is not None:
            callback(xk)
        k += 1
        gnorm = np.linalg.norm(gfk, ord=norm)
        if gnorm <= gtol:
            break
        if abs(val - old_fval) <= etol:
            break
        if not np.isfinite(old_fval):
            warnflag = 2
            break
        try:
            rhok = 1.0 / (np.dot(yk, sk))
        except ZeroDivisionError:
            rhok = 1000.0
        if np.isinf(rhok):
            rhok = 1000.0
        A1 = I - sk[:, np.newaxis] * yk[np.newaxis, :] * rhok
        A2 = I - yk[:, np.newaxis] * sk[np.newaxis, :] * rhok
        Hk = np.dot(A1, np.dot(Hk, A2)) + (rhok * sk[:, np.newaxis] * sk[np.newaxis, :])
    fval = old_fval
    if np.isnan(fval):
        warnflag = 2
    if warnflag == 2:
        msg = _status_message['pr_loss']
    elif k >= maxiter:
        warnflag = 1
        msg = _status_message['maxiter']
    else:
        msg = _status_message['success']
    if disp:
        print(msg)
        print("         Current function value: %f\n         Iterations: %d" % (fval, k))
    result = OptimizeResult(fun=fval, jac=gfk, hess_inv=Hk, nfev=func_calls[0], njev=grad_calls[0], status=warnflag,
                            success=(warnflag == 0), message=msg, x=xk, nit=k)
    if return_all:
        result['allvecs'] = allvecs
    return result


_SOLVERS = {'BFGS': minimize_bfgs}


def minimize(fun, x0, args=(), method='BFGS', jac=None, tol=None, gtol=None, callback=None, argnum=None, name=None,
             **options):
    """Solver dispatch with the call signature of the reference's Minimize.minimize (Minimize.py:10-16); BFGS only."""
    solver = _SOLVERS.get(method)
    if solver is None:
        raise ValueError('Unknown solver %s' % method)
    return solver(fun, x0, args=args, argnum=argnum, jac=jac, callback=callback, name=name, gtol=gtol, **options)
