"""minimize(): the dispatcher of the reference (diffiqult/Minimize.py:10-16)."""
from .Optimize import minimize_bfgs


def minimize(fun, x0, args=(), method='BFGS', jac=None, tol=None, gtol=None, callback=None, argnum=None, name=None,
             **options):
    if method == 'BFGS':
        return minimize_bfgs(fun, x0, args, argnum, jac, callback, name, gtol, **options)
    raise ValueError('Unknown solver %s' % method)
