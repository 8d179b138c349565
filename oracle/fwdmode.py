"""TEST INFRASTRUCTURE ONLY (oracle/): first-order, multi-direction forward-mode
arithmetic standing in for the third-party `algopy.UTPM` the reference depends on.

`algopy` is NOT under /root/reference (un-vendored, unpinned: README.rst:29-34) and is
not installed here, so its published semantics are restated:

  * a UTPM object carries a value and P first-order tangent directions
    (UTPM.init_jacobian seeds one direction per input entry, UTPM.extract_jacobian
    reads them back: reference call sites Task.py:32,37);
  * every arithmetic op pushes the tangents forward;
  * comparisons (`<`, `>`, `max(t,1e-12)`, loop exits) look at the VALUE only
    (reference call sites Tools.py:82,92,120,139-147 and Energy.py:192);
  * `algopy.zeros(shape, dtype=utpm)` allocates a UTPM container (Integrals.py:11,52,
    136,196,479; Energy.py:30,174); with a float dtype it is a plain numpy array;
  * `algopy.eigh` returns (w, Q) with the first-order eigen push-forward
    dw = diag(Q^T dA Q), dQ = Q (H o Q^T dA Q), H_ij = 1/(w_j - w_i) (Eigen.py:70).

Representation: a scalar is a `Dual` (value + numpy tangent vector); vectors/matrices are
numpy object arrays of `Dual`, so the reference's own numpy calls (np.multiply, np.add,
np.subtract, np.square, np.sum, np.exp, np.power ...) work unchanged through numpy's
object loops, exactly the way algopy rides on them.

Nothing outside tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may
import this module.
"""
import numpy as np


class Dual(object):
    __slots__ = ("val", "tan")
    __array_priority__ = 1000.0

    def __init__(self, val, tan):
        self.val = float(val)
        self.tan = tan

    # ---- helpers
    @staticmethod
    def _split(o):
        if isinstance(o, Dual):
            return o.val, o.tan
        return float(o), None

    # ---- arithmetic
    def __add__(self, o):
        v, t = Dual._split(o)
        return Dual(self.val + v, self.tan if t is None else self.tan + t)

    __radd__ = __add__

    def __sub__(self, o):
        v, t = Dual._split(o)
        return Dual(self.val - v, self.tan if t is None else self.tan - t)

    def __rsub__(self, o):
        v, t = Dual._split(o)
        return Dual(v - self.val, -self.tan if t is None else t - self.tan)

    def __mul__(self, o):
        v, t = Dual._split(o)
        if t is None:
            return Dual(self.val * v, self.tan * v)
        return Dual(self.val * v, self.tan * v + t * self.val)

    __rmul__ = __mul__

    def __truediv__(self, o):
        v, t = Dual._split(o)
        q = self.val / v
        if t is None:
            return Dual(q, self.tan / v)
        return Dual(q, (self.tan - t * q) / v)

    def __rtruediv__(self, o):
        v, t = Dual._split(o)
        q = v / self.val
        if t is None:
            return Dual(q, self.tan * (-q / self.val))
        return Dual(q, (t - self.tan * q) / self.val)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __neg__(self):
        return Dual(-self.val, -self.tan)

    def __pos__(self):
        return self

    def __abs__(self):
        return self if self.val >= 0 else -self

    def __pow__(self, p):
        if isinstance(p, Dual):
            # x**y = exp(y log x)
            v = self.val ** p.val
            return Dual(v, v * (p.tan * np.log(self.val) + self.tan * (p.val / self.val)))
        p = float(p)
        v = self.val ** p
        return Dual(v, self.tan * (p * self.val ** (p - 1.0)))

    def __rpow__(self, b):
        v = float(b) ** self.val
        return Dual(v, self.tan * (v * np.log(float(b))))

    # ---- numpy object-loop hooks (np.exp(x) on an object calls x.exp())
    def exp(self):
        v = np.exp(self.val)
        return Dual(v, self.tan * v)

    def log(self):
        return Dual(np.log(self.val), self.tan / self.val)

    def sqrt(self):
        v = np.sqrt(self.val)
        return Dual(v, self.tan * (0.5 / v))

    def conjugate(self):
        return self

    # ---- comparisons on the value coefficient only
    def _v(self, o):
        return o.val if isinstance(o, Dual) else o

    def __lt__(self, o):
        return self.val < self._v(o)

    def __le__(self, o):
        return self.val <= self._v(o)

    def __gt__(self, o):
        return self.val > self._v(o)

    def __ge__(self, o):
        return self.val >= self._v(o)

    def __eq__(self, o):
        return self.val == self._v(o)

    def __ne__(self, o):
        return self.val != self._v(o)

    __hash__ = None

    def __float__(self):
        return self.val

    def __repr__(self):
        return "Dual(%r, P=%d)" % (self.val, self.tan.size)


# ------------------------------------------------------------------ containers
def _is_dual_container(x):
    if isinstance(x, Dual):
        return True
    return isinstance(x, np.ndarray) and x.dtype == object and x.size > 0 and any(
        isinstance(e, Dual) for e in x.flat)


def _ndir(x):
    if isinstance(x, Dual):
        return x.tan.size
    for e in x.flat:
        if isinstance(e, Dual):
            return e.tan.size
    raise TypeError("no Dual inside")


def split(x, P=None):
    """object array (or Dual / float array) -> (values[...], tangents[P, ...])."""
    if isinstance(x, Dual):
        return np.float64(x.val), x.tan.copy()
    x = np.asarray(x)
    if x.dtype != object:
        v = x.astype(np.float64)
        return v, np.zeros((P,) + v.shape)
    if P is None:
        P = _ndir(x)
    v = np.empty(x.shape)
    t = np.zeros((P,) + x.shape)
    for idx in np.ndindex(*x.shape):
        e = x[idx]
        if isinstance(e, Dual):
            v[idx] = e.val
            t[(slice(None),) + idx] = e.tan
        else:
            v[idx] = float(e)
    return v, t


def join(v, t):
    """(values[...], tangents[P, ...]) -> object array of Dual."""
    v = np.asarray(v)
    out = np.empty(v.shape, dtype=object)
    for idx in np.ndindex(*v.shape):
        out[idx] = Dual(v[idx], np.array(t[(slice(None),) + idx]))
    return out


class UTPM(object):
    """The two classmethods the reference calls (Task.py:32,37)."""

    @staticmethod
    def init_jacobian(x):
        x = np.asarray(x, dtype=np.float64)
        P = x.size
        out = np.empty(x.shape, dtype=object)
        for n, idx in enumerate(np.ndindex(*x.shape)):
            t = np.zeros(P)
            t[n] = 1.0
            out[idx] = Dual(x[idx], t)
        return out

    @staticmethod
    def extract_jacobian(y):
        if isinstance(y, Dual):
            return y.tan.copy()
        if isinstance(y, np.ndarray) and y.dtype == object:
            v, t = split(y)
            return np.moveaxis(t, 0, -1)  # algopy: directions last
        raise TypeError("extract_jacobian on a non-UTPM result: %r" % (y,))


def zeros(shape, dtype=float):
    if _is_dual_container(dtype):
        P = _ndir(dtype)
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        shape = tuple(int(s) for s in shape)
        out = np.empty(shape, dtype=object)
        for idx in np.ndindex(*shape):
            out[idx] = Dual(0.0, np.zeros(P))
        return out
    if isinstance(shape, tuple):
        shape = tuple(int(s) for s in shape)
    else:
        shape = int(shape)
    return np.zeros(shape)


def exp(x):
    return np.exp(x)


def log(x):
    return np.log(x)


def sqrt(x):
    return np.sqrt(x)


def transpose(x):
    return np.transpose(x)


def diag(x):
    return np.diag(x)


def sum(x, axis=None):  # noqa: A001  (mirrors algopy.sum)
    return np.sum(x, axis=axis)


def dot(a, b):
    da, db = _is_dual_container(a), _is_dual_container(b)
    if not (da or db):
        return np.dot(a, b)
    P = _ndir(a) if da else _ndir(b)
    av, at = split(a, P)
    bv, bt = split(b, P)
    v = av @ bv
    t = np.einsum("pij,jk->pik", at, bv) + np.einsum("ij,pjk->pik", av, bt)
    return join(v, t)


# eigen push-forward: pairs closer than this are treated as a repeated eigenvalue.
DEGENERACY_TOL = 1e-9
# and a tangent coupling inside such a pair larger than this means the plain
# first-order formula is not valid for that direction -> refuse to produce a golden.
COUPLING_TOL = 1e-7


def eigh(a):
    if not _is_dual_container(a):
        return np.linalg.eigh(a)
    av, at = split(a)
    w, q = np.linalg.eigh(av)
    m = np.einsum("ji,pjk,kl->pil", q, at, q)  # Q^T dA Q per direction
    dw = np.einsum("pii->pi", m)
    gap = w[None, :] - w[:, None]  # w_j - w_i
    close = np.abs(gap) < DEGENERACY_TOL
    h = np.where(close, 0.0, 1.0 / np.where(close, 1.0, gap))
    off = close & ~np.eye(len(w), dtype=bool)
    if off.any():
        worst = np.abs(m[:, off]).max()
        if worst > COUPLING_TOL:
            raise FloatingPointError(
                "eigh push-forward: repeated eigenvalue with tangent coupling %.2e" % worst)
    dq = np.einsum("ij,pjk->pik", q, h[None] * m)
    return join(w, dw), join(q, dq)
