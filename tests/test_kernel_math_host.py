"""The kernels' scalar math (diffiqult_b200/csrc/dq_math.cuh) compiled for the host and compared with the
oracle.  This is a development / regression check that runs without a GPU; the GPU parity tests proper are
in test_gpu_*.py.  The product never executes this host build."""
import ctypes as C
import itertools
import os
import subprocess

import numpy as np
import pytest

from oracle import cpu_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "math_host.cpp")
HDR = os.path.join(os.path.dirname(HERE), "diffiqult_b200", "csrc", "dq_math.cuh")
LIB = os.path.join(HERE, "host_harness", "_build_math_host.so")


@pytest.fixture(scope="module")
def hh():
    if not os.path.isfile(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def _boys_grid():
    t = [np.array([0.0, 1e-14, 1e-12, 1e-10, 1e-6]), np.linspace(1e-4, 12.0, 24001), np.geomspace(12.0, 1e5, 4000)]
    for bp in (1.5, 2.5, 3.5, 4.5, 5.5):
        t.append(bp + np.array([-1e-6, -1e-12, 0.0, 1e-12, 1e-6]))
    return np.concatenate(t)


def test_boys_value_and_derivative_dense_grid(hh):
    t = _boys_grid()
    F = np.zeros((t.size, 5))
    dF = np.zeros((t.size, 5))
    hh.hh_boys(C.c_int(t.size), t.ctypes, F.ctypes, dF.ctypes)
    Fv = np.zeros((t.size, 5))
    hh.hh_boys_value(C.c_int(t.size), t.ctypes, Fv.ctypes)
    assert np.array_equal(F, Fv)
    for m in range(5):
        Fo, dFo = orc.boys(t, m)
        # absolute 2e-14 (F <= 1); the reference's own round trips carry ~1e-14 at tiny t
        assert np.abs(F[:, m] - Fo).max() < 2e-14, m
        # d/dt: the reference differentiates exp(a log t - gln) exp(gln) t^-a term by term; the a/t pieces
        # cancel only up to rounding, so ITS tangent carries noise ~ eps * (a/t) * F at tiny t (1e-4 at
        # t = 1e-12).  Ours is the exact derivative of the same truncated series, hence the t-dependent bound.
        tt = np.maximum(t, 1e-12)
        bound = 5e-14 + 8 * 2.2e-16 * (m + 0.5) / tt * np.abs(Fo)
        assert (np.abs(dF[:, m] - dFo) < bound).all(), m


@pytest.mark.parametrize("L", [0, 1, 2, 3, 4])
def test_boys_regimes_of_every_angular_class_match_the_reference_loops(hh, L):
    """boys_ref<L> switches, per angular class, from the reference's loops to the loop-free mid-range form (t >= 13 / 16 /
    18 / 20 / 20) and to the closed form (t >= 36 / 39.5 / 42.5 / 45 / 47.5).  Dense scan across both boundaries against
    the oracle (= the reference's series / continued fraction with its stopping rules, Tools.py:78-151): relative to
    F_m, because the high orders are small there."""
    t = np.concatenate([np.arange(6.0, 60.0, 0.002), [13.0, 16.0, 18.0, 20.0, 36.0, 39.5, 42.5, 45.0, 47.5, 50.0]])
    F = np.zeros((t.size, 5))
    dF = np.zeros((t.size, 5))
    hh.hh_boys_class(C.c_int(L), C.c_int(t.size), t.ctypes, F.ctypes, dF.ctypes)
    for m in range(L + 1):
        Fo, dFo = orc.boys(t, m)
        assert (np.abs(F[:, m] - Fo) / np.abs(Fo)).max() < 4e-15, (L, m)
        assert (np.abs(dF[:, m] - dFo) / np.abs(dFo)).max() < 3e-12, (L, m)


def _rand_prims(rng, spread):
    ex = np.exp(rng.uniform(np.log(0.15), np.log(40.0), 4))
    co = rng.uniform(0.2, 1.5, 4) * rng.choice([-1.0, 1.0], 4)
    R = rng.normal(scale=spread, size=(4, 3))
    return ex, co, R


@pytest.mark.parametrize("spread", [0.0, 0.3, 1.5, 6.0])
def test_primitive_quartet_value_and_partials_all_classes(hh, spread):
    """All 4^4 (s,px,py,pz) combinations, random exponents/centres: value 1e-13 rel, 20 partials 1e-11."""
    rng = np.random.default_rng(1234 + int(spread * 10))
    lib = orc.lib()
    worst_v, worst_g = 0.0, 0.0
    for l in itertools.product(range(4), repeat=4):
        for _ in range(3):
            ex, co, R = _rand_prims(rng, spread)
            if spread == 0.0:  # coincident centres for some functions: t = 0 clamp
                R[:] = R[0]
            la = np.array(l, dtype=np.int32)
            v, v2 = C.c_double(0), C.c_double(0)
            g = np.zeros(20)
            hh.hh_prim_quartet(ex.ctypes, co.ctypes, R.ctypes, la.ctypes, C.byref(v), C.byref(v2), g.ctypes)
            vo = C.c_double(0)
            go = np.zeros(20)
            lib.orc_eri_primitive_grad(ex.ctypes, co.ctypes, R.ctypes, la.ctypes, C.byref(vo), go.ctypes)
            assert v.value == v2.value
            scale = max(abs(vo.value), 1e-3)
            worst_v = max(worst_v, abs(v.value - vo.value) / scale)
            worst_g = max(worst_g, np.abs(g - go).max() / max(np.abs(go).max(), 1e-3))
    assert worst_v < 1e-12, worst_v
    assert worst_g < 1e-10, worst_g


def test_one_electron_primitive_value_and_partials(hh):
    rng = np.random.default_rng(7)
    lib = orc.lib()
    Z = np.array([8, 1, 1], dtype=np.int32)
    Cn = rng.normal(scale=1.5, size=(3, 3))
    worst = 0.0
    for la, lb in itertools.product(range(4), repeat=2):
        for _ in range(5):
            ex, co, R = _rand_prims(rng, 1.0)
            out, ref = np.zeros((3, 11)), np.zeros((3, 11))
            hh.hh_one_electron_primitive(C.c_double(ex[0]), C.c_double(co[0]), R[0].ctypes, C.c_int(la),
                                         C.c_double(ex[1]), C.c_double(co[1]), R[1].ctypes, C.c_int(lb),
                                         C.c_int(3), Z.ctypes, Cn.ctypes, out.ctypes)
            lib.orc_one_electron_primitive_grad(C.c_double(ex[0]), C.c_double(co[0]), R[0].ctypes, C.c_int(la),
                                                C.c_double(ex[1]), C.c_double(co[1]), R[1].ctypes, C.c_int(lb),
                                                C.c_int(3), Z.ctypes, Cn.ctypes, ref.ctypes)
            assert not np.isnan(out).any()
            worst = max(worst, (np.abs(out - ref) / np.maximum(np.abs(ref), 1e-2)).max())
    assert worst < 1e-11, worst


def test_pair_underflow_marker(hh):
    """make_pair marks a Gaussian product whose exponential underflowed to exactly 0 with K = -0.0 (the kernels leave such
    primitive pairs out of their loops); a zero coefficient gives K = +0.0 and is NOT marked (its d/d(coef) path is alive);
    anything representable, however small, is computed."""
    A = np.zeros(3)
    K, u = C.c_double(0), C.c_int(0)

    def call(a, ca, b, cb, R):
        B = np.array([0.0, 0.0, R])
        hh.hh_pair_prefactor(C.c_double(a), C.c_double(ca), A.ctypes, C.c_double(b), C.c_double(cb), B.ctypes, C.byref(K), C.byref(u))
        return K.value, u.value
    k, m = call(6.44, 1.0, 3.43, 1.0, 24.0)            # mu R^2 = 2.24 * 576 = 1289 > 745: exp underflows
    assert k == 0.0 and np.signbit(k) and m == 1
    k, m = call(6.44, -1.0, 3.43, 1.0, 24.0)           # the marker does not depend on the coefficients' signs
    assert k == 0.0 and np.signbit(k) and m == 1
    k, m = call(0.38, 1.0, 0.17, 1.0, 24.0)            # mu R^2 = 67.6: tiny but representable
    assert 0.0 < k < 1e-25 and m == 0
    k, m = call(0.38, 0.0, 0.17, 1.0, 1.0)             # zero coefficient: value vanishes, not marked
    assert k == 0.0 and not np.signbit(k) and m == 0
    k, m = call(0.38, 0.0, 0.17, -1.0, 1.0)
    assert k == 0.0 and not np.signbit(k) and m == 0
