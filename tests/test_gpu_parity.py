"""GPU parity tests: the CUDA path (through the C ABI / the reference-shaped Python API) against the CPU oracle
and the golden vectors.  Tolerances are the north-star's: 1e-10 max-abs on integral matrices, 1e-10 Eh on
energies, 1e-8 on gradients."""
import ctypes as C
import os

import numpy as np
import pytest

import systems
from diffiqult_b200 import System_mol, Tasks, synthetic
from diffiqult_b200 import Energy, Integrals, _capi
from diffiqult_b200.Basis import basis_set_3G_STO
from oracle import cpu_oracle as orc

pytestmark = pytest.mark.gpu

TOL_INT, TOL_E, TOL_G = 1e-10, 1e-10, 1e-8


def sysmol(name):
    mol, basis, ne = systems.SYSTEMS[name]
    return System_mol(mol, basis, ne, name)


def test_library_reports_a_device():
    assert _capi.lib().dq_device_count() >= 1


def test_boys_function_and_tangent():
    t = np.concatenate([[0.0, 1e-14, 1e-12, 1e-10, 1e-6], np.linspace(1e-4, 12.0, 6001), np.geomspace(12.0, 1e5, 2000),
                        np.array([1.5, 2.5, 3.5, 4.5, 5.5])[:, None].repeat(5, 1).ravel() + np.tile([-1e-6, -1e-12, 0, 1e-12, 1e-6], 5)])
    F, dF = np.zeros((t.size, 5)), np.zeros((t.size, 5))
    _capi.check(_capi.lib().dq_boys(C.c_int32(t.size), t.ctypes, F.ctypes, dF.ctypes, C.c_int32(0)))
    for m in range(5):
        Fo, dFo = orc.boys(t, m)
        assert np.abs(F[:, m] - Fo).max() < 5e-14
        bound = 1e-13 + 8 * 2.2e-16 * (m + 0.5) / np.maximum(t, 1e-12) * np.abs(Fo)   # see test_kernel_math_host.py
        assert (np.abs(dF[:, m] - dFo) < bound).all()


@pytest.mark.parametrize("name", list(systems.SYSTEMS))
def test_integrals_vs_oracle_and_reference_run(golden, name):
    s = sysmol(name)
    a = orc.SysArrays(s)
    g = golden["reference_integrals"]
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    assert np.abs(cn - orc.normalization(a)).max() < 1e-13
    assert np.abs(cn - g[name + "_coefn"]).max() < 1e-13
    S, T, V = Integrals.one_electron_matrices(s.alpha, cn, s.xyz, s.l, s.nbasis, s.charges, s.atom, s.list_contr)
    So, To, Vo = orc.one_electron(a, cn)
    for got, want, ref in ((S, So, g[name + "_S"]), (T, To, g[name + "_T"]), (V, Vo, g[name + "_V"])):
        assert np.abs(got - want).max() < TOL_INT
        assert np.abs(got - ref).max() < TOL_INT
    E = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    assert E.shape == g[name + "_ERI"].shape
    assert np.abs(E - g[name + "_ERI"]).max() < TOL_INT
    assert np.abs(E - orc.eri(a, cn)).max() < TOL_INT


def test_reference_named_integral_functions_agree(golden):
    s = sysmol("ch2o")
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    g = golden["reference_integrals"]
    assert np.abs(Integrals.overlapmatrix(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr, np.float64(1.0)) - g["ch2o_S"]).max() < TOL_INT
    assert np.abs(Integrals.kineticmatrix(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr, np.float64(1.0)) - g["ch2o_T"]).max() < TOL_INT
    assert np.abs(Integrals.nuclearmatrix(s.alpha, cn, s.xyz, s.l, s.nbasis, s.charges, s.atom, s.natoms, s.list_contr,
                                          np.float64(1.0)) - g["ch2o_V"]).max() < TOL_INT


@pytest.mark.parametrize("name", systems.FIXTURE_NAMES)
def test_integrals_vs_pyquante_fixtures(golden, name):
    """The reference's own unit tests (tests/Test.py:214-236) assert places=5 against these dumps."""
    g = golden["pyquante_fixtures"]
    s = sysmol(name)
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    S, T, V = Integrals.one_electron_matrices(s.alpha, cn, s.xyz, s.l, s.nbasis, s.charges, s.atom, s.list_contr)
    assert np.abs(S - g[name + "_S"]).max() < 1e-9
    assert np.abs(T - g[name + "_T"]).max() < 1e-9
    assert np.abs(V - g[name + "_V"]).max() < 2e-7
    if name + "_ERI" in g.files:
        E = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
        assert np.abs(E - g[name + "_ERI"]).max() < 2e-8


@pytest.mark.parametrize("name", ["h2o", "ch2o", "hcn_sto3g"])
def test_fock_build_vs_oracle(name):
    s = sysmol(name)
    a = orc.SysArrays(s)
    cn = orc.normalization(a)
    rng = np.random.default_rng(5)
    D = rng.normal(size=(a.nbasis, a.nbasis)); D = D + D.T
    F, H = Energy.fockmatrix(s.alpha, cn, s.xyz, s.l, s.charges, s.atom, s.list_contr, D)
    So, To, Vo = orc.one_electron(a, cn)
    Fo = orc.fock(To + Vo, orc.eri(a, cn), D)
    assert np.abs(H - (To + Vo)).max() < TOL_INT
    assert np.abs(F - Fo).max() < 1e-10 * max(1.0, np.abs(Fo).max())


@pytest.mark.parametrize("n,batch", [(1, 3), (2, 4), (3, 2), (7, 5), (16, 2), (33, 2), (48, 2), (49, 1), (64, 3), (101, 2), (224, 1), (236, 1), (260, 1)])
def test_eigensolver(n, batch):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(batch, n, n)); A = A + A.transpose(0, 2, 1)
    if n >= 7:   # repeated eigenvalues in the first matrix
        q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        w = rng.normal(size=n); w[1] = w[0]; w[3] = w[2]
        A[0] = (q * w) @ q.T
    A = np.ascontiguousarray(A)
    w, V = np.zeros((batch, n)), np.zeros((batch, n, n))
    _capi.check(_capi.lib().dq_eigh(C.c_int32(n), C.c_int32(batch), A.ctypes, w.ctypes, V.ctypes, C.c_int32(0)))
    for b in range(batch):
        wn = np.linalg.eigvalsh(A[b])
        scale = max(1.0, np.abs(wn).max())
        assert np.abs(w[b] - wn).max() < 1e-12 * scale * max(1, n / 16)
        assert np.all(np.diff(w[b]) >= 0)
        assert np.abs(V[b].T @ V[b] - np.eye(n)).max() < 1e-12
        assert np.abs(A[b] @ V[b] - V[b] * w[b]).max() < 1e-11 * scale * max(1, n / 16)


@pytest.mark.parametrize("name", list(systems.SYSTEMS))
def test_rhf_energy_vs_reference_run(golden, name):
    """E, every per-iteration E_elec and the iteration count of the reference's own code (make_golden.py)."""
    g = golden["reference_energy"]
    s = sysmol(name)
    r = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    if float(g[name + "_E"]) == 99999.0:   # the reference did not converge in 100 iterations (hcn test basis)
        assert r["status"] == 1 and r["niter"] == 100
        assert abs(r["E"] - float(g[name + "_steps"][0])) < 1e-6 * abs(r["E"])
        return
    assert r["status"] == 0
    assert r["niter"] == len(g[name + "_steps"])
    assert np.abs(r["esteps"] - g[name + "_steps"]).max() < TOL_E
    assert abs(r["E"] - float(g[name + "_E"])) < TOL_E
    assert np.abs(r["eps"] - g[name + "_moe"]).max() < 1e-8
    C1, C2 = r["C"][:, :s.ne], g[name + "_C"][:, :s.ne]
    assert np.abs(C1 @ C1.T - C2 @ C2.T).max() < 1e-7


def test_ch4_recorded_run(golden):
    """tests/output.molden of the reference: E_tot -38.6835839813 and all 12 SCF step energies."""
    g = golden["molden_ch4"]
    s = sysmol("ch4")
    r = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    assert r["status"] == 0 and r["niter"] == 12
    assert abs(r["E"] - float(g["E_tot"])) < 2e-10
    assert np.abs(r["esteps"] - g["steps"]).max() < 2e-10
    assert np.abs(r["eps"] - g["moe"]).max() < 1e-9


def test_scf_failure_sentinel():
    s = sysmol("h2o_sto3g")
    t = Tasks(s, "/tmp/dq_fail")
    args = t._energy_args(max_scf=3)
    assert Energy.rhfenergy(*args) == 99999
    assert t.runtask("Energy", max_scf=3) is False


@pytest.mark.parametrize("name", ["h2_sto3g", "h2", "h2o_sto3g", "ch4_sto3g_p0", "ch2o", "h2o"])
def test_gradient_vs_reference_run(golden, name):
    """dE/d(ln alpha), dE/d(raw coef), dE/d(centres) vs the reference's Task.function_grads_algopy run here."""
    if not golden.has("reference_grad_" + name):
        pytest.skip("golden not generated")
    g = golden["reference_grad_" + name]
    s = sysmol(name)
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=int(g["max_scf"]))
    assert r["status"] == 0
    for n in (0, 1, 2):
        if "g%d" % n in g.files:
            assert np.abs(r["grad"][n] - g["g%d" % n]).max() < TOL_G, (name, n)


@pytest.mark.parametrize("name", ["ch4_sto3g", "ch4_sto3g_p1"])
def test_gradient_vs_oracle_forward_mode(name):
    """Systems with symmetry-degenerate orbitals (Td CH4) included: the adjoint path only divides by
    occupied-virtual gaps, the oracle's naive eigenvector tangent may refuse (status 4) and is then skipped."""
    s = sysmol(name)
    a = orc.SysArrays(s)
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    assert r["status"] == 0
    checked = 0
    for n in (0, 1, 2):
        o = orc.rhf_grad(a, n, max_scf=100)
        if o["status"] == 4:
            continue
        assert o["status"] == 0 and o["niter"] == r["niter"]
        assert abs(o["E"] - r["E"]) < TOL_E
        assert np.abs(r["grad"][n] - o["grad"]).max() < TOL_G, (name, n)
        checked += 1
    assert checked >= 1 or name != "ch4_sto3g_p1"


def test_gradient_vs_finite_differences_of_gpu_energy():
    """Size-independent property: directional derivative of the SAME truncated-SCF energy."""
    s = sysmol("ch4_sto3g_p1")
    args = (np.log(s.alpha), s.coef, s.xyz)
    r = Energy.rhf_grad(*args, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    rng = np.random.default_rng(3)
    for n in (0, 1, 2):
        d = rng.normal(size=np.shape(args[n]))
        h = 2e-4
        e = []
        for sgn in (+1, -1):
            a2 = [np.array(x, dtype=float) for x in args]
            a2[n] = a2[n] + sgn * h * d
            rr = Energy.rhf(*a2, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
            assert rr["niter"] == r["niter"]
            e.append(rr["E"])
        fd = (e[0] - e[1]) / (2 * h)
        assert abs(fd - float(np.dot(r["grad"][n].ravel(), d.ravel()))) < 2e-6


def test_water_dimer_energy_and_gradient_vs_oracle():
    """A size the Python reference cannot reach quickly: (H2O)2 STO-3G, N = 14, 42 primitives."""
    mol = synthetic.water_cluster(2, spacing=24.0)
    s = System_mol(mol, basis_set_3G_STO, 20, "w2")
    a = orc.SysArrays(s)
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    o = orc.rhf(a, max_scf=100)
    assert r["status"] == 0 and o["status"] == 0 and r["niter"] == o["niter"]
    assert abs(r["E"] - o["E"]) < TOL_E
    for n in (0, 2):
        og = orc.rhf_grad(a, n, max_scf=100)
        assert og["status"] == 0
        assert np.abs(r["grad"][n] - og["grad"]).max() < TOL_G
