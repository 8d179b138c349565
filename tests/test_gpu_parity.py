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


@pytest.mark.parametrize("n,batch", [(1, 3), (2, 4), (3, 2), (7, 5), (16, 2), (33, 2), (48, 2), (49, 1), (64, 3), (101, 2), (224, 1), (236, 1)])
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


def test_tasks_api_energy_and_grad(golden):
    """The outer boundary exactly as a reference user drives it (examples/*.py)."""
    s = System_mol(synthetic.H2O, basis_set_3G_STO, 10, mol_name='agua', angs=False, shifted=False)
    m = Tasks(s, name='/tmp/dq_h2o_sto_3g', verbose=True)
    assert m.runtask('Energy', max_scf=50, printcoef=True, name='/tmp/Output.molden', output=False) is True
    assert abs(m.energy - (-74.9598578)) < 6e-8          # examples/h2o_optimization/h2_sto_3g.out:33
    assert abs(m.energy - float(golden["reference_energy"]["h2o_sto3g_E"])) < TOL_E
    assert m.runtask('Grad', max_scf=50, argnum=[0, 1, 2]) is True
    g = golden["reference_grad_h2o_sto3g"]
    for n in (0, 1, 2):
        assert np.abs(np.ravel(s.grad[n]) - g["g%d" % n]).max() < TOL_G
    m.end()
    assert 'SCF converged' in open('/tmp/dq_h2o_sto_3g.out').read()


def test_water_dimer_energy_and_gradient_vs_oracle():
    """A size the Python reference cannot reach quickly: (H2O)2 STO-3G, N = 14, 42 primitives."""
    mol = synthetic.water_cluster(2)
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


def test_two_rank_sharding_emulated_on_one_gpu():
    """The N>1 path on a single GPU: two host threads play rank 0 and rank 1 (each computes its half of the ERI tiles),
    the allreduce callback sums their device buffers through the host.  No kernel waits on another kernel: the host
    threads wait, with their streams idle.  Results must equal the single-rank run."""
    import threading
    import torch
    from diffiqult_b200 import distributed
    s = sysmol("ch2o")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    ref = Energy.rhf_grad(*args, max_scf=100)
    world = 2
    barrier = threading.Barrier(world)
    slots = [None] * world
    results = [None] * world

    def make_cb(rank):
        def _cb(ptr, count, user):
            t = torch.as_tensor(distributed._CudaBuffer(ptr, count), device="cuda:0")
            torch.cuda.synchronize()
            slots[rank] = t.cpu()
            barrier.wait()
            total = slots[0] + slots[1]
            barrier.wait()
            t.copy_(total.cuda())
            torch.cuda.synchronize()
        return _capi.ALLREDUCE_FN(_cb)

    def run(rank):
        cb = make_cb(rank)
        results[rank] = Energy.rhf_grad(*args, max_scf=100, world_size=world, rank=rank, allreduce=cb)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for r in results:
        assert r is not None and r["status"] == 0 and r["niter"] == ref["niter"]
        assert abs(r["E"] - ref["E"]) < 1e-10
        for n in range(3):
            assert np.abs(r["grad"][n] - ref["grad"][n]).max() < 1e-9
    assert results[0]["stats"]["n_tiles"] + results[1]["stats"]["n_tiles"] == ref["stats"]["n_tiles"]


def test_perturbed_methane_batch_vs_oracle():
    """BASELINE config 5 in miniature: 12 of the 1024 perturbed CH4 / STO-3G molecules, energy + full gradient through
    the threaded batch driver (per-thread streams), against the oracle's forward-mode gradient."""
    from diffiqult_b200 import batch
    mols = synthetic.perturbed_methanes(12)
    syss = [System_mol(m, basis_set_3G_STO, 10, "ch4_%d" % i) for i, m in enumerate(mols)]
    E, status, grads = batch.rhf_grad_batch(syss, max_scf=50, threads=4)
    assert (status == 0).all()
    for i in (0, 5, 11):
        a = orc.SysArrays(syss[i])
        o = orc.rhf(a, max_scf=50)
        assert abs(E[i] - o["E"]) < TOL_E
        for n in (0, 1, 2):
            og = orc.rhf_grad(a, n, max_scf=50)
            if og["status"] == 0:
                assert np.abs(grads[i][n] - og["grad"]).max() < TOL_G
    # serial evaluation gives the same numbers (thread safety of the library)
    r = Energy.rhf_grad(np.log(syss[3].alpha), syss[3].coef, syss[3].xyz, syss[3].l, syss[3].charges, syss[3].atom,
                        syss[3].list_contr, syss[3].ne, max_scf=50)
    assert abs(r["E"] - E[3]) < 1e-11 and np.abs(r["grad"][2] - grads[3][2]).max() < 1e-10
    assert list(batch.shard(1024, 8, 3)) == list(range(384, 512))


def test_full_size_cluster_properties():
    """(H2O)_32 / STO-3G, the bench workload, through size-independent properties: convergence and iteration count,
    directional derivative of the same truncated-SCF energy, and run-to-run reproducibility."""
    import bench
    s = bench.workload(32)
    args = [np.log(s.alpha), np.array(s.coef), np.array(s.xyz)]
    r = Energy.rhf_grad(*args, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    assert r["status"] == 0 and r["niter"] == 13
    assert abs(r["E"] - (-2398.5489189099)) < 1e-8
    r2 = Energy.rhf_grad(*args, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    assert abs(r["E"] - r2["E"]) < 1e-10 and max(np.abs(a - b).max() for a, b in zip(r["grad"], r2["grad"])) < 1e-9
    rng = np.random.default_rng(11)
    for n, h in ((0, 1e-4), (2, 1e-4)):
        d = rng.normal(size=args[n].shape)
        e = []
        for sgn in (+1, -1):
            a2 = [x.copy() for x in args]
            a2[n] = a2[n] + sgn * h * d
            rr = Energy.rhf(*a2, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
            assert rr["niter"] == r["niter"]
            e.append(rr["E"])
        fd = (e[0] - e[1]) / (2 * h)
        an = float(np.dot(r["grad"][n].ravel(), d.ravel()))
        assert abs(fd - an) < 1e-5 * max(1.0, abs(an)), (n, fd, an)


@pytest.mark.parametrize("name", ["h2_6g", "h2o_mixed"])
def test_long_and_mixed_contractions_vs_oracle(name):
    """Contraction lengths 6 (36 primitive pairs: the gradient kernel walks the ket primitives in chunks) and a basis
    mixing 1-, 2-, 3- and 4-primitive functions (function blocks split by contraction length)."""
    mol, basis, ne = systems.EXTRA_SYSTEMS[name]
    s = System_mol(mol, basis, ne, name)
    a = orc.SysArrays(s)
    cn = orc.normalization(a)
    E = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    assert np.abs(E - orc.eri(a, cn)).max() < TOL_INT
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    o = orc.rhf(a, max_scf=100)
    assert r["status"] == o["status"] == 0 and r["niter"] == o["niter"]
    assert abs(r["E"] - o["E"]) < TOL_E
    for n in (0, 1, 2):
        og = orc.rhf_grad(a, n, max_scf=100)
        assert og["status"] == 0
        assert np.abs(r["grad"][n] - og["grad"]).max() < TOL_G, (name, n)


def test_integral_direct_fock_build_matches_stored():
    """jk_mode='direct' (tiles recomputed batch by batch in every Fock build, Energy.py:29-38 without a stored Eri
    vector) against the stored-tile mode: (H2O)_8, 5565 tiles walked in batches of 256."""
    mol = synthetic.water_cluster(8)
    s = System_mol(mol, basis_set_3G_STO, 80, "w8")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    a = Energy.rhf_grad(*args, max_scf=100, jk_mode="stored")
    b = Energy.rhf_grad(*args, max_scf=100, jk_mode="direct", direct_batch_bytes=256 * 2048)
    assert a["stats"]["jk_direct"] == 0 and b["stats"]["jk_direct"] == 1
    assert a["status"] == b["status"] == 0 and a["niter"] == b["niter"]
    assert abs(a["E"] - b["E"]) < 1e-10
    assert np.abs(a["esteps"] - b["esteps"]).max() < 1e-10 if "esteps" in a else True
    for n in range(3):
        assert np.abs(a["grad"][n] - b["grad"][n]).max() < 1e-9
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    rng = np.random.default_rng(2)
    D = rng.normal(size=(s.nbasis, s.nbasis)); D = D + D.T
    F1, _ = Energy.fockmatrix(s.alpha, cn, s.xyz, s.l, s.charges, s.atom, s.list_contr, D, jk_mode="stored")
    F2, _ = Energy.fockmatrix(s.alpha, cn, s.xyz, s.l, s.charges, s.atom, s.list_contr, D, jk_mode="direct", direct_batch_bytes=1)
    assert np.abs(F1 - F2).max() < 1e-10 * max(1.0, np.abs(F1).max())


def test_bfgs_reproduces_reference_examples():
    """'Opt' end to end against the outputs the reference ships.
    examples/h2o_optimization/h2_sto_3g.out:61-77 -- H2O/STO-3G, argnum=[0], maxiter=3: energy, gradient norm and the seven
    printed exponents after 3 BFGS iterations (reproduced to every printed digit, incl. the reference's line-search quirk);
    examples/h2_optimization/h2_sto_3g.out:28-38 -- H2/STO-3G, argnum=[0,1] to convergence: E = -1.095971."""
    s = System_mol(synthetic.H2O, basis_set_3G_STO, 10, mol_name='agua')
    m = Tasks(s, name='/tmp/dq_opt_h2o', verbose=False)
    assert m.runtask('Energy', max_scf=50, printcoef=True, name='Output.molden', output=False)
    m.runtask('Opt', max_scf=50, printcoef=False, maxiter=3, argnum=[0], output=False)
    r = m.opt_result
    assert r.nit == 3 and r.message == 'Maximum number of iterations has been exceeded.'
    assert abs(r.fun - (-74.976643)) < 1e-6
    assert abs(np.abs(r.jac).max() - 0.256754) < 1e-6
    assert np.abs(s.alpha[:7] - [135.263104, 23.755473, 6.558207, 5.162051, 1.199706, 0.378480, 5.320693]).max() < 1e-6
    s = System_mol(synthetic.H2_EXAMPLE, basis_set_3G_STO, 2, mol_name='agua')
    m = Tasks(s, name='/tmp/dq_opt_h2', verbose=False)
    m.runtask('Opt', max_scf=50, printcoef=False, argnum=[0, 1], output=True)
    r = m.opt_result
    # The reference needed 26 BFGS iterations here (maxiter = 30).  Its line search (with the pk[0] quirk) amplifies
    # last-bit differences of E and the gradient - ours vary from run to run with the order of the floating-point atomics
    # - into a different iteration count (24-28 observed, occasionally the cap): the converged point is what is pinned.
    assert r.status in (0, 1) and r.nit <= 30
    assert abs(r.fun - (-1.095971)) < 1e-6 and np.abs(r.jac).max() < (1e-5 if r.status == 0 else 1e-3)
    assert np.abs(s.alpha[:2] - [4.803590, 0.694786]).max() < 5e-4
    # all three groups at once (crashes in the reference: Optimize.py:281-287)
    s = System_mol(synthetic.H2_EXAMPLE, basis_set_3G_STO, 2)
    m = Tasks(s, name='/tmp/dq_opt_h2_012', verbose=False)
    e0 = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=50)["E"]
    m.runtask('Opt', max_scf=50, argnum=[0, 1, 2], maxiter=4, reference_linesearch=False)
    assert m.opt_result.fun < e0 - 1e-3


def test_optin_screening_keeps_results_within_tolerance():
    """screen_threshold is off by default (the reference screens nothing).  Opt-in: tiles whose primitive prefactors are
    bounded below the threshold are skipped; energy and gradient stay within the parity tolerances."""
    mol = synthetic.water_cluster(8)
    s = System_mol(mol, basis_set_3G_STO, 80, "w8")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    a = Energy.rhf_grad(*args, max_scf=100)
    b = Energy.rhf_grad(*args, max_scf=100, screen_threshold=1e-18)
    assert a["stats"]["n_tiles_skipped"] == 0 and b["stats"]["n_tiles_skipped"] > a["stats"]["n_tiles"] // 2
    assert a["niter"] == b["niter"] and abs(a["E"] - b["E"]) < 1e-10
    for n in range(3):
        assert np.abs(a["grad"][n] - b["grad"][n]).max() < 1e-8


def test_molden_output_diffs_against_the_reference_file(golden, tmp_path):
    """Tasks.runtask('Energy', output=True) writes name-<ntask>.molden in the reference's layout; parsed with the parser
    that packed tests/output.molden (the reference's recorded CH4 run) it carries the same numbers."""
    g = golden["molden_ch4"]
    s = sysmol("ch4")
    m = Tasks(s, name=str(tmp_path / "ch4"), verbose=False)
    assert m.runtask('Energy', max_scf=100, output=True)
    lines = open(str(tmp_path / "ch4-1.molden")).read().split("\n")
    assert lines[0] == '[Energy] ' and lines[4] == 'SCF Details' and lines[5] == 'Eigen: True'
    assert abs(float(lines[1].split()[1]) - float(g["E_elec"])) < 2e-10
    assert abs(float(lines[2].split()[1]) - float(g["E_nuc"])) < 2e-9
    assert abs(float(lines[3].split()[1]) - float(g["E_tot"])) < 2e-10
    steps = [float(x.split()[2]) for x in lines if x.startswith("Step:")]
    assert len(steps) == 12 and np.abs(np.array(steps) - g["steps"]).max() < 2e-10
    for tag in ("[CM] ", "[DM] ", "[NMO] ", "[MOE] ", "[INPUT] ", "[Atoms]", "[GTO]", "[MO]"):
        assert tag in lines
    i = lines.index("[DM] ")
    D = np.array([[float(x) for x in ln.split()[1:]] for ln in lines[i + 3:i + 3 + 18]])
    assert np.abs(D - g["D"]).max() < 1e-8
    i = lines.index("[MOE] ")
    moe = np.array([float(ln.split()[1]) for ln in lines[i + 2:i + 20]])
    assert np.abs(moe - g["moe"]).max() < 1e-9
    i = lines.index("[GTO]")
    gto = [tuple(float(x) for x in ln.split()) for ln in lines[i + 1:lines.index("[MO]")] if len(ln.split()) == 2 and "." in ln.split()[0]]
    assert np.abs(np.array(gto) - g["gto"]).max() < 1e-11


def test_readguess_warm_start_vs_oracle(tmp_path):
    """printguess / readguess (Energy.py:148-157,197-198): the SCF restarted from a saved C, energy and gradient against
    the oracle run with the same guess (a slightly different geometry, so the guess is good but not converged)."""
    s = sysmol("h2o_sto3g")
    t = Tasks(s, str(tmp_path / "w"))
    args = t._energy_args(max_scf=50, printguess=str(tmp_path / "c.npy"))
    e0 = Energy.rhfenergy(*args)
    C0 = np.load(str(tmp_path / "c.npy"))
    assert C0.shape == (7, 7)
    s2 = sysmol("h2o_sto3g")
    s2.xyz = s2.xyz + 0.02 * np.random.default_rng(4).normal(size=s2.xyz.shape)
    a = orc.SysArrays(s2)
    args2 = (np.log(s2.alpha), s2.coef, s2.xyz, s2.l, s2.charges, s2.atom, s2.list_contr, s2.ne)
    cold = Energy.rhf_grad(*args2, max_scf=50)
    warm = Energy.rhf_grad(*args2, max_scf=50, guess_C=C0)
    assert warm["status"] == 0 and warm["niter"] < cold["niter"]
    orc.set_guess(C0)
    try:
        o = orc.rhf(a, max_scf=50)
        assert o["niter"] == warm["niter"] and abs(o["E"] - warm["E"]) < TOL_E
        for n in (0, 1, 2):
            og = orc.rhf_grad(a, n, max_scf=50)
            assert og["status"] == 0 and np.abs(og["grad"] - warm["grad"][n]).max() < TOL_G
    finally:
        orc.set_guess(None)
    # through the reference-shaped entry point: readguess is a path to the .npy
    t2 = Tasks(s2, str(tmp_path / "w2"))
    e = Energy.rhfenergy(*t2._energy_args(max_scf=50, readguess=str(tmp_path / "c.npy")))
    assert abs(e - warm["E"]) < 1e-12 and abs(e0 - e) > 1e-6


def test_real_multi_gpu_nccl_sharding():
    """One process per GPU over NCCL (torchrun, 2 ranks): sharded energy + gradient equal the single-GPU run
    (scripts/multi_gpu_check.py).  Skipped on a single-GPU box; the emulated test above covers the logic there."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "scripts", "multi_gpu_check.py"), "--waters", "3"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["world"] == 2 and rep["allreduce_calls"] > 0


def test_fused_batch_path_vs_oracle_and_per_molecule_path():
    """BASELINE config 5 through the fused device path (csrc/dq_batch.cu: thread per (quartet, molecule), one CTA per
    molecule for the SCF + reverse sweep): 24 perturbed CH4 / STO-3G.  Energies, iteration counts and all three
    gradient groups against the oracle (3 molecules) and against the per-molecule kernels (all molecules)."""
    from diffiqult_b200 import batch
    mols = synthetic.perturbed_methanes(24)
    syss = [System_mol(m, basis_set_3G_STO, 10, "ch4_%d" % i) for i, m in enumerate(mols)]
    E, status, grads, st = batch.rhf_grad_batch_fused(syss, max_scf=50)
    assert (status == 0).all() and st["launches"] < 40
    for i in (0, 7, 23):
        a = orc.SysArrays(syss[i])
        o = orc.rhf(a, max_scf=50)
        assert o["niter"] == st["niter"][i] and abs(E[i] - o["E"]) < TOL_E
        for n in (0, 1, 2):
            og = orc.rhf_grad(a, n, max_scf=50)
            assert og["status"] == 0 and np.abs(grads[i][n] - og["grad"]).max() < TOL_G
    for i, s in enumerate(syss):
        r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=50)
        assert r["niter"] == st["niter"][i] and abs(r["E"] - E[i]) < 1e-11
        for n in range(3):
            assert np.abs(r["grad"][n] - grads[i][n]).max() < 1e-10
    # energy-only mode, and the dispatcher picks the fused path for a one-structure batch
    E2, status2, g2 = batch.rhf_grad_batch(syss[:5], max_scf=50, with_grad=False)
    assert g2 is None and np.abs(E2 - E[:5]).max() < 1e-12


def test_fused_batch_other_structures_and_failure_sentinel():
    """The fused path on other shapes: formaldehyde (N = 12, mixed s/p classes incl. (pp|pp)), water with the uncontracted
    test basis (one primitive per function, N = 14), and a batch in which max_scf is too small (status 1 = 99999)."""
    from diffiqult_b200 import batch
    rng = np.random.default_rng(11)
    for name, max_scf in (("ch2o", 100), ("h2o", 60)):
        mol, basis, ne = systems.SYSTEMS[name]
        syss = []
        for r in range(3):
            m2 = [(z, tuple(np.array(p) + 0.03 * rng.normal(size=3))) for z, p in mol]
            syss.append(System_mol(m2, basis, ne, name))
        E, status, grads, st = batch.rhf_grad_batch_fused(syss, max_scf=max_scf)
        for i, s in enumerate(syss):
            a = orc.SysArrays(s)
            o = orc.rhf(a, max_scf=max_scf)
            assert status[i] == o["status"] and st["niter"][i] == o["niter"] and abs(E[i] - o["E"]) < TOL_E
            if o["status"] == 0:
                for n in (0, 2):
                    og = orc.rhf_grad(a, n, max_scf=max_scf)
                    assert np.abs(grads[i][n] - og["grad"]).max() < TOL_G
    syss = [System_mol(m, basis_set_3G_STO, 10, "ch4") for m in synthetic.perturbed_methanes(4)]
    E, status, grads, st = batch.rhf_grad_batch_fused(syss, max_scf=3)
    assert (status == 1).all() and (st["niter"] == 3).all()
    with pytest.raises(_capi.DiffiqultB200Error, match="nbasis <= 40"):
        big = [System_mol(synthetic.water_cluster(8), basis_set_3G_STO, 80, "w8")] * 2
        batch.rhf_grad_batch_fused(big, max_scf=5)


def _random_system(rng):
    """A random closed-shell 'molecule': 3-4 atoms (Z in 1..9), random (non-collinear: a linear molecule has exactly
    degenerate pi orbitals, and when the aufbau rule occupies one of such a pair in some SCF iteration the reference's
    result itself depends on an arbitrary eigenvector choice), a random s/p basis with contraction
    lengths 1-4 per function and random exponents / coefficients.  Nothing physical - the point is parity of the code
    paths (ragged blocks, mixed contraction lengths, every angular class) on inputs no fixture was tuned for."""
    nat = int(rng.integers(3, 5))
    zs = [int(z) for z in rng.integers(1, 10, size=nat)]
    if sum(zs) % 2:
        zs[0] += 1
    pos = rng.normal(scale=1.6, size=(nat, 3)) + np.arange(nat)[:, None] * np.array([1.3, 0.4, -0.7])
    mol = [(z, (float(p[0]), float(p[1]), float(p[2]))) for z, p in zip(zs, pos)]
    basis = {}
    for z in set(zs):
        shells = []
        for _ in range(int(rng.integers(1, 3))):
            k = int(rng.integers(1, 5))
            shells.append(('S', [(float(np.exp(rng.uniform(np.log(0.2), np.log(30.0)))), float(rng.uniform(0.2, 1.0))) for _ in range(k)]))
        if rng.random() < 0.7:
            k = int(rng.integers(1, 4))
            shells.append(('P', [(float(np.exp(rng.uniform(np.log(0.25), np.log(6.0)))), float(rng.uniform(0.2, 1.0))) for _ in range(k)]))
        basis[z] = shells
    return mol, basis, sum(zs)


@pytest.mark.parametrize("seed", [1, 2, 4, 5, 8, 10])
def test_random_systems_vs_oracle(seed):
    """Randomised parity: integrals (1e-10), the truncated-SCF energy with every per-iteration E_elec (1e-10, also when
    the SCF does NOT converge within max_scf = the reference's 99999 case) and the three gradient groups (1e-8)."""
    rng = np.random.default_rng(1000 + seed)
    mol, basis, nel = _random_system(rng)
    s = System_mol(mol, basis, nel, "rnd%d" % seed)
    if s.ne > s.nbasis or s.ne < 1:
        pytest.skip("more electron pairs than basis functions")
    a = orc.SysArrays(s)
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    S, T, V = Integrals.one_electron_matrices(s.alpha, cn, s.xyz, s.l, s.nbasis, s.charges, s.atom, s.list_contr)
    e = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    So, To, Vo = orc.one_electron(a, cn)
    assert max(np.abs(S - So).max(), np.abs(T - To).max(), np.abs(V - Vo).max() / max(1.0, np.abs(Vo).max())) < TOL_INT
    assert np.abs(e - orc.eri(a, cn)).max() < TOL_INT
    max_scf = 40
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=max_scf)
    o = orc.rhf(a, max_scf=max_scf)
    assert r["status"] == o["status"] and r["niter"] == o["niter"]
    scale = max(1.0, abs(o["E"]))
    assert abs(r["E"] - o["E"]) < TOL_E * scale
    if o["status"] != 0:
        return
    # Forward mode through eigh divides by EVERY eigenvalue difference (Eigen.py:69-74 / algopy's push-forward): with a
    # (near-)degenerate pair inside the occupied or the virtual space - e.g. the two pi orbitals of a linear molecule - the
    # reference's own gradient is rounding noise times 1/gap, although the energy does not depend on rotations inside those
    # spaces.  The reverse sweep here divides by occupied-virtual gaps only, so in that case the
    # oracle is not a usable reference and central finite differences of the GPU energy (same iteration count) are.
    eps = np.sort(np.asarray(o["eps"]))
    well_separated = np.diff(eps).min() > 1e-3
    if well_separated:
        for n in (0, 1, 2):
            og = orc.rhf_grad(a, n, max_scf=max_scf)
            assert np.abs(r["grad"][n] - og["grad"]).max() < TOL_G * max(1.0, np.abs(og["grad"]).max())
    base = [np.log(s.alpha), np.array(s.coef, dtype=float), np.array(s.xyz, dtype=float)]
    h = 2e-4
    for n in (0, 1, 2):
        g = np.ravel(r["grad"][n])
        picks = {int(np.abs(g).argmax()), int(rng.integers(g.size)), int(rng.integers(g.size))}
        for idx in picks:
            e_pm = []
            for sign in (1.0, -1.0):
                args = [b.copy() for b in base]
                flat = args[n].reshape(-1)
                flat[idx] += sign * h
                q = Energy.rhf(args[0], args[1], args[2], s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=max_scf)
                e_pm.append(q["E"] if (q["status"] == 0 and q["niter"] == r["niter"]) else None)
            if None in e_pm:
                continue        # the stopping rule fired at a different iteration: E_k is not the same function there
            fd = (e_pm[0] - e_pm[1]) / (2 * h)
            assert abs(fd - g[idx]) < 5e-6 * max(1.0, np.abs(g).max()), (n, idx, fd, g[idx])


def test_sparse_weight_gradient_kernel_equals_dense(monkeypatch):
    """The lane-independent ERI-gradient kernel (chosen automatically when the density matrix is mostly exact zeros) forced
    on systems with dense weights: same gradient as the tile-synchronous kernel, and both against the oracle; and the
    automatic choice on two far-apart water monomers, whose density is exactly block diagonal."""
    for name in ("ch2o", "h2o_sto3g", "ch4"):
        s = sysmol(name)
        args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
        monkeypatch.setenv("DQ_GRAD_KERNEL", "dense")
        d = Energy.rhf_grad(*args, max_scf=100)
        monkeypatch.setenv("DQ_GRAD_KERNEL", "sparse")
        sp = Energy.rhf_grad(*args, max_scf=100)
        assert d["stats"]["grad_kernel_sparse"] == 0 and sp["stats"]["grad_kernel_sparse"] == 1
        for n in range(3):
            assert np.abs(d["grad"][n] - sp["grad"][n]).max() < 1e-11 * max(1.0, np.abs(d["grad"][n]).max())
    monkeypatch.delenv("DQ_GRAD_KERNEL")
    # N = 56 >= 48: the cluster eigensolver, which keeps the exact zeros between the eight far-apart monomers
    mol = synthetic.water_cluster(8, spacing=40.0)
    s = System_mol(mol, basis_set_3G_STO, 80, "w8far")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    r = Energy.rhf_grad(*args, max_scf=100)
    assert r["status"] == 0 and r["stats"]["grad_kernel_sparse"] == 1
    assert r["stats"]["n_grad_quartets_zero_weight"] > 0.3 * r["stats"]["n_quartets"]     # 37 % for 8 monomers, 80 % for 32
    monkeypatch.setenv("DQ_GRAD_KERNEL", "dense")
    d = Energy.rhf_grad(*args, max_scf=100)
    assert d["stats"]["grad_kernel_sparse"] == 0 and abs(d["E"] - r["E"]) < 1e-11
    for n in range(3):
        assert np.abs(d["grad"][n] - r["grad"][n]).max() < 1e-11 * max(1.0, np.abs(d["grad"][n]).max())
    # every monomer is a rigidly rotated copy plus jitter: the centre gradient of the whole cluster sums to ~0 per monomer
    # only up to the jitter, but translational invariance of the TOTAL energy holds exactly for floating Gaussians + nuclei
    # moved together; here only the Gaussians move, so no such identity applies - the dense kernel is the reference.


def test_exact_zero_primitive_pairs_are_skipped_bit_identically(monkeypatch):
    """Primitive pairs whose Gaussian product underflows to exactly 0 are left out of the kernels' loops (marker K = -0.0):
    same energy and gradient to the last bits as with DQ_NO_ZERO_SKIP=1, on monomers far enough apart for underflow; and a
    primitive with coefficient exactly 0 (value contributions vanish, d/d(coef) does not) still gets its gradient."""
    mol = synthetic.water_cluster(8, spacing=40.0)
    s = System_mol(mol, basis_set_3G_STO, 80, "w8far")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    a = Energy.rhf_grad(*args, max_scf=100)
    assert a["stats"]["n_prim_pairs_underflowed"] > 0.3 * a["stats"]["n_prim_pairs"]
    monkeypatch.setenv("DQ_NO_ZERO_SKIP", "1")
    b = Energy.rhf_grad(*args, max_scf=100)
    monkeypatch.delenv("DQ_NO_ZERO_SKIP")
    assert a["niter"] == b["niter"] and abs(a["E"] - b["E"]) < 1e-11
    for n in range(3):
        assert np.abs(a["grad"][n] - b["grad"][n]).max() < 1e-12
    # zero coefficient: CH2O with the middle primitive of the first function switched off
    s = sysmol("ch2o")
    coef = np.array(s.coef, dtype=float)
    coef[1] = 0.0
    r = Energy.rhf_grad(np.log(s.alpha), coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    s.coef = coef
    ao = orc.SysArrays(s)
    o = orc.rhf(ao, max_scf=100)
    assert r["niter"] == o["niter"] and abs(r["E"] - o["E"]) < TOL_E
    og = orc.rhf_grad(ao, 1, max_scf=100)
    assert abs(og["grad"][1]) > 1e-3 and np.abs(r["grad"][1] - og["grad"]).max() < TOL_G


@pytest.mark.parametrize("name", ["h2", "h2o_sto3g", "ch2o", "ch4", "h2_6g", "h4_6g", "h2o_mixed"])
def test_compacting_eri_kernel_equals_plain_and_oracle(monkeypatch, name):
    """The ERI kernel that lists the surviving primitive pairs of a tile and walks their Cartesian product (chosen
    automatically when many Gaussian products underflow) forced on ordinary molecules, incl. long (36 primitive pairs:
    several list chunks) and mixed contractions: packed ERI vector against the plain kernel and the oracle."""
    table = systems.SYSTEMS if name in systems.SYSTEMS else systems.EXTRA_SYSTEMS
    mol, basis, ne = table[name]
    s = System_mol(mol, basis, ne, name)
    a = orc.SysArrays(s)
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    monkeypatch.setenv("DQ_ERI_KERNEL", "tiles")
    e0 = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    monkeypatch.setenv("DQ_ERI_KERNEL", "compact")
    e1 = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    assert np.abs(e1 - e0).max() < 1e-13 * max(1.0, np.abs(e0).max())
    assert np.abs(e1 - orc.eri(a, cn)).max() < TOL_INT
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=60)
    assert r["stats"]["eri_kernel_compact"] == 1
    o = orc.rhf(a, max_scf=60)
    assert r["niter"] == o["niter"] and abs(r["E"] - o["E"]) < TOL_E
