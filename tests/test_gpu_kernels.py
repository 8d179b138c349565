"""GPU parity of the kernel VARIANTS (run before the core-path and workflow files: pytest orders files by name):
the sparse-weight gradient kernel, exact-zero skipping, the compacting ERI kernel, the fused batch path, long / mixed
contractions, integral-direct Fock builds, opt-in screening and randomised systems, against the CPU oracle."""
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


@pytest.mark.parametrize("name", ["h2o_sto3g", "ch2o", "ch4_sto3g", "hcn_sto3g", "w2", "w4"])
def test_deferred_loop_regime_eri_kernel_equals_plain_and_oracle(monkeypatch, name):
    """k_eri_tiles_def (the default for STO-3G-like contractions: loop-regime primitive quartets of a tile marked in phase A,
    listed by Boys-argument bin and evaluated by whichever thread comes next in phase B, summed in fixed point) against the
    plain kernel (DQ_ERI_KERNEL=plain: Boys loops in place) and the oracle; molecules put most quartets in the loop regime,
    the water dimer / tetramer of the dense cluster mix all regimes.  Two calls return identical bits."""
    if name.startswith("w"):
        mol, basis, ne = synthetic.water_cluster(int(name[1:]), spacing=5.67), basis_set_3G_STO, 10 * int(name[1:])
    else:
        mol, basis, ne = (systems.SYSTEMS if name in systems.SYSTEMS else systems.EXTRA_SYSTEMS)[name]
    s = System_mol(mol, basis, ne, name)
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    monkeypatch.setenv("DQ_ERI_KERNEL", "plain")
    e0 = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    monkeypatch.delenv("DQ_ERI_KERNEL")
    e1 = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    e2 = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    assert np.array_equal(e1, e2)
    assert np.abs(e1 - e0).max() < 1e-13 * max(1.0, np.abs(e0).max())
    if s.nbasis <= 14:
        assert np.abs(e1 - orc.eri(orc.SysArrays(s), cn)).max() < TOL_INT


@pytest.mark.parametrize("name", ["h2o_sto3g", "ch2o", "hcn_sto3g", "w2", "w4"])
def test_deferred_loop_regime_gradient_kernels_equal_in_place_form(monkeypatch, name):
    """The gradient pass in two kernels (default for STO-3G-like contractions and classes with >= 2 p functions:
    k_eri_grad_chunks<DEFER> marks the loop-regime primitive quartets, k_eri_grad_deferred evaluates them per tile, listed by
    Boys-argument bin, and delivers their 14 adjoint sums through fixed-point shared-memory accumulators) against the
    one-kernel form with the Boys loops in place (DQ_GRAD_DEFER=0): energy and all three gradient groups; where the oracle is
    cheap, against it as well."""
    if name.startswith("w"):
        mol, basis, ne = synthetic.water_cluster(int(name[1:]), spacing=5.67), basis_set_3G_STO, 10 * int(name[1:])
    else:
        mol, basis, ne = (systems.SYSTEMS if name in systems.SYSTEMS else systems.EXTRA_SYSTEMS)[name]
    s = System_mol(mol, basis, ne, name)
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    monkeypatch.setenv("DQ_GRAD_DEFER", "0")
    r0 = Energy.rhf_grad(*args, max_scf=200)
    monkeypatch.delenv("DQ_GRAD_DEFER")
    r1 = Energy.rhf_grad(*args, max_scf=200)
    assert r0["status"] == r1["status"] == 0 and r0["niter"] == r1["niter"] and r0["E"] == r1["E"]
    for n in (0, 1, 2):
        scale = max(1.0, np.abs(r0["grad"][n]).max())
        assert np.abs(r1["grad"][n] - r0["grad"][n]).max() < 1e-12 * scale, n
    if name in ("h2o_sto3g", "ch2o"):
        a = orc.SysArrays(s)
        for n in (0, 1, 2):
            og = orc.rhf_grad(a, n, max_scf=200)
            assert og["status"] == 0 and np.abs(og["grad"] - np.ravel(r1["grad"][n])).max() < TOL_G, n


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
        big = [System_mol(synthetic.water_cluster(8, spacing=24.0), basis_set_3G_STO, 80, "w8")] * 2
        batch.rhf_grad_batch_fused(big, max_scf=5)


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
    mol = synthetic.water_cluster(8, spacing=24.0)
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


def test_optin_screening_keeps_results_within_tolerance():
    """screen_threshold is off by default (the reference screens nothing).  Opt-in: tiles whose primitive prefactors are
    bounded below the threshold are skipped; energy and gradient stay within the parity tolerances."""
    mol = synthetic.water_cluster(8, spacing=24.0)
    s = System_mol(mol, basis_set_3G_STO, 80, "w8")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    a = Energy.rhf_grad(*args, max_scf=100)
    b = Energy.rhf_grad(*args, max_scf=100, screen_threshold=1e-18)
    assert a["stats"]["n_tiles_skipped"] == 0 and b["stats"]["n_tiles_skipped"] > a["stats"]["n_tiles"] // 2
    assert a["niter"] == b["niter"] and abs(a["E"] - b["E"]) < 1e-10
    for n in range(3):
        assert np.abs(a["grad"][n] - b["grad"][n]).max() < 1e-8


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


def _cluster(name, golden):
    """(System_mol, argument tuple, golden) of one oracle_cluster_* case (tests/golden/make_cluster_golden.py)."""
    g = golden["oracle_cluster_" + name]
    n = int(g["n_waters"])
    mol = synthetic.water_cluster(n, spacing=float(g["spacing"]))
    s = System_mol(mol, basis_set_3G_STO, 10 * n, name)
    assert abs(float(np.sum(s.xyz)) - float(g["xyz_sum"])) < 1e-9          # same synthetic geometry as the generator's
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    return s, mol, args, g


CLUSTER_VARIANTS = [  # (DQ_ERI_KERNEL, DQ_GRAD_KERNEL, DQ_NO_ZERO_SKIP): every combination of the kernels the bench can select
    (None, None, None), ("tiles", "dense", None), ("compact", "sparse", None), ("tiles", "sparse", "1"), ("compact", "dense", None)]


@pytest.mark.parametrize("name", ["w8", "w8core", "w8far"])
def test_water_octamer_energy_steps_and_gradients_vs_oracle_every_kernel_variant(golden, monkeypatch, name):
    """(H2O)_8 / STO-3G, N = 56 - the smallest size at which the big-system code paths are live (cluster eigensolver inside
    the SCF, warm starts, multi-chunk J/K) - against the CPU oracle (tests/golden/oracle_cluster_*.npz): energy, iteration
    count, EVERY per-iteration E_elec (1e-10) and all three gradient groups (1e-8), with the plain / compacting ERI kernel
    and the dense / sparse gradient kernel forced both ways.
      w8     5.67 bohr (SURVEY 8d density), started from the monomer-superposition readguess (Energy.py:148-157)
      w8core 24 bohr, core guess          w8far  40 bohr, core guess (underflowed pairs, exactly block-sparse density)"""
    if not golden.has("oracle_cluster_" + name):
        pytest.skip("golden not generated")
    s, mol, args, g = _cluster(name, golden)
    guess = g["C0"] if "C0" in g.files else None
    for eri_k, grad_k, nozs in CLUSTER_VARIANTS:
        for var, val in (("DQ_ERI_KERNEL", eri_k), ("DQ_GRAD_KERNEL", grad_k), ("DQ_NO_ZERO_SKIP", nozs)):
            if val is None:
                monkeypatch.delenv(var, raising=False)
            else:
                monkeypatch.setenv(var, val)
        r = Energy.rhf_grad(*args, max_scf=100, guess_C=guess)
        e = Energy.rhf(*args, max_scf=100, guess_C=guess)
        tag = (name, eri_k, grad_k, nozs)
        assert r["status"] == 0 and r["niter"] == int(g["niter"]), tag
        assert abs(r["E"] - float(g["E"])) < TOL_E, tag
        assert e["niter"] == int(g["niter"]) and np.abs(e["esteps"] - g["steps"]).max() < TOL_E, tag
        for n in (0, 1, 2):
            assert np.abs(r["grad"][n] - g["g%d" % n]).max() < TOL_G, tag + (n,)
        if eri_k is not None:
            assert r["stats"]["eri_kernel_compact"] == (1 if eri_k == "compact" else 0)
            assert r["stats"]["grad_kernel_sparse"] == (1 if grad_k == "sparse" else 0)
    if guess is not None:          # the guess the GPU path builds itself (monomer SCFs on the device) spans the same density
        C0 = synthetic.monomer_guess(mol, basis_set_3G_STO)
        ne = s.ne
        assert np.abs(C0[:, :ne] @ C0[:, :ne].T - guess[:, :ne] @ guess[:, :ne].T).max() < 1e-7


def test_two_calls_return_identical_bits_with_fixed_point_accumulation(golden, monkeypatch):
    """J, K, pair adjoints and gradients are summed in integer limbs (dq_options.accumulate, the default), so
    the result does not depend on the order in which warps finish: two consecutive calls on (H2O)_8 return bit-identical
    energies, step energies and gradients (stored and direct Fock builds, dense and sparse gradient kernels); the
    floating-point-atomics mode agrees with it to rounding."""
    mol = synthetic.water_cluster(8, spacing=5.67)
    s = System_mol(mol, basis_set_3G_STO, 80, "w8")
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    C0 = synthetic.monomer_guess(mol, basis_set_3G_STO)
    for grad_k in ("dense", "sparse"):
        monkeypatch.setenv("DQ_GRAD_KERNEL", grad_k)
        runs = [Energy.rhf_grad(*args, max_scf=100, guess_C=C0) for _ in range(3)]
        for r in runs[1:]:
            assert r["E"] == runs[0]["E"] and r["niter"] == runs[0]["niter"]
            for n in range(3):
                assert np.array_equal(r["grad"][n], runs[0]["grad"][n])
    monkeypatch.delenv("DQ_GRAD_KERNEL")
    e1 = Energy.rhf(*args, max_scf=100, guess_C=C0)
    e2 = Energy.rhf(*args, max_scf=100, guess_C=C0)
    assert np.array_equal(e1["esteps"], e2["esteps"]) and np.array_equal(e1["C"], e2["C"])
    fl = Energy.rhf_grad(*args, max_scf=100, guess_C=C0, accumulate="float")
    assert fl["niter"] == runs[0]["niter"] and abs(fl["E"] - runs[0]["E"]) < 1e-11
    for n in range(3):
        assert np.abs(fl["grad"][n] - runs[0]["grad"][n]).max() < 1e-11
    # the fused batch path as well
    from diffiqult_b200 import batch
    syss = [System_mol(m, basis_set_3G_STO, 10, "ch4") for m in synthetic.perturbed_methanes(6)]
    a = batch.rhf_grad_batch_fused(syss, max_scf=50)
    b = batch.rhf_grad_batch_fused(syss, max_scf=50)
    assert np.array_equal(a[0], b[0])
    for ga, gb in zip(a[2], b[2]):
        for n in range(3):
            assert np.array_equal(ga[n], gb[n])


def test_error_statuses_instead_of_process_exit():
    """The reference exits the process when a Boys loop does not converge (Tools.py:124-125,150-151) and fails with a
    NameError for max_scf < 1 (Energy.py:164,196); the C ABI returns a status instead."""
    t = np.array([0.5, np.nan, 3.0])
    F, dF = np.zeros((3, 5)), np.zeros((3, 5))
    st = _capi.lib().dq_boys(C.c_int32(3), t.ctypes, F.ctypes, dF.ctypes, C.c_int32(0))
    assert st == -5 and b"Boys" in _capi.lib().dq_last_error()
    st = _capi.lib().dq_boys(C.c_int32(1), t.ctypes, F.ctypes, dF.ctypes, C.c_int32(0))      # the flag was cleared
    assert st == 0
    s = sysmol("h2o_sto3g")
    with pytest.raises(ValueError):
        Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=0)
    a = _capi.SystemArrays(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, True)
    opt = _capi.options(max_scf=0)
    E, nit = C.c_double(0.0), C.c_int32(0)
    assert _capi.lib().dq_rhf(C.byref(a.struct), C.byref(opt), C.byref(E), C.byref(nit), None, None, None) == -1
