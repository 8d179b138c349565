"""Pin the CPU oracle (oracle/rhf_oracle.cpp) to the reference: its PyQuante fixtures, its recorded CH4
run, and outputs of the reference's own source run under Python 3 (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import systems
from diffiqult_b200.Molecule import System_mol
from oracle import cpu_oracle as orc


def arrays(name):
    mol, basis, ne = systems.SYSTEMS[name]
    return orc.SysArrays(System_mol(mol, basis, ne, name))


def test_inputs_match_reference_system_builder(golden):
    """Our System_mol flattens (mol, basis) exactly as the reference's Molecule.py does."""
    g = golden["reference_integrals"]
    for name in systems.SYSTEMS:
        a = arrays(name)
        assert np.array_equal(a.alpha, g[name + "_in_alpha"])
        assert np.array_equal(a.coef, g[name + "_in_coef"])
        assert np.array_equal(a.xyz, g[name + "_in_xyz"])
        assert np.array_equal(a.ltype, g[name + "_in_ltype"])
        assert np.array_equal(a.contr, g[name + "_in_contr"])
        assert np.array_equal(a.charges, g[name + "_in_charges"])
        assert np.array_equal(a.atoms, g[name + "_in_atoms"])
        assert a.ne == int(g[name + "_in_ne"])


def test_boys_matches_reference_source(golden):
    """F_m(t) and its tangent, incl. the gser/gcf branch points and the 1e-12 clamp (Tools.py:78-151)."""
    g = golden["reference_boys"]
    for m in range(5):
        F, dF = orc.boys(g["t"], m)
        assert np.abs(F - g["F"][m]).max() <= 1e-14 * max(1.0, np.abs(g["F"][m]).max())
        assert np.abs(dF - g["dF"][m]).max() <= 1e-13 * max(1.0, np.abs(g["dF"][m]).max())


def test_boys_vs_scipy_places5():
    """The reference's own unit test (tests/Test.py:185-200): getgammp vs scipy, places=5."""
    from scipy.special import gammainc, gamma
    t = np.arange(1.0e-12, 9, 0.01)
    for m in range(5):
        F, _ = orc.boys(t, m)
        a = m + 0.5
        exact = 0.5 * gammainc(a, t) * gamma(a) * t ** (-a)
        assert np.abs(F - exact).max() < 0.5e-5


@pytest.mark.parametrize("name", list(systems.SYSTEMS))
def test_integrals_match_reference_source(golden, name):
    """Normalisation, S, T, V and the packed ERI vector vs the reference's own code: 1e-12."""
    g = golden["reference_integrals"]
    a = arrays(name)
    cn = orc.normalization(a)
    assert np.abs(cn - g[name + "_coefn"]).max() < 1e-13
    S, T, V = orc.one_electron(a, cn)
    assert np.abs(S - g[name + "_S"]).max() < 1e-13
    assert np.abs(T - g[name + "_T"]).max() < 1e-12
    assert np.abs(V - g[name + "_V"]).max() < 1e-12
    E = orc.eri(a, cn)
    assert E.shape == g[name + "_ERI"].shape
    assert np.abs(E - g[name + "_ERI"]).max() < 1e-13


@pytest.mark.parametrize("name", systems.FIXTURE_NAMES)
def test_integrals_vs_pyquante_fixtures(golden, name):
    """The reference's own fixtures (tests/test/*_pyquante.out; reference asserts places=5).
    Attainable: S 1e-13, T 1e-12, V 1e-7, ERI 2e-8 (the gap is the reference's Boys)."""
    g = golden["pyquante_fixtures"]
    a = arrays(name)
    cn = orc.normalization(a)
    S, T, V = orc.one_electron(a, cn)
    assert np.abs(S - g[name + "_S"]).max() < 1e-9   # fixture text carries ~10 digits
    assert np.abs(T - g[name + "_T"]).max() < 1e-9
    assert np.abs(V - g[name + "_V"]).max() < 2e-7
    if name + "_ERI" in g.files:
        assert np.abs(orc.eri(a, cn) - g[name + "_ERI"]).max() < 2e-8


def test_ch4_recorded_run(golden):
    """tests/output.molden: E_tot, all 12 per-iteration E_elec, MO energies, normalised coefficients."""
    g = golden["molden_ch4"]
    a = arrays("ch4")
    r = orc.rhf(a, max_scf=100)
    assert r["status"] == 0 and r["niter"] == 12
    assert abs(r["E"] - float(g["E_tot"])) < 2e-10
    assert np.abs(r["esteps"] - g["steps"]).max() < 2e-10
    assert np.abs(r["eps"] - g["moe"]).max() < 1e-9
    assert np.abs(orc.normalization(a) - g["gto"][:, 1]).max() < 1e-11
    D = r["C"][:, :a.ne] @ r["C"][:, :a.ne].T
    assert np.abs(D - g["D"]).max() < 1e-8


@pytest.mark.parametrize("name", list(systems.SYSTEMS))
def test_energy_matches_reference_source(golden, name):
    g = golden["reference_energy"]
    r = orc.rhf(arrays(name), max_scf=100)
    if float(g[name + "_E"]) == 99999.0:
        # plain Roothaan from the core guess does not converge for HCN within 100 iterations in the reference
        # either (Energy.py:317-322 returns the 99999 sentinel; tests/hcn.out is a truncated run)
        assert r["status"] == 1 and r["niter"] == 100
        assert abs(r["E"] - float(g[name + "_steps"][0])) < 1e-6 * abs(r["E"])
        return
    assert r["status"] == 0
    assert r["niter"] == len(g[name + "_steps"])
    assert np.abs(r["esteps"] - g[name + "_steps"]).max() < 1e-11 * max(1.0, abs(r["E"]))
    assert abs(r["E"] - float(g[name + "_E"])) < 1e-11 * max(1.0, abs(r["E"]))


def test_published_energies():
    """BASELINE.md section 2 (7 printed decimals)."""
    for name, e in (("h2", -1.0964861), ("h2o", -72.9771868), ("ch4", -38.6835840), ("h2o_sto3g", -74.9598578)):
        assert abs(orc.rhf(arrays(name))["E"] - e) < 6e-8


def test_scf_nonconvergence_is_reported():
    r = orc.rhf(arrays("h2o_sto3g"), max_scf=3)
    assert r["status"] == 1 and r["niter"] == 3


@pytest.mark.parametrize("name", ["h2_sto3g", "h2", "h2o_sto3g", "ch4_sto3g_p0", "hcn_sto3g", "ch2o", "h2o"])
def test_gradient_matches_reference_source(golden, name):
    """Forward-mode gradients vs the reference's Task.function_grads_algopy run here (1e-9)."""
    if not golden.has("reference_grad_" + name):
        pytest.skip("golden not generated")
    g = golden["reference_grad_" + name]
    a = arrays(name)
    for n in (0, 1, 2):
        if "g%d" % n not in g.files:
            continue
        r = orc.rhf_grad(a, n, max_scf=int(g["max_scf"]))
        assert r["status"] == 0
        assert np.abs(r["grad"] - g["g%d" % n]).max() < 1e-9, (name, n)
