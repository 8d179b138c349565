"""Generate the golden vectors under tests/golden/ (run in the BUILD container only).

Sources of truth, in order of strength:
  1. the reference's own fixtures: PyQuante dumps tests/test/*_pyquante.out and the CH4
     run recorded in tests/output.molden (packed here into small .npz files);
  2. outputs of the reference's OWN source executed here under Python 3 through
     oracle/ref_loader.py (+ oracle/fwdmode.py standing in for algopy): normalised
     coefficients, S, T, V, packed ERI vectors, SCF energies with per-iteration E_elec,
     Boys-function tables and parameter gradients.

Usage:  python tests/golden/make_golden.py [fixtures] [integrals] [boys] [energy] [grad:<system>:<argnums>]
With no arguments everything except the slow gradients is regenerated.
"""
import io
import os
import re
import sys
import tempfile
import time
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader, fwdmode  # noqa: E402
import systems  # noqa: E402

REF_TESTS = os.path.join(ref_loader.REF_ROOT, "tests")


def ref():
    return ref_loader.load()


def ref_system(name):
    mol, basis, ne = systems.SYSTEMS[name]
    return ref()["Molecule"].System_mol(mol, basis, ne, name)


def system_inputs(s):
    return dict(alpha=np.array(s.alpha, float), coef=np.array(s.coef, float), xyz=np.array(s.xyz, float),
                ltype=np.array([systems.ltype(l) for l in s.l], np.int32),
                contr=np.array(s.list_contr, np.int32), charges=np.array(s.charges, np.int32),
                atoms=np.array(s.atom, float), ne=np.int32(s.ne))


# ---------------------------------------------------------------- 1. reference fixtures
def pack_fixtures():
    T = ref()["Tools"]
    out = {}
    for name in systems.FIXTURE_NAMES:
        n = ref_system(name).nbasis
        for kind, tag in (("overlap", "S"), ("kinetics", "T"), ("nuclear", "V")):
            rows = np.loadtxt(os.path.join(REF_TESTS, "test", "%s_%s_pyquante.out" % (name, kind)))
            m = np.zeros((n, n))
            m[rows[:, 0].astype(int), rows[:, 1].astype(int)] = rows[:, 2]
            assert rows.shape[0] == n * n
            out["%s_%s" % (name, tag)] = m
        path = os.path.join(REF_TESTS, "test", "%s_eri_pyquante.out" % name)
        if os.path.isfile(path):
            rows = np.loadtxt(path)
            assert rows.shape[0] == n ** 4
            mm = n * (n + 1) // 2
            vec = np.full(mm * (mm + 1) // 2, np.nan)
            spread = 0.0
            for i, j, k, l, v in rows:
                idx = T.eri_index(int(i), int(j), int(k), int(l), n)
                if np.isnan(vec[idx]):
                    vec[idx] = v
                else:
                    spread = max(spread, abs(vec[idx] - v))
            assert not np.isnan(vec).any()
            out["%s_ERI" % name] = vec
            out["%s_ERI_spread" % name] = np.float64(spread)  # asymmetry inside the dump itself
            print(name, "eri packed", vec.size, "dump-internal spread", spread)
    np.savez_compressed(os.path.join(HERE, "pyquante_fixtures.npz"), **out)

    # tests/output.molden: the recorded CH4 run
    lines = open(os.path.join(REF_TESTS, "output.molden")).read().split("\n")
    g = dict(E_elec=float(lines[1].split()[1]), E_nuc=float(lines[2].split()[1]), E_tot=float(lines[3].split()[1]))
    steps = [float(l.split()[2]) for l in lines if l.startswith("Step:")]
    def block(tag):
        i = next(k for k, l in enumerate(lines) if l.startswith(tag))
        rows = []
        for l in lines[i + 3:]:
            p = l.split()
            if len(p) != 19:
                break
            rows.append([float(x) for x in p[1:]])
        return np.array(rows)
    i = next(k for k, l in enumerate(lines) if l.startswith("[MOE]"))
    moe = np.array([float(l.split()[1]) for l in lines[i + 2:i + 20]])
    # normalised contraction coefficients from the [GTO] block: "  alpha coef"
    i = next(k for k, l in enumerate(lines) if l.startswith("[GTO]"))
    gto = []
    for l in lines[i + 1:]:
        if l.startswith("[MO]"):
            break
        p = l.split()
        if len(p) == 2 and "." in p[0] and "." in p[1]:
            gto.append((float(p[0]), float(p[1])))
    np.savez_compressed(os.path.join(HERE, "molden_ch4.npz"), steps=np.array(steps), C=block("[CM]"),
                        D=block("[DM]"), moe=moe, gto=np.array(gto), **g)
    print("molden ch4: steps", len(steps), "gto", len(gto), g)


# ---------------------------------------------------------------- 2. reference run here
def run_integrals():
    I = ref()["Integrals"]
    out = {}
    for name in systems.SYSTEMS:
        s = ref_system(name)
        t0 = time.time()
        inp = system_inputs(s)
        one = np.float64(1.0)
        cn = I.normalization(np.array(s.alpha), s.coef, s.l, s.list_contr)
        S = I.overlapmatrix(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr, one)
        T = I.kineticmatrix(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr, one)
        V = I.nuclearmatrix(s.alpha, cn, s.xyz, s.l, s.nbasis, s.charges, s.atom, s.natoms, s.list_contr, one)
        E = I.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr, float(1.0))
        for k, v in inp.items():
            out["%s_in_%s" % (name, k)] = v
        out["%s_coefn" % name] = np.array(cn, float)
        out["%s_S" % name], out["%s_T" % name], out["%s_V" % name] = S, T, V
        out["%s_ERI" % name] = np.array(E, float)
        print(name, "N", s.nbasis, "eri", len(E), "%.1fs" % (time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "reference_integrals.npz"), **out)


def run_boys():
    T = ref()["Tools"]
    ts = np.concatenate([[0.0, 1e-13, 1e-12, 1e-9, 1e-6], np.linspace(1e-3, 8.0, 161),
                         np.array([0.5, 1.5, 2.5, 3.5, 4.5, 5.5])[:, None].repeat(5, 1).ravel()
                         + np.tile([-1e-9, -1e-13, 0.0, 1e-13, 1e-9], 6),
                         np.geomspace(8.0, 400.0, 60)])
    val = np.zeros((5, ts.size))
    der = np.zeros((5, ts.size))
    for m in range(5):
        for k, t in enumerate(ts):
            val[m, k] = T.incompletegammaf(float(t), m)
            d = T.incompletegammaf(fwdmode.Dual(float(t), np.ones(1)), m)
            der[m, k] = d.tan[0] if isinstance(d, fwdmode.Dual) else 0.0
            assert abs(float(d) - val[m, k]) == 0.0
    np.savez_compressed(os.path.join(HERE, "reference_boys.npz"), t=ts, F=val, dF=der)
    print("boys table", val.shape)


def _run_energy(name, max_scf=100):
    s = ref_system(name)
    tk = ref()["Task"].Tasks(s, os.path.join(tempfile.gettempdir(), "golden_" + name), verbose=False)
    args = tk._energy_args(max_scf=max_scf, max_d=300, log=True, name=os.path.join(tempfile.gettempdir(), "golden_" + name),
                           write=True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), ref_loader.bare_names():
        e = ref()["Energy"].rhfenergy(*args)
    n = s.nbasis
    if e == 99999:  # Energy.py:317-322: SCF did not converge within max_scf; E_tot is only printed
        etot = [float(l.split()[1]) for l in buf.getvalue().split("\n") if l.startswith("E_tot:")][0]
        return 99999.0, np.array([etot]), np.zeros((n, n)), np.zeros(n)
    lines = open(args[16] + ".molden").read().split("\n")
    steps = np.array([float(l.split()[2]) for l in lines if l.startswith("Step:")])
    i = next(k for k, l in enumerate(lines) if l.startswith("[CM]"))
    C = np.array([[float(x) for x in l.split()[1:]] for l in lines[i + 3:i + 3 + n]])
    i = next(k for k, l in enumerate(lines) if l.startswith("[MOE]"))
    moe = np.array([float(l.split()[1]) for l in lines[i + 2:i + 2 + n]])
    return float(e), steps, C, moe


def run_energy():
    out = {}
    for name in systems.SYSTEMS:
        t0 = time.time()
        e, steps, C, moe = _run_energy(name)
        out["%s_E" % name], out["%s_steps" % name], out["%s_C" % name], out["%s_moe" % name] = e, steps, C, moe
        print(name, "E = %.12f" % e, "iters", len(steps), "%.1fs" % (time.time() - t0))
    np.savez_compressed(os.path.join(HERE, "reference_energy.npz"), **out)


def run_grad(name, argnum, max_scf=100):
    s = ref_system(name)
    tk = ref()["Task"].Tasks(s, os.path.join(tempfile.gettempdir(), "goldeng_" + name), verbose=False)
    path = os.path.join(HERE, "reference_grad_%s.npz" % name)
    out = dict(np.load(path)) if os.path.isfile(path) else {}
    for n in argnum:
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            g = tk._energy_gradss([n], max_scf=max_scf)  # Task.py:223-240 -> grad_evaluator_algopy
        out["g%d" % n] = np.array(g[0], float)
        out["max_scf"] = np.int32(max_scf)
        print(name, "argnum", n, "len", len(g[0]), "|g|inf %.6f" % np.abs(g[0]).max(), "%.1fs" % (time.time() - t0))
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    todo = sys.argv[1:] or ["fixtures", "integrals", "boys", "energy"]
    for item in todo:
        if item == "fixtures":
            pack_fixtures()
        elif item == "integrals":
            run_integrals()
        elif item == "boys":
            run_boys()
        elif item == "energy":
            run_energy()
        elif item.startswith("grad:"):
            _, name, nums = item.split(":")
            run_grad(name, [int(c) for c in nums])
        else:
            raise SystemExit("unknown item " + item)
