"""Generate the (H2O)_n cluster goldens with the CPU oracle (run in the BUILD container only; oracle = test infrastructure).

    python tests/golden/make_cluster_golden.py w8 [w8far] [w8core] [w32]

Each golden holds the inputs' fingerprint, the `readguess` matrix C0 (monomer superposition, synthetic.monomer_guess with
the ORACLE solving the monomers; None for core-guess runs), the oracle's energy, iteration count and every per-iteration
E_elec (Energy.py:159-195), and - where the forward-mode direction count fits ORC_PMAX - the three gradient groups
(Task.py:25-51).  (H2O)_32 / 5.67 bohr (BASELINE config 4) takes ~20 CPU-minutes on 8 cores: energy and steps only.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from diffiqult_b200 import System_mol, synthetic  # noqa: E402
from diffiqult_b200.Basis import basis_set_3G_STO  # noqa: E402
from oracle import cpu_oracle as orc  # noqa: E402

# name -> (monomers, spacing, guess?, gradient?)
CASES = {"w8": (8, 5.67, True, True), "w8far": (8, 40.0, False, True), "w8core": (8, 24.0, False, True),
         "w16": (16, 5.67, True, False), "w32": (32, 5.67, True, False)}


def oracle_monomer(sm):
    o = orc.rhf(orc.SysArrays(sm), max_scf=100)
    assert o["status"] == 0
    return o["C"]


def make(name):
    n, spacing, guess, grad = CASES[name]
    mol = synthetic.water_cluster(n, spacing=spacing)
    s = System_mol(mol, basis_set_3G_STO, 10 * n, name)
    a = orc.SysArrays(s)
    out = dict(n_waters=np.int32(n), spacing=np.float64(spacing), xyz_sum=np.float64(np.sum(s.xyz)), max_scf=np.int32(100))
    if guess:
        C0 = synthetic.monomer_guess(mol, basis_set_3G_STO, solve=oracle_monomer)
        orc.set_guess(C0)
        out["C0"] = C0
    t0 = time.time()
    try:
        o = orc.rhf(a, max_scf=100)
        out.update(E=np.float64(o["E"]), status=np.int32(o["status"]), niter=np.int32(o["niter"]), steps=o["esteps"], eps=o["eps"])
        print(name, "E = %.10f" % o["E"], "status", o["status"], "niter", o["niter"], "%.1f s" % (time.time() - t0), flush=True)
        if grad:
            for g in (0, 1, 2):
                og = orc.rhf_grad(a, g, max_scf=100)
                assert og["status"] == 0 and og["niter"] == o["niter"], (og["status"], og["niter"])
                out["g%d" % g] = og["grad"]
                print(name, "grad", g, "|g|inf = %.6f" % np.abs(og["grad"]).max(), "%.1f s" % (time.time() - t0), flush=True)
    finally:
        orc.set_guess(None)
    np.savez_compressed(os.path.join(HERE, "oracle_cluster_%s.npz" % name), **out)


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["w8"]:
        make(nm)
