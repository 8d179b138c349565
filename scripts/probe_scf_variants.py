"""Scratch probe: SCF of the systems that disagree with the oracle, under every A/B switch (one process per variant)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, systems
    from diffiqult_b200 import System_mol, Energy
    from oracle import cpu_oracle as orc
    for name in ("hcn", "h2o_mixed", "h2o", "ch2o"):
        table = systems.SYSTEMS if name in systems.SYSTEMS else systems.EXTRA_SYSTEMS
        mol, basis, ne = table[name]
        s = System_mol(mol, basis, ne, name)
        r = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=30)
        o = orc.rhf(orc.SysArrays(s), max_scf=30)
        k = min(len(r["esteps"]), len(o["esteps"]))
        d = np.abs(r["esteps"][:k] - o["esteps"][:k])
        first = int(np.argmax(d > 1e-8)) if (d > 1e-8).any() else -1
        print("   %-10s niter %3d/%3d  first step differing > 1e-8: %d  (max diff %.3e)" % (name, r["niter"], o["niter"], first, d.max()), flush=True)
else:
    for env in ({}, {"DQ_SORT": "0"}, {"DQ_ACCUM": "float"}, {"DQ_JK_LDG": "1"}, {"DQ_SORT": "0", "DQ_ACCUM": "float", "DQ_JK_LDG": "1"},
                {"DQ_COLD_EIGH": "1"}, {"DQ_EIGH_PACKED": "1"}):
        print(env, flush=True)
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e)
