"""Print the achieved GPU-vs-oracle / GPU-vs-reference-run deviations (numbers quoted in DESIGN.md)."""
import sys, os
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import systems
from diffiqult_b200 import System_mol, Energy, Integrals
from oracle import cpu_oracle as orc
G = lambda n: np.load(os.path.join("tests", "golden", n + ".npz"))
gi, ge = G("reference_integrals"), G("reference_energy")
worst = dict(S=0, T=0, V=0, ERI=0, E=0, steps=0)
for name in systems.SYSTEMS:
    mol, basis, ne = systems.SYSTEMS[name]
    s = System_mol(mol, basis, ne)
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    S, T, V = Integrals.one_electron_matrices(s.alpha, cn, s.xyz, s.l, s.nbasis, s.charges, s.atom, s.list_contr)
    E = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    d = dict(S=np.abs(S - gi[name + "_S"]).max(), T=np.abs(T - gi[name + "_T"]).max(), V=np.abs(V - gi[name + "_V"]).max(),
             ERI=np.abs(E - gi[name + "_ERI"]).max())
    r = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    if float(ge[name + "_E"]) != 99999.0:
        d["E"] = abs(r["E"] - float(ge[name + "_E"])); d["steps"] = np.abs(r["esteps"] - ge[name + "_steps"]).max()
    print(name, {k: "%.1e" % v for k, v in d.items()}, "niter", r["niter"])
    for k, v in d.items(): worst[k] = max(worst[k], v)
print("WORST vs reference run:", {k: "%.1e" % v for k, v in worst.items()})
wg = 0
for name in ["h2_sto3g", "h2", "h2o_sto3g", "ch4_sto3g_p0", "ch2o", "h2o"]:
    g = G("reference_grad_" + name)
    mol, basis, ne = systems.SYSTEMS[name]
    s = System_mol(mol, basis, ne)
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=int(g["max_scf"]))
    for n in (0, 1, 2):
        if "g%d" % n in g.files:
            dv = np.abs(r["grad"][n] - g["g%d" % n]).max(); wg = max(wg, dv)
            print(name, "argnum", n, "max|dgrad| %.1e" % dv)
print("WORST gradient deviation vs reference run: %.1e" % wg)
