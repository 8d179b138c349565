"""Scratch probe: does plain Roothaan converge for (H2O)_n at a given grid spacing?"""
import sys
import numpy as np
sys.path.insert(0, ".")
from diffiqult_b200 import System_mol, synthetic, Energy
from diffiqult_b200.Basis import basis_set_3G_STO
n = int(sys.argv[1])
for sp in [float(x) for x in sys.argv[2].split(",")]:
    mol = synthetic.water_cluster(n, spacing=sp)
    s = System_mol(mol, basis_set_3G_STO, 10 * n, "w")
    r = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=60)
    d = np.abs(np.diff(r["esteps"]))
    print(n, sp, "status", r["status"], "niter", r["niter"], "E %.8f" % r["E"], "last |dE|", d[-3:], flush=True)
