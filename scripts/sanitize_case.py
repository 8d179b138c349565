"""Small mixed-contraction energy + gradient + integral-direct + screening run for compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import systems
from diffiqult_b200 import System_mol, Energy, Integrals
for name, table in (("h2o_mixed", systems.EXTRA_SYSTEMS), ("ch4_sto3g_p0", systems.SYSTEMS), ("h2", systems.SYSTEMS)):
    mol, basis, ne = table[name]
    s = System_mol(mol, basis, ne, name)
    args = (np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne)
    a = Energy.rhf_grad(*args, max_scf=60)
    b = Energy.rhf_grad(*args, max_scf=60, jk_mode="direct", direct_batch_bytes=1, screen_threshold=1e-30)
    cn = Integrals.normalization(s.alpha, s.coef, s.l, s.list_contr)
    e = Integrals.erivector(s.alpha, cn, s.xyz, s.l, s.nbasis, s.list_contr)
    print(name, a["E"], abs(a["E"] - b["E"]), a["niter"], float(np.abs(e).max()), flush=True)
print("done")
