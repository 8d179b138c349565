import sys, os; sys.path.insert(0, ".")
os.environ["DQ_EIGH_DEBUG"] = "1"
import numpy as np, bench
from diffiqult_b200 import Energy
s = bench.workload(32)
r = Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
print(r["niter"], r["E"])
