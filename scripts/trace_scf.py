import sys, os
import numpy as np
sys.path.insert(0, ".")
os.environ["DQ_TRACE"] = "1"
import bench
from diffiqult_b200 import Energy
s = bench.workload(32)
for rep in range(2):
    r = Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=100)
    print(r["stats"], flush=True)
