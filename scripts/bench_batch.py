"""Config 5 probe: N perturbed CH4 / STO-3G energy + full gradient through the threaded batch driver."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from diffiqult_b200 import System_mol, synthetic, batch
from diffiqult_b200.Basis import basis_set_3G_STO
n = int(sys.argv[1])
syss = [System_mol(m, basis_set_3G_STO, 10, "ch4") for m in synthetic.perturbed_methanes(n)]
batch.rhf_grad_batch(syss[:16], threads=4)
for th in [int(x) for x in sys.argv[2].split(",")]:
    t0 = time.perf_counter()
    E, st, g = batch.rhf_grad_batch(syss, max_scf=50, threads=th)
    dt = time.perf_counter() - t0
    print("molecules %d threads %d: %.3f s  (%.2f ms/molecule)  converged %d  E[0] %.10f" % (n, th, dt, dt / n * 1e3, int((st == 0).sum()), E[0]), flush=True)
