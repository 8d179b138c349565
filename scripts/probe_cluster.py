"""Scratch probe: (H2O)_n STO-3G energy (+gradient) timings and SCF behaviour on the GPU."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from diffiqult_b200 import System_mol, synthetic, Energy, _capi
from diffiqult_b200.Basis import basis_set_3G_STO
import ctypes as C

tf = C.c_double(0)
_capi.check(_capi.lib().dq_fp64_peak(C.c_int32(0), C.byref(tf)))
print("fp64 FMA peak TFLOP/s", tf.value, flush=True)
for n in [int(x) for x in sys.argv[1].split(",")]:
    mol = synthetic.water_cluster(n, spacing=24.0)
    s = System_mol(mol, basis_set_3G_STO, 10 * n, "w%d" % n)
    for grad in (False, True):
        t0 = time.time()
        fn = Energy.rhf_grad if grad else Energy.rhf
        r = fn(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=int(sys.argv[2]))
        dt = time.time() - t0
        st = r["stats"]
        print(json.dumps(dict(n=n, grad=grad, wall=dt, E=r["E"], status=r["status"], niter=r["niter"],
                              gmax=(max(float(np.abs(g).max()) for g in r["grad"]) if grad else None), **st)), flush=True)
        print("  esteps tail", r.get("esteps", [])[-4:] if not grad else "", flush=True)
