"""Golden trajectory of BASELINE config 3: `Tasks.runtask('Opt', argnum=[0,1])` on CH2O / STO-3G (geometry tests/Test.py:73-76),
first 3 BFGS iterations, computed with the reference's host loop over the CPU ORACLE (tests/oracle_opt.py).

    python tests/golden/make_opt_golden.py            (build container only; ~1 minute)

max_scf = 100 (the reference's default for 'Opt', Task.py:362): with BASELINE's max_scf=50 a line-search trial point of the
second iteration does not converge its SCF and the reference's gradient wrapper dies in UTPM.extract_jacobian(99999).

HCN (tests/Test.py:64-66), the other molecule BASELINE names, has NO oracle trajectory and no reference_grad_hcn_sto3g.npz:
  * the reference's plain Roothaan iteration needs 128 iterations for HCN / STO-3G (the oracle: E = -91.6751865757 after 128;
    with max_scf <= 100 the reference returns its 99999 sentinel and has no gradient at all), and
  * HCN is linear: its two pi orbitals are exactly degenerate, and forward mode through eigh (Eigen.py:69-74) divides by
    that zero gap inside the occupied space in EVERY iteration.  What algopy does there is not pinned by any reference
    fixture; oracle/fwdmode.py and oracle/rhf_oracle.cpp refuse (status 4) rather than guess.
  The GPU path only divides by occupied-virtual gaps, so it has a gradient for HCN; it is checked against central finite
  differences of the GPU energy instead (tests/test_gpu_workflows.py::test_config3_hcn_*).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import systems  # noqa: E402
import oracle_opt  # noqa: E402
from diffiqult_b200 import System_mol  # noqa: E402

if __name__ == "__main__":
    mol, basis, ne = systems.SYSTEMS["ch2o"]
    s = System_mol(mol, basis, ne, "ch2o")
    res = oracle_opt.run_opt(s, [0, 1], max_scf=100, maxiter=3)
    print(res.nit, res.fun, np.abs(res.jac).max(), res.message)
    np.savez(os.path.join(HERE, "oracle_opt_ch2o.npz"), x=res.x, fun=res.fun, jac=res.jac, nit=res.nit, allvecs=np.array(res.allvecs),
             nfev=res.nfev, njev=res.njev, max_scf=100, maxiter=3)
