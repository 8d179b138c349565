import sys
import numpy as np
sys.path.insert(0, ".")
from diffiqult_b200 import System_mol, Tasks, synthetic
from diffiqult_b200.Basis import basis_set_3G_STO, basis_set_2G_STO_uncontracted
for ref in (True, False):
    s = System_mol(synthetic.H2O, basis_set_3G_STO, 10, mol_name='agua')
    m = Tasks(s, name='/tmp/h2_sto_3g', verbose=False)
    m.runtask('Energy', max_scf=50, printcoef=True, name='Output.molden', output=False)
    m.runtask('Opt', max_scf=50, printcoef=False, maxiter=3, argnum=[0], output=False, reference_linesearch=ref)
    r = m.opt_result
    print("H2O STO-3G [0] maxiter=3 ref_ls=%s: E %.6f |g| %.6f nit %d nfev %d njev %d msg %s" % (ref, r.fun, np.abs(r.jac).max(), r.nit, r.nfev, r.njev, r.message))
    print("  alpha[:7]", np.round(s.alpha[:7], 6), " expected 135.263104 23.755473 6.558207 5.162051 1.199706 0.378480 5.320693 ; E -74.976643 |g| 0.256754")
    s = System_mol(synthetic.H2_EXAMPLE, basis_set_3G_STO, 2, mol_name='agua')
    m = Tasks(s, name='/tmp/h2_sto_3g', verbose=False)
    m.runtask('Opt', max_scf=50, printcoef=False, argnum=[0, 1], output=True, reference_linesearch=ref)
    r = m.opt_result
    print("H2 STO-3G [0,1] ref_ls=%s: E %.6f |g| %.6f nit %d msg %s" % (ref, r.fun, np.abs(r.jac).max(), r.nit, r.message))
    print("  alpha[:2]", np.round(s.alpha[:2], 6), " expected 4.803590 0.694786 ; E -1.095971 |g| 0.000008 nit 26")
    s = System_mol(synthetic.H2O, basis_set_2G_STO_uncontracted, 10, mol_name='agua')
    m = Tasks(s, name='/tmp/h2o_sto_2g', verbose=False)
    m.runtask('Opt', max_scf=100, argnum=[0], maxiter=5, reference_linesearch=ref)
    r = m.opt_result
    print("H2O 2G-unc [0] maxiter=5 ref_ls=%s: E %.6f |g| %.6f nit %d" % (ref, r.fun, np.abs(r.jac).max(), r.nit))
    print("  alpha", np.round(s.alpha, 6), " expected 135.545392 19.574369 4.944644 0.689005 2.851426 2.798600 2.883584 0.543853 0.489780 0.453495 1.371735 0.230254 ; E -74.667865 |g| 1.589181")
