import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
src=open('tests/test_gpu_parity.py').read()
ns={}
exec(src[src.index('def _random_system'):src.index('@pytest.mark.parametrize("seed"')], {'np':np}, ns)
from diffiqult_b200 import System_mol, Energy
from oracle import cpu_oracle as orc
seed=int(sys.argv[1])
rng=np.random.default_rng(1000+seed)
mol,basis,nel=ns['_random_system'](rng)
s=System_mol(mol,basis,nel,'r')
print(mol); print(basis); print('l',s.l,'contr',s.list_contr,'ne',s.ne)
a=orc.SysArrays(s)
r=Energy.rhf_grad(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=40)
o=orc.rhf(a,max_scf=40)
print('E',r['E'],o['E'],r['niter'],o['niter'], 'eps', r.get('eps'))
print('esteps diff', np.abs(np.array(o['esteps'])[:r['niter']]-Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=40)['esteps']).max())
print('orbital energies', Energy.rhf(np.log(s.alpha), s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, max_scf=40)['eps'])
for n in (0,1,2):
    og=orc.rhf_grad(a,n,max_scf=40)
    d=np.abs(r['grad'][n]-og['grad'])
    print(n,'maxdiff',d.max(),'at',d.argmax(),'scale',np.abs(og['grad']).max()); print(' gpu',r['grad'][n]); print(' orc',og['grad'])
