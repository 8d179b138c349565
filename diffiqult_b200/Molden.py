"""Writer of the reference's molden-like result file (text I/O only).

Block order, labels and number formatting follow the nested `write_molden` of the reference (`Energy.py:208-308`) and
`Tools.printmatrix` (`Tools.py:176-187`), so files diff against `tests/output.molden`.  Quirks kept: the `[INPUT]` block
prints `coef[function index]` instead of the primitive's coefficient (`Energy.py:261`), and `[MO]` marks ne+1 orbitals as
occupied (`j > ne`, `Energy.py:303`).  All numbers come from the device results of `Energy.rhf`; the natural-orbital
occupations (eigenvalues of D) are computed with the library's own eigensolver.
"""
import ctypes as C

import numpy as np

from . import _capi
from .Param import select_atom


def _printmatrix(matrix, tape):
    tape.write('  ' + ''.join(str(i) + '     ' for i in range(len(matrix))) + '\n')
    for i, row in enumerate(matrix):
        tape.write(str(i) + ' ' + ''.join(str(float(x)) + ' ' for x in row) + '\n')


def _eigenvalues(D, device):
    n = D.shape[0]
    A = np.ascontiguousarray(D, dtype=np.float64)
    w, V = np.zeros(n), np.zeros((n, n))
    _capi.check(_capi.lib().dq_eigh(C.c_int32(n), C.c_int32(1), A.ctypes, w.ctypes, V.ctypes, C.c_int32(device)))
    return w


def write_molden(path, result, alpha, coefn, xyz, l, charges, atoms, contr_list, ne, device=0):
    """result: dict from Energy.rhf; alpha: exponents; coefn: NORMALISED contraction coefficients."""
    Cm, eps, steps = result["C"], result["eps"], result["esteps"]
    n, natoms = Cm.shape[0], len(charges)
    D = Cm[:, :ne] @ Cm[:, :ne].T
    e_elec = float(steps[-1])
    e_tot = float(result["E"])
    with open(path, "w") as tape:
        tape.write('[Energy] \n')
        tape.write('E_elec: ' + str(e_elec) + '\n')
        tape.write('E_nuc: ' + str(e_tot - e_elec) + '\n')
        tape.write('E_tot: ' + str(e_tot) + '\n')
        tape.write('SCF Details\nEigen: True\n')
        for i, step in enumerate(steps):
            tape.write('Step: ' + str(i) + ' ' + str(float(step)) + '\n')
        tape.write('[CM] \nC AO times MO\n')
        _printmatrix(Cm, tape)
        tape.write('[DM] \nD \n')
        _printmatrix(D, tape)
        tape.write('[NMO] \nNMO \n')
        _printmatrix([_eigenvalues(D, device)], tape)
        tape.write('[MOE] \nMOE \n')
        for i, e in enumerate(eps):
            tape.write(str(i) + ' ' + str(float(e)) + '\n')
        tape.write('[INPUT] \n')
        line = 'mol = ['
        for i, c in enumerate(atoms):
            line += '(' + str(charges[i]) + ',(' + str(c[0]) + ',' + str(c[1]) + ',' + str(c[2]) + ')),\n'
        tape.write(line)
        cont, line = 0, 'basis = ['
        for i, ci in enumerate(contr_list):
            line += '[(' + str(l[i][0]) + ',' + str(l[i][1]) + ',' + str(l[i][2]) + '),'
            for _ in range(ci):
                line += str(alpha[cont]) + ',' + str(coefn[i]) + ',(' + str(xyz[i, 0]) + ',' + str(xyz[i, 1]) + ',' + str(xyz[i, 2]) + ')],\n'
                cont += 1
            line += ']\n'
        tape.write(line)
        tape.write('[Atoms]\n')
        for i, c in enumerate(atoms):
            tape.write(str(select_atom.get(int(charges[i]))) + ' ' + str(i + 1) + ' ' + str(charges[i]) + ' ' + str(c[0]) + ' ' +
                       str(c[1]) + ' ' + str(c[2]) + '\n')
        for i, c in enumerate(xyz):
            tape.write('XX ' + str(i + natoms + 1) + ' 0 ' + str(c[0]) + ' ' + str(c[1]) + ' ' + str(c[2]) + '\n')
        cont = 0
        tape.write('[GTO]\n')
        for i, ci in enumerate(contr_list):
            tape.write('  ' + str(i + 1 + natoms) + '  0\n')
            kind = ' s   ' if np.sum(l[i]) == 0 else ' p   '
            tape.write(kind + str(ci) + ' 1.0 ' + str(l[i][0]) + ' ' + str(l[i][1]) + ' ' + str(l[i][2]) + '\n')
            for _ in range(ci):
                tape.write('  ' + str(alpha[cont]) + ' ' + str(coefn[cont]) + '\n')
                cont += 1
            tape.write('  \n')
        tape.write('[MO]\n')
        for j in range(n):
            tape.write('  Sym=  None\n  Ene=      ' + str(float(eps[j])) + '\n  Spin= Alpha\n')
            tape.write('  Occup=    0.0\n' if j > ne else '  Occup=    2.0\n')
            for i in range(n):
                tape.write(str(i + 1 + natoms) + ' ' + str(float(Cm[i, j])) + '\n')
