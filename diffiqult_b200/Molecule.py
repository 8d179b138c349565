"""System containers: the user-visible half of the drop-in boundary.

Mirrors the reference's `diffiqult/Molecule.py` (Getbasis :5-84, Getgeom :87-107,
System_mol :110-218): same constructor arguments, same attribute names and layouts, so a
script written against the reference runs unchanged.

    mol       = [(Z, (x, y, z)), ...]                       bohr unless angs=True
    basis_set = {Z: [('S'|'P', [(exp, coef), ...]), ...]}

Every Cartesian component of a 'P' entry becomes an independent basis function with its
own copy of the primitives (Molecule.py:60-71,81-84); exponents of px/py/pz may therefore
diverge during an optimisation and the kernels never assume shell structure.
"""
import numpy as np

from .Param import select_atom

_CARTESIAN = {"S": [(0, 0, 0)], "P": [(1, 0, 0), (0, 1, 0), (0, 0, 1)]}


class Getbasis(object):
    """Flatten (molecule, basis dict) into primitive/function arrays (Molecule.py:23-84)."""

    def __init__(self, molecule, basis, shifted=False):
        self.alpha, self.coef, self.xyz, self.l, self.list_contr = [], [], [], [], []
        self.tot_prim = 0
        if shifted:
            # reference quirk kept (Molecule.py:73-79): list_contr is never filled, so a
            # "shifted" system reports nbasis == 0 in the reference as well.
            for gauss in basis:
                self.l.append(gauss[0])
                self.alpha.append(gauss[1])
                self.coef.append(gauss[2])
                self.xyz.append(gauss[3])
            return
        for z, pos in molecule:
            for kind, prims in basis[z]:
                for lxyz in _CARTESIAN[kind]:
                    self.list_contr.append(len(prims))
                    self.xyz.append([pos[0], pos[1], pos[2]])
                    self.l.append(lxyz)
                    for e, c in prims:
                        self.alpha.append(e)
                        self.coef.append(c)
                        self.tot_prim += len(lxyz)


class Getgeom(object):
    """Charges and nuclear coordinates (Molecule.py:92-107)."""

    def __init__(self, molecule):
        self.charge = [int(z) for z, _ in molecule]
        self.xyz = [[p[0], p[1], p[2]] for _, p in molecule]


class System_mol(object):
    """All a task needs to know about a molecule + floating-Gaussian basis (Molecule.py:130-189).

    Attributes (names and layouts of the reference): alpha[nprim], coef[nprim] (raw,
    un-normalised), xyz[nbasis,3], l (list of nbasis 3-tuples), list_contr[nbasis],
    nbasis, charges[natoms] (int), atom[natoms,3], natoms, ne (= electrons // 2).
    """

    def __init__(self, mol, basis_set, ne, mol_name="molecule", angs=False, shifted=False):
        self.mol_name = mol_name
        b = Getbasis(mol, basis_set, shifted=shifted)
        self.list_contr = b.list_contr
        self.nbasis = len(b.list_contr)
        self.alpha = np.array(b.alpha)
        self.xyz = np.reshape(np.array(b.xyz, dtype="float64"), (self.nbasis, 3))
        self.l = b.l
        g = Getgeom(mol)
        self.charges = np.array(g.charge)
        self.atom = np.array(g.xyz)
        self.natoms = len(self.charges)
        self.coef = np.array(b.coef)
        self.ne = int(ne / 2)
        if angs:
            factor = 0.529177249  # sic: the reference multiplies (Molecule.py:185-188)
            self.xyz = factor * self.xyz
            self.atom = factor * self.atom

    def printcurrentgeombasis(self, tape):
        """Same text block as Molecule.py:191-218 (alpha is indexed by function, sic)."""
        tape.write(" Atomic coordinates\n")
        for i, c in enumerate(self.atom):
            tape.write("  %s %d %s %s %s %s\n" % (select_atom.get(int(self.charges[i])), i + 1,
                                                 self.charges[i], c[0], c[1], c[2]))
        tape.write(" Basis functions\n")
        for i, c in enumerate(self.xyz):
            kind = "s" if np.sum(self.l[i]) == 0 else "p"
            tape.write("  %d %2.6f %2.6f %2.6f %s %5.6f %d %d %d\n" % (
                i + 1, c[0], c[1], c[2], kind, self.alpha[i],
                self.l[i][0], self.l[i][1], self.l[i][2]))
        tape.write("\n")
