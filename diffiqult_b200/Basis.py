"""Basis-set tables of the reference (data, not code).

The numbers are the reference's `diffiqult/Basis.py` tables (basis_set_2G_STO_uncontracted
Basis.py:4-69, basis_set_3G_STO_uncontracted :71-169, basis_set_6G_STO_uncontracted
:171-364, basis_set_3G_STO :366-470), stored as data/basis_sets.json and rebuilt here into
the reference's dict layout:

    {Z: [('S'|'P', [(exponent, contraction_coefficient), ...]), ...]}
"""
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    with open(os.path.join(_HERE, "data", "basis_sets.json")) as fh:
        raw = json.load(fh)
    out = {}
    for name, table in raw.items():
        out[name] = {
            int(z): [(kind, [(float(e), float(c)) for e, c in prims]) for kind, prims in shells]
            for z, shells in table.items()
        }
    return out


_tables = _load()
basis_set_2G_STO_uncontracted = _tables["basis_set_2G_STO_uncontracted"]
basis_set_3G_STO_uncontracted = _tables["basis_set_3G_STO_uncontracted"]
basis_set_6G_STO_uncontracted = _tables["basis_set_6G_STO_uncontracted"]
basis_set_3G_STO = _tables["basis_set_3G_STO"]
