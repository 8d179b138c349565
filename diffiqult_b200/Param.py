"""Element look-ups used by the text writers (same names as the reference's diffiqult/Param.py)."""

# Z, symbol, mass used by the reference (None: the reference has no entry)
_ELEMENTS = ((1, "H", 1.00794), (3, "Li", None), (6, "C", 12), (7, "N", 15), (8, "O", 16), (9, "F", 18.998403))

select_atom = {z: sym for z, sym, _ in _ELEMENTS}
select_weight = {z: w for z, _, w in _ELEMENTS if w is not None}
