"""Element look-ups used by the text writers (reference: diffiqult/Param.py:1-16)."""

select_atom = {1: "H", 3: "Li", 6: "C", 7: "N", 8: "O", 9: "F"}

select_weight = {1: 1.00794, 6: 12, 7: 15, 8: 16, 9: 18.998403}
