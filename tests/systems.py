"""Test systems: the reference's own fixtures (tests/Test.py:29-165) + BASELINE.json configs."""
import numpy as np

from diffiqult_b200 import synthetic as syn
from diffiqult_b200.Basis import basis_set_3G_STO

# test-local uncontracted basis of the reference's unit tests (tests/Test.py:79-140)
basis_set_test = {
    1: [('S', [(1.309756377, 0.430128498)]), ('S', [(0.233135974, 0.678913531)])],
    6: [('S', [(27.38503303, 0.430128498)]), ('S', [(4.874522052, 0.678913531)]),
        ('S', [(1.136748189, 0.51154070)]), ('S', [(0.288309360, 0.612819896)]),
        ('P', [(1.136748189, 0.51154070)]), ('P', [(0.288309360, 0.612819896)])],
    7: [('S', [(99.106168999999994, 0.15432897000000001)]), ('S', [(18.052312000000001, 0.53532813999999995)]),
        ('S', [(4.8856602000000002, 0.44463454000000002)]), ('S', [(3.7804559000000002, -0.099967230000000004)]),
        ('S', [(0.87849659999999996, 0.39951282999999999)]), ('S', [(0.28571439999999998, 0.70011546999999996)]),
        ('P', [(3.7804559000000002, 0.15591627)]), ('P', [(0.87849659999999996, 0.60768372000000004)]),
        ('P', [(0.28571439999999998, 0.39195739000000002)])],
    8: [('S', [(49.9809710, 0.4301280)]), ('S', [(8.8965880, 0.6789140)]),
        ('S', [(1.9452370, 0.0494720)]), ('S', [(0.4933630, 0.9637820)]),
        ('P', [(1.9452370, 0.0494720)]), ('P', [(0.4933630, 0.9637820)])],
    9: [('S', [(63.7352020, 0.4301280)]), ('S', [(11.3448340, 0.6789140)]), ('S', [(11.3448340, 0.6789140)]),
        ('S', [(2.4985480, 0.0494720)]), ('S', [(0.6336980, 0.9637820)]),
        ('P', [(2.4985480, 0.5115410)]), ('P', [(0.6336980, 0.6128200)])],
}

# name -> (mol, basis, n_electrons)
# the five fixture systems of tests/Test.py:141-165 ...
SYSTEMS = {
    "h2": (syn.H2, basis_set_test, 2),
    "h2o": (syn.H2O, basis_set_test, 10),
    "ch4": (syn.CH4, basis_set_test, 10),
    "hcn": (syn.HCN, basis_set_test, 14),
    "ch2o": (syn.CH2O, basis_set_3G_STO, 16),
    # ... and the BASELINE.json configurations in STO-3G
    "h2_sto3g": (syn.H2_EXAMPLE, basis_set_3G_STO, 2),
    "h2o_sto3g": (syn.H2O, basis_set_3G_STO, 10),
    "hcn_sto3g": (syn.HCN, basis_set_3G_STO, 14),
    "ch4_sto3g": (syn.CH4, basis_set_3G_STO, 10),
    "ch4_sto3g_p0": (syn.perturbed_methanes(2)[0], basis_set_3G_STO, 10),
    "ch4_sto3g_p1": (syn.perturbed_methanes(2)[1], basis_set_3G_STO, 10),
}
FIXTURE_NAMES = ["h2", "h2o", "ch4", "hcn", "ch2o"]


def ltype(l):
    """(lx,ly,lz) tuple -> 0 (s), 1 (px), 2 (py), 3 (pz)."""
    return 0 if max(l) == 0 else 1 + list(l).index(max(l))

# systems without reference goldens (checked against the pinned CPU oracle only): long and mixed contractions
from diffiqult_b200.Basis import basis_set_6G_STO_uncontracted as _u6  # noqa: E402

_h6 = [p for _, prims in _u6[1] for p in prims]
basis_h_6g = {1: [('S', [(e, c) for e, c in _h6])]}          # one s function contracted from 6 primitives (KK = 36)
basis_mixed = {
    1: [('S', basis_set_3G_STO[1][0][1][:2]), ('S', basis_set_3G_STO[1][0][1][2:])],        # 2 + 1 primitives
    8: basis_set_3G_STO[8] + [('P', [(0.9, 1.0)]), ('S', [(0.35, 0.6), (1.7, 0.5), (9.0, 0.2), (40.0, 0.05)])],
}
# four hydrogens, one 6-primitive s function each: a full block of 4 functions with 36 primitive pairs per function pair
# (576 entries per side of a tile: the list-building and ket-chunk loops of the compacting ERI kernel run several rounds)
H4 = [(1, (0.0, 0.0, 0.0)), (1, (0.0, 0.0, 1.5)), (1, (1.4, 0.3, 0.2)), (1, (-0.4, 1.3, 1.1))]
EXTRA_SYSTEMS = {
    "h4_6g": (H4, basis_h_6g, 4),
    "h2_6g": (syn.H2_EXAMPLE, basis_h_6g, 2),
    "h2o_mixed": (syn.H2O, basis_mixed, 10),
}
