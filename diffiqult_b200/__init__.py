"""diffiqult_b200: B200-native RHF hot path behind DiffiQult's System_mol / Tasks API.

    from diffiqult_b200 import System_mol, Tasks
    from diffiqult_b200.Basis import basis_set_3G_STO

mirrors `from diffiqult import ...` of the reference (diffiqult/__init__.py:69-70).  All numerics run in
libdiffiqult_b200.so (hand-written CUDA for sm_100a); there is no CPU fallback.
"""
from .Molecule import System_mol  # noqa: F401
from .Task import Tasks  # noqa: F401
