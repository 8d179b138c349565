"""diffiqult_b200: B200-native RHF hot path behind DiffiQult's System_mol / Tasks API."""
from .Molecule import System_mol  # noqa: F401
