"""Import-compatibility shim: the reference keeps its solver dispatcher in a module of this name
(diffiqult/Minimize.py); here it lives next to the solver in Optimize.py."""
from .Optimize import minimize  # noqa: F401
