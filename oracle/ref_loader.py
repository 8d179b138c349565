"""TEST INFRASTRUCTURE ONLY (oracle/): run the UNMODIFIED reference sources under Python 3.

The reference (/root/reference/diffiqult) is Python-2-only and needs `algopy`, which is
absent here, so it cannot be imported as a package.  This loader reads the reference's
own source text where it lies (nothing is copied into the repo), applies the minimal
syntactic patches listed below IN MEMORY, and executes the modules with
`oracle/fwdmode.py` standing in for `algopy`.  The numerics that run are therefore the
reference's own lines: this is what pins the restated oracle and generates the golden
vectors under tests/golden/ (tests/golden/make_golden.py).

/root/reference does not exist on the GPU box; only make_golden.py (run in the build
container) uses this module.

In-memory patches (all syntactic, none numerical):
  P1  tabs expanded to 8-column stops (Python 2 semantics; Energy.py:123,149,159,
      Task.py:118-122 mix tabs and spaces)
  P2  `print x, y` statements -> print(x, y)           (Tools.py:90,124,150,196-198,210;
                                                        Task.py:316-317)
  P3  np.float(1.0) -> float(1.0)                       (Energy.py:130; Task.py:143)
  P4  Integrals.py:195,197: `/8`, `/2` used as sizes -> `//` (Python 2 int division)
  P5  Task.py imports of modules that no longer exist / are broken in the reference
      itself are dropped: scipy.optimize.optimize (Task.py:2), Dipole (Task.py:3,
      whose own `from Data import ...` fails even on py2.7), Minimize (Task.py:4,
      needs scipy._lib.six and minpack2); time.clock -> time.process_time.
  P6  implicit relative imports (`import Tools`, `from Integrals import ...`) are
      satisfied by registering the loaded modules under those bare names in
      sys.modules while the reference code is being executed.
"""
import os
import re
import sys
import types

from . import fwdmode

REF_ROOT = os.environ.get("DIFFIQULT_REFERENCE", "/root/reference")
_PKG = os.path.join(REF_ROOT, "diffiqult")

_PRINT_STMT = re.compile(r"^(\s*)print[ \t]+(?!\()(.*?)\s*$")


def available():
    return os.path.isfile(os.path.join(_PKG, "Integrals.py"))


def _patch(name, text):
    lines = []
    for ln in text.split("\n"):
        ln = ln.expandtabs(8)  # P1
        m = _PRINT_STMT.match(ln)
        if m and not ln.lstrip().startswith("#"):  # P2
            ln = "%sprint(%s)" % (m.group(1), m.group(2))
        ln = ln.replace("np.float(1.0)", "float(1.0)")  # P3
        lines.append(ln)
    text = "\n".join(lines)
    if name == "Integrals":  # P4
        text = text.replace("+ 3*nbasis + 2)/8", "+ 3*nbasis + 2)//8")
        text = text.replace("nbasis*(nbasis + 1)/2 ", "nbasis*(nbasis + 1)//2 ")
    if name == "Task":  # P5
        text = text.replace("from scipy.optimize import optimize as opt", "opt = None")
        text = text.replace("from Dipole import dipolemoment", "dipolemoment = None")
        text = text.replace("from Minimize import minimize", "minimize = None")
        text = text.replace("time.clock()", "time.process_time()")
    return text


_ORDER = ["Param", "Basis", "Tools", "Integrals", "Eigen", "Energy", "Molecule", "Task"]
_cache = None


def load():
    """Return a dict name -> executed reference module (Tools, Integrals, Energy, ...)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference sources not found under %s" % _PKG)
    algopy = types.ModuleType("algopy")
    for k in ("UTPM", "zeros", "exp", "log", "sqrt", "dot", "transpose", "diag", "sum", "eigh"):
        setattr(algopy, k, getattr(fwdmode, k))
    saved = {k: sys.modules.get(k) for k in _ORDER + ["algopy"]}
    mods = {}
    try:
        sys.modules["algopy"] = algopy  # P6
        for name in _ORDER:
            path = os.path.join(_PKG, name + ".py")
            with open(path) as fh:
                src = _patch(name, fh.read())
            mod = types.ModuleType(name)
            mod.__file__ = path
            sys.modules[name] = mod
            exec(compile(src, path, "exec"), mod.__dict__)
            mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mods["algopy"] = algopy
    _cache = mods
    return mods


class bare_names(object):
    """Context manager: keep the reference modules importable under their bare Python-2 names
    (P6) while reference code that imports lazily runs (Energy.py:209 `import Param`)."""

    def __enter__(self):
        mods = load()
        self._saved = {k: sys.modules.get(k) for k in mods}
        sys.modules.update(mods)
        return mods

    def __exit__(self, *exc):
        for k, v in self._saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        return False
