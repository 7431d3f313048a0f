"""Importable alias of the package directory `scalable-collaborative-zksnark_b200/`
(its name is not a valid Python identifier): `import scz_b200 as scz`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("scalable-collaborative-zksnark_b200")
sys.modules[__name__] = _pkg
