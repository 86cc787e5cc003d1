"""Import shim: the package directory is named ``image-caption-emotion-indonesia_b200`` (not a valid
Python identifier), so ``import icei_b200`` loads it under this importable alias."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_HERE, "image-caption-emotion-indonesia_b200")

if "icei_b200" not in sys.modules or getattr(sys.modules["icei_b200"], "__path__", None) is None:
    _spec = importlib.util.spec_from_file_location(
        "icei_b200", os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["icei_b200"] = _mod
    _spec.loader.exec_module(_mod)
