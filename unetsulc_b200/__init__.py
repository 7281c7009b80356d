"""Importable alias of the package directory ``2022_pauriau_unetsulc_b200`` (a Python identifier cannot start
with a digit).  ``import unetsulc_b200`` and ``import unetsulc_b200.models`` resolve to the same modules."""
import importlib
import sys

_real = importlib.import_module("2022_pauriau_unetsulc_b200")
sys.modules[__name__] = _real
for _k, _v in list(sys.modules.items()):
    if _k.startswith("2022_pauriau_unetsulc_b200."):
        sys.modules["unetsulc_b200." + _k.split(".", 1)[1]] = _v
