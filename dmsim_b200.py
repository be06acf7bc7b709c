"""Alias: ``import dmsim_b200`` == the package in the hyphenated directory ``dm-sim_b200/``."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _il.import_module("dm-sim_b200")
_sys.modules[__name__] = _pkg
