"""Alias used by scripts generated for the reference (tool/dmsim_qasm.py:345, tool/randomtest_n14.py:18):
``import dmsim_py_omp_wrapper as dmsim``.  Re-exports the pybind11 module built from dm-sim_b200/csrc/pybind_module.cpp."""
import importlib.util as _u
import glob as _g
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_cands = _g.glob(_os.path.join(_here, "dm-sim_b200", "lib", "libdmsim_py_nvgpu_omp*.so"))
if not _cands:
    raise ImportError("libdmsim_py_nvgpu_omp is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
_spec = _u.spec_from_file_location("libdmsim_py_nvgpu_omp", _cands[0])
_mod = _u.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
_sys.modules.setdefault("libdmsim_py_nvgpu_omp", _mod)
globals().update({k: getattr(_mod, k) for k in dir(_mod) if not k.startswith("__")})
