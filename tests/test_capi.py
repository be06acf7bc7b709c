"""CPU: the C-ABI library loads, exports every symbol include/dmsim_b200.h declares, keeps its POD layouts, and
fails LOUDLY (no fallback) when no GPU is usable.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "dmsim_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dmb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(dm):
    L = dm.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dmsim_b200.h but not exported"


def test_pod_layouts(dm):
    assert dm.GATE_DTYPE.itemsize == 56
    assert ctypes.sizeof(dm.dmb_stats) == 3 * 8 + 10 * 8
    assert dm.lib().dmb_version().decode().startswith("dmsim-b200")


def test_header_compiles_as_c_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "dmsim_b200.h"\nint main(void){ dmb_gate g; g.op = DMB_OP_RYY; return sizeof(g) == 56 ? 0 : 1; }\n')
    for cc, flags in (("/usr/bin/gcc", ["-std=c11"]), ("/usr/bin/g++", ["-std=c++17", "-x", "c++"])):
        exe = tmp_path / "t.out"
        subprocess.run([cc, *flags, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
        assert subprocess.run([str(exe)]).returncode == 0


def test_cxx_dropin_header_builds_against_the_library(tmp_path):
    """The reference's example/adder_n10 driver, written against include/dmsim_b200.hpp, compiles and links."""
    src = os.path.join(ROOT, "examples", "adder_n10.cpp")
    exe = tmp_path / "adder"
    lib = os.path.join(ROOT, "dm-sim_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", str(exe),
                    "-L", lib, "-ldmsim_b200", "-Wl,-rpath," + lib], check=True)
    assert os.path.exists(exe)


def test_xacc_plugin_abi_builds_against_the_library(tmp_path):
    """include/DmSimApi.hpp (the reference's xacc/DmSimApi.hpp plugin ABI: DmSimBackend + getGpuDmSim) compiles and
    links; the enum mirrors DMSim::OP value for value (xacc/DmSimApi.hpp:6-45)."""
    src = tmp_path / "abi.cpp"
    src.write_text('#include "DmSimApi.hpp"\n'
                   'static_assert((int)DmSim::OP::RYY == 37 && (int)DmSim::OP::CH == (int)DMSim::OP::CH && '
                   '(int)DmSim::OP::C4X == (int)DMSim::OP::C4X, "enum order");\n'
                   'int main(int argc, char**){ std::shared_ptr<DmSim::DmSimBackend> b = DmSim::getGpuDmSim(); '
                   'if (argc > 99) { b->init(2); b->addGate(DmSim::OP::H, {0}); b->measure(1); b->finalize(); } return b ? 0 : 1; }\n')
    lib = os.path.join(ROOT, "dm-sim_b200", "lib")
    exe = tmp_path / "abi"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", lib, "-ldmsim_b200", "-Wl,-rpath," + lib], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "xacc_backend.cpp"), "-o", str(tmp_path / "x"), "-L", lib, "-ldmsim_b200",
                    "-Wl,-rpath," + lib], check=True)


def test_no_silent_fallback_without_gpu(dm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dm.DMSimError, match="no usable CUDA device"):
        dm.Simulation(3, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dm-sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), fn
                assert "oracle/" not in txt.replace("Nothing here imports ``oracle/``", "").replace(
                    "Nothing here touches oracle/", ""), fn


def test_pybind_module_surface(dm):
    """Same names as the reference's module (src/py_nvgpu_omp_wrapper.cu:29-87)."""
    import dmsim_py_omp_wrapper as m
    for name in ("append", "upload", "clear_circuit", "run", "reset", "measure"):
        assert hasattr(m.Simulation, name)
    for op in dm.OP_NAMES:
        assert hasattr(m.Simulation, op), op
    assert m.Simulation.CRZ(0.5, 1, 2).dump() == "CRZ(1,2,0,0,0,0,0,0.5);\n"
    assert m.Simulation.RZ(0.5, 1).dump() == "RZ(1,0,0,0,0,0,0.5,0);\n"
