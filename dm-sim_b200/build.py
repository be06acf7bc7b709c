"""Builds the native libraries of the engine IN-TREE (so that they travel to the GPU box):

  dm-sim_b200/lib/libdmsim_b200.so              C-ABI (include/dmsim_b200.h): planner + sm_100a kernels
  dm-sim_b200/lib/libdmsim_py_nvgpu_omp*.so     pybind11 module with the reference's Python surface
                                                (reference src/py_nvgpu_omp_wrapper.cu:29-87)

nvcc cross-compiles for sm_100a without a GPU.  Nothing here touches oracle/ or /root/reference.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX, "--expt-relaxed-constexpr"]
# experiments: DMB_GEOM=<thread bits>,<register bits> builds another CTA geometry (csrc/devop.hpp); default 7,4
GEOM = [f"-DDMB_THREAD_BITS={int(a)}" for a in os.environ.get("DMB_GEOM", "").split(",")[:1] if a] + \
       [f"-DDMB_REG_BITS={int(a)}" for a in os.environ.get("DMB_GEOM", "").split(",")[1:2] if a]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def core_sources():
    srcs = [os.path.join(CSRC, f) for f in ("plan.cpp", "encode.cpp", "kernels.cu", "sweep_kernel.cu", "capi.cu")]
    hdrs = [os.path.join(CSRC, f) for f in ("plan.hpp", "kernels.cuh", "devop.hpp", "encode.hpp")] + [os.path.join(ROOT, "include", "dmsim_b200.h")]
    return srcs, hdrs


def build_core(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libdmsim_b200.so")
    srcs, hdrs = core_sources()
    if not force and not _newer(out, srcs + hdrs):
        return out
    objs = []
    for s in srcs:
        o = os.path.join(LIB, os.path.basename(s) + ".o")
        if force or _newer(o, [s] + hdrs):
            if s.endswith(".cpp"):  # pure host code: the system compiler, no CUDA front-end
                _run([HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-Wall"] + GEOM + ["-c", s, "-o", o], verbose)
            else:
                _run([NVCC] + ARCH + NVCC_FLAGS + GEOM + ["-Xptxas", "-v" if verbose else "-O3", "-c", s, "-o", o], verbose)
        objs.append(o)
    _run([NVCC] + ARCH + ["-shared", "-ccbin", HOST_CXX, "-o", out] + objs + ["-ldl"], verbose)
    return out


def build_pybind(force=False, verbose=False):
    import pybind11

    os.makedirs(LIB, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    out = os.path.join(LIB, "libdmsim_py_nvgpu_omp" + ext)
    src = os.path.join(CSRC, "pybind_module.cpp")
    hdrs = [os.path.join(ROOT, "include", "dmsim_b200.h"), os.path.join(ROOT, "include", "dmsim_b200.hpp")]
    if not os.path.exists(src):
        return None
    if not force and not _newer(out, [src] + hdrs + [os.path.join(LIB, "libdmsim_b200.so")]):
        return out
    cmd = [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", src, "-o", out,
           "-I" + os.path.join(ROOT, "include"), "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           "-L" + LIB, "-ldmsim_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd, verbose)
    return out


def build_all(force=False, verbose=False):
    core = build_core(force, verbose)
    py = build_pybind(force, verbose)
    return core, py


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
