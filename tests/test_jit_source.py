"""CPU: the run-time compiler's input (csrc/jit.cu).  The generated program of a sweep must be accepted by NVRTC for sm_100a
together with the device headers exactly as the library embeds them -- no GPU is needed to compile.  (What the compiled
kernels compute is checked on the GPU: tests/test_gpu_jit.py.)"""
import importlib
import os

import numpy as np
import pytest

from helpers import each_op_once, random_gates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dm-sim_b200", "csrc")
ENTRY = '''#include "sweep_device.cuh"
extern "C" __global__ void __launch_bounds__(dmb::kTileThreads, 3) dmb_jit_sweep(const __grid_constant__ dmb::SweepArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    dmb::sweep_body<DMB_J_MASK, 1>(a, smem_raw);
}
'''


def nvrtc_compile(defines, program):
    nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
    hdrs = {"devop.hpp": open(os.path.join(CSRC, "devop.hpp")).read(),
            "sweep_device.cuh": open(os.path.join(CSRC, "sweep_device.cuh")).read(), "dmb_jit_program.inc": program}
    names = [n.encode() for n in hdrs]
    err, prog = nvrtc.nvrtcCreateProgram((defines + ENTRY).encode(), b"dmb_jit_sweep.cu", len(names), [hdrs[n].encode() for n in hdrs], names)
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-default-device", b"-lineinfo"]
    (res,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    assert res == nvrtc.nvrtcResult.NVRTC_SUCCESS, log.decode()[:3000]
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    assert n > 0
    return n


def sweeps_of(dm, n, world, gates, peer=False):
    out, i = [], 0
    while True:
        r = dm.jit_source(n, world, gates, i, peer=peer)
        if r is None:
            return out
        out.append(r)
        i += 1


def test_generated_programs_compile_for_sm_100a(dm):
    """Every op body through the generator (all 38 ops + C1 / C2 + SRN on small tiles), the TMA / direct-store skeleton
    (full-size tiles of a QFT) and the peer-store variant of a sharded plan."""
    circuits = importlib.import_module("dm-sim_b200.circuits")
    rng = np.random.default_rng(3)
    allops = random_gates(6, 10, rng, names=["U3", "CX", "H", "T"], with_raw=False) + each_op_once(6, rng) + [("SRN", [2], 0, 0, 0)]
    cases = [(6, 1, allops, False), (8, 1, circuits.qft(8), False), (8, 2, random_gates(8, 40, rng), True)]
    n_compiled = 0
    for n, world, gates, peer in cases:
        progs = sweeps_of(dm, n, world, gates, peer)
        assert progs, "the generator must cover these sweeps"
        for defines, program in progs[:4]:
            assert "#define DMB_JIT 1" in defines
            assert "__syncthreads();" in program or "#define DMB_J_N_GROUPS 0\n" in defines  # (a pure remap pack sweep has no ops)
            nvrtc_compile(defines, program)
            n_compiled += 1
    assert n_compiled >= 4


def test_program_text_is_independent_of_the_payload(dm):
    """Same structure, other angles -> the same text (one cached kernel for a whole VQE loop); another structure -> another text."""
    def circuit(scale, extra=False):
        rng = np.random.default_rng(9)
        gs = []
        for _ in range(25):
            gs.append(("RY", [int(rng.integers(7))], scale * float(rng.uniform(0.2, 1.2)), 0, 0))
            a, b = (int(x) for x in rng.choice(7, 2, replace=False))
            gs.append(("CX", [a, b], 0, 0, 0))
            gs.append(("RZ", [int(rng.integers(7))], scale * float(rng.uniform(0.2, 1.2)), 0, 0))
        return gs + ([("H", [3], 0, 0, 0), ("CX", [3, 5], 0, 0, 0)] if extra else [])
    a, b, c = (sweeps_of(dm, 7, 1, g) for g in (circuit(1.0), circuit(0.6), circuit(1.0, extra=True)))
    assert a and a == b
    assert a != c
