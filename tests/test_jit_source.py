"""CPU: the run-time compiler (csrc/jit.cu).  The generated program of a sweep must be accepted by NVRTC for sm_100a together
with the device headers exactly as the library embeds them -- compiling needs no GPU, so the library's own compile path
(dlopen of NVRTC, worker threads, caches) runs here.  (What the compiled kernels compute is checked on the GPU:
tests/test_gpu_jit.py.)"""
import importlib
import os

import numpy as np
import pytest

from helpers import each_op_once, random_gates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    if not dm.query("jit_available"):
        pytest.skip("libnvrtc not loadable here")
    circuits = importlib.import_module("dm-sim_b200.circuits")
    rng = np.random.default_rng(3)
    allops = random_gates(6, 10, rng, names=["U3", "CX", "H", "T"], with_raw=False) + each_op_once(6, rng) + [("SRN", [2], 0, 0, 0)]
    cases = [(6, 1, allops, False), (8, 1, circuits.qft(8), False), (8, 2, random_gates(8, 40, rng), True)]
    n_compiled = 0
    for n, world, gates, peer in cases:
        progs = sweeps_of(dm, n, world, gates, peer)
        assert progs, "the generator must cover these sweeps"
        for i, (defines, program) in enumerate(progs[:4]):
            assert "#define DMB_JIT 1" in defines
            assert "__syncthreads();" in program or "#define DMB_J_N_GROUPS 0\n" in defines  # (a pure remap pack sweep has no ops)
            rc = dm.jit_compile(n, world, gates, i, peer=peer, wait=True)
            assert rc == 1, dm.lib().dmb_last_error().decode()[:3000]
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


def test_library_compiler_builds_caches_and_survives_exit(dm, tmp_path):
    """The library's own compile path (dlopen of NVRTC, worker threads, disk cache) without a GPU; and a process that EXITS
    while compilations are queued / running must not crash (the workers are drained at exit)."""
    import subprocess
    import sys
    if not dm.query("jit_available"):
        pytest.skip("libnvrtc not loadable here")
    code = (
        "import sys, os, importlib\n"
        "sys.path.insert(0, %r)\n"
        "dm = importlib.import_module('dm-sim_b200'); circuits = importlib.import_module('dm-sim_b200.circuits')\n"
        "g = circuits.qft(8)\n"
        "mode = sys.argv[1]\n"
        "if mode == 'wait':\n"
        "    assert dm.jit_compile(8, 1, g, 0) == 1\n"
        "    print('COUNTERS', int(dm.query('jit_compiled')), int(dm.query('jit_disk_hits')))\n"
        "else:\n"
        "    rs = [dm.jit_compile(8, 1, circuits.random_c1c2(8, 40, seed=s), 0, wait=False) for s in range(6)]\n"
        "    assert all(r in (0, 1) for r in rs), rs\n"
        "    print('QUEUED')\n" % ROOT)
    env = dict(os.environ, DMB_JIT_CACHE=str(tmp_path / "jitcache"))
    a = subprocess.run([sys.executable, "-c", code, "wait"], env=env, capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and "COUNTERS 1 0" in a.stdout, a.stdout + a.stderr
    b = subprocess.run([sys.executable, "-c", code, "wait"], env=env, capture_output=True, text=True, timeout=300)
    assert b.returncode == 0 and "COUNTERS 0 1" in b.stdout, b.stdout + b.stderr  # second process: disk cache hit, nothing compiled
    c = subprocess.run([sys.executable, "-c", code, "exit"], env=env, capture_output=True, text=True, timeout=300)
    assert c.returncode == 0 and "QUEUED" in c.stdout, (c.returncode, c.stdout + c.stderr)


def test_disk_cache_is_bounded(dm, tmp_path):
    """DMB_JIT_CACHE_MAX_MB: when the cache directory holds more, the oldest kernels go (checked when a process first uses it)."""
    import subprocess
    import sys
    import time
    if not dm.query("jit_available"):
        pytest.skip("libnvrtc not loadable here")
    cache = tmp_path / "jitcache"
    cache.mkdir()
    now = time.time()
    for i in range(8):  # 8 x 256 KiB of old "kernels", the first ones the oldest
        f = cache / ("%032x.cubin" % i)
        f.write_bytes(b"\0" * (256 << 10))
        os.utime(f, (now - 10000 + i, now - 10000 + i))
    code = ("import sys, importlib\nsys.path.insert(0, %r)\n"
            "dm = importlib.import_module('dm-sim_b200'); circuits = importlib.import_module('dm-sim_b200.circuits')\n"
            "assert dm.jit_compile(6, 1, circuits.qft(6), 0) == 1\n" % ROOT)
    env = dict(os.environ, DMB_JIT_CACHE=str(cache), DMB_JIT_CACHE_MAX_MB="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    left = sorted(p.name for p in cache.glob("*.cubin"))
    old = [n for n in left if n.startswith("0000000000000000000000000000000")]
    assert len(old) <= 2 and all(int(n[:32], 16) >= 6 for n in old), left   # 2 MiB > 1 MiB cap: pruned to <= 0.5 MiB, newest kept
    assert len(left) == len(old) + 1                                         # + the kernel this process compiled
