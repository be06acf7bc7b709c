"""-m gpu: the run-time specialised sweep kernels (csrc/jit.cu) against the oracle and against the interpreter kernels.

The default engine only specialises sweeps of large shards; these tests force it on (jit_min_bits = 0) at sizes the
oracle finishes in seconds, in the waiting mode (jit = 2) and in the tiered mode (jit = 1)."""
import time

import numpy as np
import pytest

from helpers import each_op_once, random_gates
from test_gpu_parity import TOL, run_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture
def jit(dm):
    if not dm.query("jit_available"):
        pytest.fail("libnvrtc is not loadable on this box: the specialised kernels cannot be built")

    def set_(mode, min_bits=0, **kw):
        dm.set_option("jit", mode)
        dm.set_option("jit_min_bits", min_bits)
        for k, v in kw.items():
            dm.set_option(k, v)
    yield set_
    dm.set_option("jit", 1); dm.set_option("jit_min_bits", 16); dm.set_option("jit_hot_small", 8); dm.set_option("tma", 1)
    dm.set_option("persistent", 0); dm.set_option("graph", 1)


@pytest.mark.parametrize("n,small", [(6, 1), (8, 1), (9, 0), (10, 1)])
def test_specialised_kernels_match_oracle_and_interpreter(dm, oracle_mod, jit, n, small):
    """Random circuits over every op: specialised == oracle (1e-12) and == the interpreter kernels (same op bodies, same
    order: 1e-14).  small = 1: full-size tiles go through TMA (+ direct store of the last round); 0: plain tile I/O."""
    rng = np.random.default_rng(500 + n)
    gates = random_gates(n, 70, rng)
    ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
    jit(0, tma=small, persistent=0)
    ref = run_gpu(dm, n, gates)
    assert dm.query("jit_sweeps", ref._h) == 0
    rre, rim = ref.get_dm()
    jit(2, tma=small, persistent=0)
    before = dm.query("jit_failed")
    sim = run_gpu(dm, n, gates)
    assert dm.query("jit_failed") == before
    assert dm.query("jit_sweeps", sim._h) == sim.last_stats["n_sweeps"] and dm.query("jit_pending", sim._h) == 0
    re, im = sim.get_dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
    assert max(np.abs(re - rre).max(), np.abs(im - rim).max()) < 1e-14
    assert abs(sim.trace() - 1.0) < TOL


def test_specialised_each_op_alone(dm, oracle_mod, jit):
    """One op per run after a scrambling prefix, every op body through the generator."""
    n = 6
    jit(2, persistent=0)
    rng = np.random.default_rng(11)
    prefix = random_gates(n, 10, rng, names=["U3", "CX", "H", "T"], with_raw=False)
    for g in each_op_once(n, rng) + [("SRN", [2], 0, 0, 0)]:
        gates = prefix + [g]
        sim = run_gpu(dm, n, gates)
        assert dm.query("jit_sweeps", sim._h) == sim.last_stats["n_sweeps"], g[0]
        re, im = sim.get_dm()
        ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
        err = max(np.abs(re - ore).max(), np.abs(im - oim).max())
        assert err < TOL, f"op {g[0]} on {g[1]}: {err}"


def test_tiered_execution_switches_to_specialised_kernels(dm, oracle_mod, jit):
    """jit = 1: the first run is interpreted (the compiler starts when the plan runs a second time); once the compiler is done the same circuit runs specialised (the
    captured graph is refreshed), with the same result, and a second circuit of the same STRUCTURE (other angles) reuses
    the kernels without compiling."""
    n = 9
    jit(1, persistent=0, jit_hot_small=2)  # (small shards are only compiled for from their 8th run on by default)

    def circuit(scale):
        rng = np.random.default_rng(77)
        gs = []
        for _ in range(40):
            q = int(rng.integers(n))
            gs.append(("RY", [q], scale * float(rng.uniform(0.2, 1.2)), 0, 0))
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            gs.append(("CX", [a, b], 0, 0, 0))
            gs.append(("RZ", [int(rng.integers(n))], scale * float(rng.uniform(0.2, 1.2)), 0, 0))
        return gs

    gates = circuit(1.0)
    ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
    sim = dm.Simulation(n, 1)
    rec, mats = dm.pack_gates(gates)
    results = []
    deadline = time.time() + 120
    while True:
        sim.reset_dm()
        dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data, mats.size // 32))
        sim._uploaded = True
        sim.run()
        re, im = sim.get_dm()
        assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
        results.append(dm.query("jit_sweeps", sim._h))
        if dm.query("jit_pending", sim._h) == 0 and results[-1] == sim.last_stats["n_sweeps"]:
            break
        assert time.time() < deadline, f"specialised kernels never became ready: {results}"
        time.sleep(0.2)
    compiled = dm.query("jit_compiled")
    gates2 = circuit(0.7)
    sim2 = run_gpu(dm, n, gates2)
    assert dm.query("jit_sweeps", sim2._h) == 0, "a plan's first run is interpreted (tiered by hotness)"
    sim2.reset_dm()
    sim2.run()  # second run of the plan: it is hot now, and its kernels exist already
    re, im = sim2.get_dm()
    ore, oim = oracle_mod.Oracle(n).sim(gates2).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
    assert dm.query("jit_compiled") == compiled, "same structure, other angles: no new kernel"
    assert dm.query("jit_sweeps", sim2._h) == sim2.last_stats["n_sweeps"]


def test_specialised_n13_fullsize_tiles(dm, oracle_mod, jit):
    """n = 13 (1 GiB): TMA tiles, direct store, stars; diagonal and trace against the oracle's state."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    n = 13
    jit(2)
    gates = circuits.qft(n) + circuits.random_c1c2(n, 30, seed=3)
    sim = run_gpu(dm, n, gates)
    assert dm.query("jit_sweeps", sim._h) == sim.last_stats["n_sweeps"]
    jit(0)
    ref = run_gpu(dm, n, gates)
    idx = np.random.default_rng(1).integers(0, 4 ** n, size=1 << 16, dtype=np.uint64)
    a = sim.elements(idx)
    b = ref.elements(idx)
    assert max(np.abs(a[0] - b[0]).max(), np.abs(a[1] - b[1]).max()) < 1e-14
    assert np.abs(sim.diag() - ref.diag()).max() < 1e-14
    assert abs(sim.trace() - 1.0) < TOL
