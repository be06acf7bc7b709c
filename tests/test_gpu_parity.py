"""-m gpu: the CUDA path (through the C-ABI) against the oracle, 1e-12 absolute on every element."""
import os

import numpy as np
import pytest

from helpers import each_op_once, random_gates, to_complex

pytestmark = pytest.mark.gpu
TOL = 1e-12  # BASELINE.json north_star: elements and diagonal within 1e-12 absolute, trace within 1e-12 of 1


def run_gpu(dm, n, gates, n_gpus=1):
    sim = dm.Simulation(n, n_gpus)
    rec, mats = dm.pack_gates(gates)
    dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                       mats.size // 32))
    sim._uploaded = True
    sim.run()
    return sim


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_random_circuits_all_ops(dm, oracle_mod, n):
    rng = np.random.default_rng(100 + n)
    gates = random_gates(n, 60, rng)
    sim = run_gpu(dm, n, gates)
    re, im = sim.get_dm()
    ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
    assert np.abs(re - ore).max() < TOL and np.abs(im - oim).max() < TOL
    assert abs(sim.trace() - 1.0) < TOL
    assert np.abs(sim.diag() - np.diagonal(ore)).max() < TOL


@pytest.mark.parametrize("n", [5, 8])
def test_each_op_alone(dm, oracle_mod, n):
    """One op per run, after a scrambling prefix, so a wrong op cannot hide behind the others."""
    rng = np.random.default_rng(7)
    prefix = random_gates(n, 12, rng, names=["U3", "CX", "H", "T"], with_raw=False)
    for g in each_op_once(n, rng):
        gates = prefix + [g]
        sim = run_gpu(dm, n, gates)
        re, im = sim.get_dm()
        ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
        err = max(np.abs(re - ore).max(), np.abs(im - oim).max())
        assert err < TOL, f"op {g[0]} on {g[1]}: {err}"


def test_srn_single_run(dm, oracle_mod):
    """SRN is real-linear, not a matrix (reference :1253-1266)."""
    n = 4
    gates = [("H", [0], 0, 0, 0), ("U3", [1], 0.3, 0.2, 0.1), ("SRN", [1], 0, 0, 0), ("CX", [1, 2], 0, 0, 0),
             ("SRN", [0], 0, 0, 0)]
    sim = run_gpu(dm, n, gates)
    re, im = sim.get_dm()
    ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL


def test_adder_n10(dm, oracle_mod):
    """example/adder_n10: deterministic outcome 0b1000000010 (README.md:235-244) and the full 16 MiB matrix."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    gates = circuits.adder_n10()
    sim = run_gpu(dm, 10, gates)
    assert sim.last_stats["n_gates"] == 30 and sim.last_stats["n_primitives"] == 142
    res = sim.measure(5, seed=1)
    assert res == [0b1000000010] * 5
    re, im = sim.get_dm()
    ore, oim = oracle_mod.Oracle(10).sim(gates).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL


def test_continuation_across_runs(dm, oracle_mod):
    """State persists across run(); clear_circuit keeps it (reference README.md:206-210)."""
    n = 6
    rng = np.random.default_rng(3)
    a, b = random_gates(n, 20, rng), random_gates(n, 20, rng)
    sim = dm.Simulation(n, 1)
    for part in (a, b):
        for g in part:
            sim.append(dm.Gate(g[0], *(list(g[1]) + [0] * (5 - len(g[1]))), theta=g[2], phi=g[3], lam=g[4],
                               matrix=g[5] if len(g) > 5 else None))
        sim.upload()
        sim.run()
        sim.clear_circuit()
    re, im = sim.get_dm()
    ore, oim = oracle_mod.Oracle(n).sim(a + b).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
    sim.reset()
    d = sim.diag()
    assert d[0] == 1.0 and np.abs(d[1:]).max() == 0.0


def test_sampling_matches_reference_rule(dm, oracle_mod):
    n = 8
    rng = np.random.default_rng(11)
    gates = random_gates(n, 40, rng)
    sim = run_gpu(dm, n, gates)
    o = oracle_mod.Oracle(n).sim(gates)
    r = rng.uniform(0, 1, size=500)
    r[:3] = [0.0, 1.0, 0.999999999999]
    got, total = sim.sample(r)
    want = o.sample_with_r(r)
    # a draw within 1e-12 of a bin edge may legitimately land in the neighbouring bin
    scan = np.concatenate([[0.0], np.cumsum(np.abs(o.diag()))])
    near_edge = np.min(np.abs(scan[None, :] - r[:, None]), axis=1) < 1e-12
    assert np.all((got == want) | near_edge)
    assert abs(total - scan[-1]) < TOL
    # measure() with the reference's generator
    m_gpu = sim.measure(20, seed=1234)
    m_ref, _ = o.measure(20, seed=1234)
    assert m_gpu == [int(x) for x in m_ref]


def test_purity_and_set_dm(dm):
    n = 5
    rng = np.random.default_rng(5)
    dim = 1 << n
    a = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    rho = a @ a.conj().T
    rho /= np.trace(rho).real
    sim = dm.Simulation(n, 1)
    stored = rho.T  # the engine stores rho^T, [col][row]
    sim.set_dm(np.ascontiguousarray(stored.real), np.ascontiguousarray(stored.imag))
    assert abs(sim.trace() - 1.0) < TOL
    assert abs(sim.purity() - np.sum(np.abs(rho) ** 2)) < 1e-12
    re, im = sim.get_dm()
    assert np.array_equal(re, stored.real) and np.array_equal(im, stored.imag)


def test_mixed_state_gate_application(dm, oracle_mod):
    """Gates on a non-pure, non-trivial state (set_dm / set_state on both sides)."""
    n = 5
    rng = np.random.default_rng(6)
    dim = 1 << n
    a = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    rho = a @ a.conj().T
    rho /= np.trace(rho).real
    stored = np.ascontiguousarray(rho.T)
    gates = random_gates(n, 30, rng)
    sim = dm.Simulation(n, 1)
    sim.set_dm(np.ascontiguousarray(stored.real), np.ascontiguousarray(stored.imag))
    rec, mats = dm.pack_gates(gates)
    dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                       mats.size // 32))
    sim._uploaded = True
    sim.run()
    re, im = sim.get_dm()
    o = oracle_mod.Oracle(n)
    sre, sim_ = np.ascontiguousarray(stored.real), np.ascontiguousarray(stored.imag)  # keep alive across the call
    o.lib.orc_set_state(o.h, sre.ctypes.data, sim_.ctypes.data)
    ore, oim = o.sim(gates).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL


@pytest.mark.parametrize("opts", [dict(tile_bits=12, low_bits=3), dict(tile_bits=8, low_bits=2),
                                  dict(tile_bits=10, low_bits=5), dict(tile_bits=6, low_bits=0)])
def test_tile_geometries(dm, oracle_mod, opts):
    """Same circuit under different tile sizes / contiguous-run lengths (many sweeps, arbitrary tile bit sets)."""
    n = 9
    rng = np.random.default_rng(21)
    gates = random_gates(n, 80, rng)
    try:
        for k, v in opts.items():
            dm.set_option(k, v)
        dm.set_option("min_tiles_log2", 4)
        sim = run_gpu(dm, n, gates)
        re, im = sim.get_dm()
    finally:
        dm.set_option("tile_bits", 12); dm.set_option("low_bits", 3); dm.set_option("min_tiles_log2", 10)
    ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL


def test_graph_and_stream_paths_agree(dm):
    """The three executors of a single-GPU run -- one cooperative launch for all sweeps (small states), CUDA graph, plain
    stream launches -- give bit-identical results."""
    n = 8
    rng = np.random.default_rng(31)
    gates = random_gates(n, 50, rng)
    outs, launches = [], []
    for persistent, graph in ((1, 1), (0, 1), (0, 0)):
        dm.set_option("persistent", persistent)
        dm.set_option("graph", graph)
        dm.set_option("tile_bits", 8)
        try:
            sim = run_gpu(dm, n, gates)
            outs.append(sim.get_dm())
            launches.append((sim.last_stats["n_launches"], sim.last_stats["n_sweeps"]))
        finally:
            dm.set_option("graph", 1); dm.set_option("tile_bits", 12); dm.set_option("persistent", 0)
    for o in outs[1:]:
        assert np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1])
    assert launches[0][1] > 1 and launches[0][0] == 1, "the small-state run should be ONE launch"
    assert launches[2][0] == launches[2][1]


def test_one_launch_executor_vqe_golden_and_continuation(dm, oracle_mod):
    """benchmark/vqe_uccsd_n8.qasm (10808 gates, ~100 sweeps) as one cooperative launch against the reference's golden
    diagonal, then a second circuit continuing from the evolved state."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "vqe_uccsd_n8.npz"))
    gates = circuits.vqe_uccsd_n8()
    dm.set_option("persistent", 1)  # (off by default: measured slower than the captured graph, profiles/README.md)
    try:
        sim = run_gpu(dm, 8, gates)
        assert sim.last_stats["n_launches"] == 1 and sim.last_stats["n_sweeps"] > 50
        assert np.abs(sim.diag() - z["diag"]).max() < TOL
        rng = np.random.default_rng(5)
        more = random_gates(8, 30, rng)
        sim.clear_circuit()
        for g in more:
            sim.append(dm.Gate(g[0], *(list(g[1]) + [0] * (5 - len(g[1]))), theta=g[2], phi=g[3], lam=g[4], matrix=g[5] if len(g) > 5 else None))
        sim.upload()
        sim.run()
        o = oracle_mod.Oracle(8).sim(gates + more)
        assert np.abs(sim.diag() - o.diag()).max() < 1e-11  # (10838 gates deep: the reference itself drifts by 1e-12)
    finally:
        dm.set_option("persistent", 0)


def test_plan_cache_reuses_and_alternates(dm, oracle_mod):
    """A circuit that is set again reuses its plan and the device tables (h2d_bytes == 0); alternating circuits, resets and
    the cache switched off give the same state as the oracle."""
    n = 9
    rng = np.random.default_rng(77)
    a, b = random_gates(n, 40, rng), random_gates(n, 40, rng)
    L = dm.lib()

    def set_(sim, gates):
        rec, mats = dm.pack_gates(gates)
        dm._check(L.dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32))
        sim._uploaded = True

    try:
        for cache in (1, 0):
            dm.set_option("plan_cache", cache)
            sim = dm.Simulation(n, 1)
            o = oracle_mod.Oracle(n)
            h2d = []
            for gates in (a, a, b, a, b, b):
                sim.reset_dm()
                set_(sim, gates)
                sim.run()
                h2d.append(sim.last_stats["h2d_bytes"])
                re, im = sim.get_dm()
                ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
                assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
            if cache:
                assert h2d[0] > 0 and h2d[1] == 0 and h2d[2] > 0 and h2d[5] == 0
            else:
                assert all(x > 0 for x in h2d)
            # continuation (no reset): the second run of `a` starts from another layout / state -> another plan
            set_(sim, a)
            sim.run()
            o.sim(b).sim(a)
            re, im = sim.get_dm()
            ore, oim = o.dm()
            assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
    finally:
        dm.set_option("plan_cache", 1)


@pytest.mark.parametrize("n,tile_bits", [(9, 8), (10, 12), (7, 5)])
def test_sparse_start_matches_dense_execution(dm, oracle_mod, n, tile_bits):
    """After dmb_reset_dm only element 0 is non-zero: the leading sweeps launch just the tiles that can hold non-zeros
    (support tracking, option "sparse").  Bit-identical to launching every tile, also for a second run on the evolved
    state, and equal to the oracle; a circuit that leaves most qubits untouched exercises the partially grown support."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    rng = np.random.default_rng(100 + n)
    cases = [circuits.qft(n), random_gates(n, 40, rng), [("H", [1], 0, 0, 0), ("CX", [1, n - 1], 0, 0, 0), ("T", [0], 0, 0, 0)]]
    for gates in cases:
        outs = []
        for sparse in (1, 0):
            dm.set_option("sparse", sparse)
            dm.set_option("tile_bits", tile_bits)
            try:
                sim = run_gpu(dm, n, gates)
                first = sim.get_dm()
                sim.run()  # second run: dense path, same graph key handling
                outs.append((first, sim.get_dm()))
            finally:
                dm.set_option("sparse", 1); dm.set_option("tile_bits", 12)
        for a, b in zip(outs[0], outs[1]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
        assert max(np.abs(outs[0][0][0] - ore).max(), np.abs(outs[0][0][1] - oim).max()) < TOL


def test_large_n13_properties_and_parity(dm, oracle_mod):
    """n = 13 (1 GiB state): full-matrix parity on a short circuit (oracle needs ~4 GiB and a few seconds)."""
    n = 13
    rng = np.random.default_rng(41)
    gates = random_gates(n, 24, rng, names=["U3", "CX", "H", "T", "RZ", "CU1", "C1", "C2"], with_raw=False)
    sim = run_gpu(dm, n, gates)
    assert abs(sim.trace() - 1.0) < TOL
    assert abs(sim.purity() - 1.0) < 1e-11  # unitary circuit from a pure state
    re, im = sim.get_dm()
    ore, oim = oracle_mod.Oracle(n).sim(gates).dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
    assert np.abs(re - re.T).max() < TOL and np.abs(im + im.T).max() < TOL  # Hermitian


import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["all_ops_n5", "random_mix_n6", "random_c1c2_n6", "srn_n4"])
def test_reference_golden_full_matrix(dm, name):
    """Outputs of the reference itself (tests/golden, produced by make_golden.py from oracle/_ref)."""
    z = np.load(_os.path.join(_GOLD, name + ".npz"))
    n = int(z["n"])
    sim = dm.Simulation(n, 1)
    rec, mats = z["gates"], np.ascontiguousarray(z["mats"])
    dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                       mats.size // 32))
    sim._uploaded = True
    sim.run()
    re, im = sim.get_dm()
    assert np.abs(re - z["real"]).max() < TOL and np.abs(im - z["imag"]).max() < TOL


@pytest.mark.parametrize("name", ["adder_n10", "qft_n10", "vqe_uccsd_n8"])
def test_reference_golden_diagonal(dm, name):
    z = np.load(_os.path.join(_GOLD, name + ".npz"))
    n = int(z["n"])
    sim = dm.Simulation(n, 1)
    rec, mats = z["gates"], np.ascontiguousarray(z["mats"])
    dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                       mats.size // 32))
    sim._uploaded = True
    sim.run()
    assert np.abs(sim.diag() - z["diag"]).max() < TOL
    assert abs(sim.trace() - z["diag"].sum()) < TOL


def test_srn_continuation_two_runs(dm, oracle_mod):
    """After an SRN run the state is not Hermitian: the next run must follow the reference's frame exactly."""
    n = 5
    rng = np.random.default_rng(17)
    a = random_gates(n, 15, rng) + [("SRN", [2], 0, 0, 0)] + random_gates(n, 5, rng)
    b = random_gates(n, 20, rng)
    c = [("SRN", [0], 0, 0, 0)] + random_gates(n, 10, rng)
    sim = dm.Simulation(n, 1)
    o = oracle_mod.Oracle(n)
    for part in (a, b, c):
        rec, mats = dm.pack_gates(part)
        dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                           mats.size // 32))
        sim._uploaded = True
        sim.run()
        o.sim(part)
        re, im = sim.get_dm()
        ore, oim = o.dm()
        assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL
        assert np.abs(sim.diag() - o.diag()).max() < TOL


def test_repeated_runs_of_one_circuit(dm, oracle_mod):
    """bench.py's loop: the same uploaded circuit run again and again on the evolving state."""
    n = 7
    rng = np.random.default_rng(23)
    gates = random_gates(n, 40, rng)
    sim = run_gpu(dm, n, gates)
    sim.run()
    sim.run()
    o = oracle_mod.Oracle(n).sim(gates).sim(gates).sim(gates)
    re, im = sim.get_dm()
    ore, oim = o.dm()
    assert max(np.abs(re - ore).max(), np.abs(im - oim).max()) < TOL


def test_cxx_dropin_example_runs(dm, tmp_path):
    """examples/adder_n10.cpp (the reference's example driver on include/dmsim_b200.hpp) prints 1000000010 x5."""
    import subprocess
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    lib = _os.path.join(root, "dm-sim_b200", "lib")
    exe = tmp_path / "adder"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", _os.path.join(root, "include"),
                    _os.path.join(root, "examples", "adder_n10.cpp"), "-o", str(exe), "-L", lib, "-ldmsim_b200",
                    "-Wl,-rpath," + lib], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("1000000010") == 5 and "nqubits:10, ngates:30" in r.stdout


def test_xacc_plugin_backend_runs(dm, tmp_path):
    """examples/xacc_backend.cpp: the reference's XACC plugin ABI (xacc/DmSimApi.hpp:47-62) on the B200 backend, with the
    checks of xacc/nvidia_omp/tests/DmSimAcceleratorTester.cpp (Bell 0.5/0.5, <Z> of an RX sweep within 0.05)."""
    import subprocess
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    lib = _os.path.join(root, "dm-sim_b200", "lib")
    exe = tmp_path / "xacc"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", _os.path.join(root, "include"),
                    _os.path.join(root, "examples", "xacc_backend.cpp"), "-o", str(exe), "-L", lib, "-ldmsim_b200",
                    "-Wl,-rpath," + lib], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    assert "DM-Sim" not in r.stdout  # the XACC runners silence the per-sim() summary line


def test_cplus_translator_output_runs(dm, tmp_path):
    """OpenQASM -> C++ driver (tool/dmsim_qasm_cplus.py) -> built against the drop-in header -> run: a deterministic
    circuit with a user-defined gate prints the expected basis state 5 times (MSB first, as print_measurement does)."""
    import importlib
    import subprocess
    qasm = importlib.import_module("dm-sim_b200.qasm")
    text = ("OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[4];\ngate flip2 a, b { x a; cx a, b; }\n"
            "flip2 q[0], q[1];\nccx q[0], q[1], q[3];\nh q[2];\nh q[2];\n")
    cpp, stats = qasm.translate_cplus(text, segment=2)
    assert stats["segments"] == 2
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    lib = _os.path.join(root, "dm-sim_b200", "lib")
    src, exe = tmp_path / "c.cpp", tmp_path / "c"
    src.write_text(cpp)
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I", _os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", lib, "-ldmsim_b200", "-Wl,-rpath," + lib], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("1011") == 5, r.stdout   # qubits 0, 1, 3 set


def test_pybind_module_runs_a_generated_script(dm, tmp_path):
    """tool/dmsim_qasm.py output executed with the drop-in pybind11 module, as the reference's workflow does."""
    import subprocess
    import sys
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    q = tmp_path / "c.qasm"
    q.write_text('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[4];\nx q[0];\ncx q[0],q[3];\nccx q[0],q[3],q[2];\n')
    out = tmp_path / "c.py"
    subprocess.run([sys.executable, _os.path.join(root, "tool", "dmsim_qasm.py"), "-i", str(q), "-o", str(out)], check=True)
    script = out.read_text().replace("sim.measure(10)", "print('RESULT', sim.measure(10))")
    out.write_text(script)
    env = dict(_os.environ, PYTHONPATH=root)
    r = subprocess.run([sys.executable, str(out), "4", "1"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "RESULT [13, 13, 13, 13, 13, 13, 13, 13, 13, 13]" in r.stdout  # |1101> : q0, q2, q3 set


# ---- BASELINE.json's full size: 15 qubits, 16 GiB density matrix (the oracle would need 64 GiB and ~10 minutes) -------
def _inverse(gates):
    inv = []
    for g in reversed(gates):
        name = g[0]
        if name in ("H", "X", "CX"):
            inv.append(g)
        elif name == "U1":
            inv.append(("U1", g[1], 0.0, 0.0, -g[4]))
        else:
            raise ValueError(name)
    return inv


@pytest.mark.parametrize("family", ["qft", "bv"])
def test_fullsize_n15_against_statevector(dm, family):
    """qft_n15 / bv_n15 gate for gate (benchmark/*.qasm) on a non-trivial input: trace, purity, ALL 2^15 diagonal
    probabilities and 2^16 random off-diagonal elements against rho^T built from a state-vector run of the same gates."""
    import importlib
    from helpers import statevector
    circuits = importlib.import_module("dm-sim_b200.circuits")
    n = 15
    prefix = [("X", [q], 0.0, 0.0, 0.0) for q in (0, 3, 4, 9, 14)] if family == "qft" else []
    gates = prefix + (circuits.qft(n) if family == "qft" else circuits.bv(n))
    sim = run_gpu(dm, n, gates)
    psi = statevector(n, gates)
    assert abs(sim.trace() - 1.0) < TOL
    assert abs(sim.purity() - 1.0) < 1e-11
    assert np.abs(sim.diag() - np.abs(psi) ** 2).max() < TOL
    rng = np.random.default_rng(15)
    col = rng.integers(0, 1 << n, size=1 << 16, dtype=np.uint64)
    row = rng.integers(0, 1 << n, size=1 << 16, dtype=np.uint64)
    got = sim.elements((col << np.uint64(n)) | row)
    want = psi[row.astype(np.int64)] * np.conj(psi[col.astype(np.int64)])  # res[col][row] = rho[row][col]
    assert np.abs(got - want).max() < TOL
    if family == "bv":  # hidden string all ones, ancilla left in |->: two outcomes with probability 1/2 each
        d = sim.diag()
        assert abs(d[(1 << 14) - 1] - 0.5) < TOL and abs(d[(1 << 15) - 1] - 0.5) < TOL


def test_fullsize_n15_circuit_then_inverse_is_identity(dm):
    """Round trip at full size: qft_n15 followed by its inverse returns |0..0><0..0| exactly (to 1e-12)."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    n = 15
    g = circuits.qft(n)
    sim = run_gpu(dm, n, g + _inverse(g))
    d = sim.diag()
    assert abs(d[0] - 1.0) < TOL and np.abs(d[1:]).max() < TOL
    assert abs(sim.purity() - 1.0) < 1e-11
    rng = np.random.default_rng(16)
    f = rng.integers(1, 1 << (2 * n), size=1 << 14, dtype=np.uint64)
    assert np.abs(sim.elements(f)).max() < TOL
