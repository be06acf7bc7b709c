"""CPU: the C++ host pipeline (expand -> fuse -> schedule -> encode) checked against the oracle and the reference's
golden vectors WITHOUT a GPU, through two numpy interpreters of what the planner emits:
  plan_emulator   : executes the schedule with the full op matrices (scheduling / fusion / remap bookkeeping),
  kernel_emulator : walks the device tables exactly as sweep_kernel indexes them (encoder + addressing scheme)."""
import os

import numpy as np
import pytest

import kernel_emulator as ke
import plan_emulator as pe
from helpers import each_op_once, random_gates, to_complex
from test_oracle import FULL, load_golden

TOL = 1e-12


def zero_state(n):
    v = np.zeros(4 ** n, dtype=np.complex128)
    v[0] = 1.0
    return v


@pytest.fixture
def opts(dm):
    def set_(**kw):
        for k, v in kw.items():
            dm.set_option(k, v)
    yield set_
    dm.set_option("tile_bits", 12); dm.set_option("low_bits", 3); dm.set_option("min_tiles_log2", 10)


@pytest.mark.parametrize("name", FULL)
@pytest.mark.parametrize("world", [1, 2, 4])
def test_schedule_reproduces_reference_golden(dm, name, world):
    z, packed = load_golden(name)
    n = int(z["n"])
    plan = dm.plan_json(n, world, packed)
    ref = to_complex(z["real"], z["imag"])
    assert np.abs(pe.run_plan(plan, zero_state(n)) - ref).max() < TOL
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - ref).max() < TOL


@pytest.mark.parametrize("n,world,o", [
    (1, 1, {}), (2, 1, {}), (3, 1, {}), (4, 1, {}), (6, 1, {}), (7, 1, {}),
    (7, 1, dict(tile_bits=9, low_bits=2, min_tiles_log2=2)),
    (6, 1, dict(tile_bits=6, low_bits=0, min_tiles_log2=2)),
    (4, 1, dict(tile_bits=4, low_bits=1, min_tiles_log2=2)),
    (6, 2, {}), (6, 4, {}), (7, 8, dict(tile_bits=7, low_bits=1)), (7, 4, dict(tile_bits=8)), (5, 8, {}), (3, 8, {}),
])
def test_random_circuits_every_geometry(dm, oracle_mod, opts, n, world, o):
    opts(**o)
    rng = np.random.default_rng(1000 + 10 * n + world)
    gates = random_gates(n, 40, rng, exclude=())  # includes SRN (full-barrier path) and raw C1/C2
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    ref = to_complex(re, im)
    plan = dm.plan_json(n, world, gates)
    assert np.abs(pe.run_plan(plan, zero_state(n)) - ref).max() < TOL
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - ref).max() < TOL
    if world > 1:
        assert plan["n_exchanges"] >= 1


@pytest.mark.parametrize("n,world,o", [
    (7, 1, dict(min_tiles_log2=2)), (7, 2, dict(min_tiles_log2=1)), (7, 1, dict(min_tiles_log2=2, tma_box_bits=12)),
    (7, 1, dict(min_tiles_log2=2, tma_box_bits=7)), (7, 1, dict(min_tiles_log2=2, tma=0)),
])
def test_full_size_tiles_tma_addressing(dm, oracle_mod, opts, n, world, o):
    """k = 12 tiles: TMA boxes (5-D tensor view of the shard, enumerated copies), the hardware 128-byte swizzle and the
    lane-bit choice that goes with it -- walked by the kernel emulator exactly as the device does."""
    opts(**o)
    try:
        rng = np.random.default_rng(77 + world)
        gates = random_gates(n, 60, rng, exclude=("SRN",))
        re, im = oracle_mod.Oracle(n).sim(gates).dm()
        plan = dm.plan_json(n, world, gates)
        sweeps = [st for st in plan["steps"] if st["kind"] == "sweep"]
        assert any(st["k"] == 12 for st in sweeps)
        want_tma = o.get("tma", 1)
        for st in sweeps:
            if st["k"] == 12:
                assert st["dev"]["tma_load"] == want_tma and st["dev"]["swz"] == want_tma
                if want_tma:
                    g = st["dev"]["tma"]
                    assert g["n_copies"] * g["box_bytes"] == 16 << 12 and g["start"][0] == 0 and g["box_log2"][0] == 3
                    assert all(b <= 8 for b in g["box_log2"])
                    assert st["dev"]["tma_store"] == int(st["in_pos"] == st["out_pos"] and not st["out_of_place"])
        assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL
    finally:
        dm.set_option("tma", 1); dm.set_option("tma_box_bits", 10)


def test_each_op_alone(dm, oracle_mod):
    rng = np.random.default_rng(3)
    n = 5
    prefix = random_gates(n, 8, rng, names=["U3", "CX", "H"], with_raw=False)
    for g in each_op_once(n, rng) + [("SRN", [2], 0, 0, 0), ("ID", [1], 0, 0, 0)]:
        gates = prefix + [g]
        re, im = oracle_mod.Oracle(n).sim(gates).dm()
        out = ke.run_plan_dev(dm.plan_json(n, 1, gates), zero_state(n))
        assert np.abs(out - to_complex(re, im)).max() < TOL, g[0]


def test_continuation_from_permuted_layout(dm, oracle_mod):
    """A second circuit planned from the layout (and conjugation flag) the first one left behind."""
    rng = np.random.default_rng(8)
    n, world = 6, 4
    a = random_gates(n, 30, rng) + [("SRN", [1], 0, 0, 0)]
    b = random_gates(n, 30, rng)
    p1 = dm.plan_json(n, world, a)
    v1 = ke.run_plan_dev(p1, zero_state(n))
    assert p1["end_layout"] != list(range(2 * n)) and p1["conj_end"]
    p2 = dm.plan_json(n, world, b, start_layout=p1["end_layout"], conj_state=p1["conj_end"], non_hermitian=True)
    v2 = ke.run_plan_dev(p2, v1)
    re, im = oracle_mod.Oracle(n).sim(a).sim(b).dm()
    assert np.abs(v2 - to_complex(re, im)).max() < TOL


def test_one_exchange_per_run_like_the_reference(dm):
    """The reference needs exactly one all-to-all per sim(); so does the planner from reset."""
    rng = np.random.default_rng(9)
    for n, world in [(8, 2), (8, 4), (9, 8)]:
        gates = random_gates(n, 60, rng)
        assert dm.plan_json(n, world, gates)["n_exchanges"] == 1


def test_fusion_and_sweep_counts_of_named_workloads(dm):
    import importlib
    C = importlib.import_module("dm-sim_b200.circuits")
    p = dm.plan_json(10, 1, C.adder_n10())
    assert (p["n_gates"], p["n_primitives"]) == (30, 142)  # example/adder_n10: 30 Gate objects = 142 primitives
    p = dm.plan_json(15, 1, C.qft(15))
    assert p["n_primitives"] == 540 and p["n_sweeps"] <= 10  # the reference needs 2 x 540 HBM sweeps
    p = dm.plan_json(15, 1, C.bv(15))
    assert p["n_primitives"] == 44 and p["n_sweeps"] <= 4
    for st in p["steps"]:
        assert st["in_pos"][:3] == [0, 1, 2], "every tile keeps the 3 lowest physical bits: >= 128 B HBM runs"


def test_invalid_arguments_are_rejected(dm):
    with pytest.raises(dm.DMSimError, match="out of range"):
        dm.plan_json(3, 1, [("H", [3], 0, 0, 0)])
    with pytest.raises(dm.DMSimError, match="out of range"):
        dm.plan_json(3, 1, [("H", [0, 7], 0, 0, 0)])  # append() asserts every qb < n_qubits, used or not
    with pytest.raises(dm.DMSimError, match="repeated"):
        dm.plan_json(3, 1, [("CX", [1, 1], 0, 0, 0)])
    with pytest.raises(dm.DMSimError, match="power of two"):
        dm.plan_json(3, 3, [("H", [0], 0, 0, 0)])
    with pytest.raises(dm.DMSimError, match="divide"):
        dm.plan_json(2, 8, [("H", [0], 0, 0, 0)])
    with pytest.raises(dm.DMSimError, match="unknown op"):
        rec, mats = dm.pack_gates([("H", [0], 0, 0, 0)])
        rec[0]["op"] = 77
        dm.plan_json(3, 1, (rec, mats))
    assert dm.plan_json(3, 1, [])["n_sweeps"] == 0  # empty circuit


def test_large_op_counts_split_into_sweeps(dm, oracle_mod):
    """More ops than the kernel's shared-memory op table holds: the planner must split the sweep."""
    rng = np.random.default_rng(4)
    n = 4
    gates = random_gates(n, 400, rng)
    plan = dm.plan_json(n, 1, gates)
    # controlled phases (cls 7) merge into star ops and count 1/8 towards the table, at most 160 of them per sweep
    for st in plan["steps"]:
        n_cp = sum(1 for o in st["ops"] if o["cls"] == 7)
        assert 8 * (len(st["ops"]) - n_cp) + n_cp <= 8 * 112 and n_cp <= 160
    assert plan["n_sweeps"] > 1
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL


@pytest.mark.parametrize("n,world,o", [
    (5, 1, {}), (7, 1, {}), (7, 1, dict(tile_bits=8, low_bits=2, min_tiles_log2=2)),
    (6, 1, dict(tile_bits=5, low_bits=0, min_tiles_log2=2)), (7, 1, dict(tile_bits=12, low_bits=3, min_tiles_log2=1)),
    (6, 2, {}), (7, 4, dict(tile_bits=8)), (7, 8, dict(tile_bits=7, low_bits=1)), (6, 8, {}),
])
def test_controlled_phase_stars(dm, oracle_mod, opts, n, world, o):
    """Controlled phases run with ONE bit in the tile / in the register round (CLS_CPHASE, RC_STAR): the partner bit may be
    a lane / warp / iteration bit, a bit outside the tile or a rank bit.  QFT plus a mix of diagonal 2-qubit gates."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    opts(**o)
    rng = np.random.default_rng(77 + n + world)
    gates = circuits.qft(n)
    for _ in range(30):
        a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
        kind = rng.integers(6)
        th = float(rng.uniform(-3, 3))
        gates.append([("CU1", [a, b], 0, 0, th), ("CZ", [a, b], 0, 0, 0), ("CRZ", [a, b], 0, 0, th), ("RZZ", [a, b], th, 0, 0),
                      ("H", [a], 0, 0, 0), ("U3", [b], th, 0.3, -0.2)][kind])
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    ref = to_complex(re, im)
    plan = dm.plan_json(n, world, gates)
    n_cp = sum(1 for st in plan["steps"] if st["kind"] == "sweep" for op in st["ops"] if op["cls"] == 7)
    assert n_cp > 0
    assert np.abs(pe.run_plan(plan, zero_state(n)) - ref).max() < TOL
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - ref).max() < TOL
    # the same circuit with the controlled-phase scheduling switched off gives the same state
    dm.set_option("cphase", 0)
    try:
        plan0 = dm.plan_json(n, world, gates)
        assert sum(1 for st in plan0["steps"] if st["kind"] == "sweep" for op in st["ops"] if op["cls"] == 7) == 0
        assert np.abs(ke.run_plan_dev(plan0, zero_state(n)) - ref).max() < TOL
    finally:
        dm.set_option("cphase", 1)


def test_deferred_diagonals_and_new_device_ops(dm, oracle_mod):
    """Pending diagonal factors materialise only when a non-diagonal op needs their register bit: QFT rounds become
    HAD / CP2 / STAR, controlled diagonals RC_DIAGP, Hadamard scales fold across the sweep (RC_HAD), and rounds with
    more than 16 device ops are split.  Every variant is checked against the oracle through the kernel mirror."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    RC = dict(DIAGR=7, RR=8, STAR=10, HAD=11, DIAGP=12, CP2=13)
    rng = np.random.default_rng(5)
    # (the shapes asserted below are those of rounds planned from the front of the op list, the default)
    dm.set_option("heavy_last", 0)
    _deferred_diagonals_body(dm, oracle_mod, circuits, RC, rng)


def _deferred_diagonals_body(dm, oracle_mod, circuits, RC, rng):

    def codes(plan):
        return [o["code"] for st in plan["steps"] if st["kind"] == "sweep" for o in st["dev"]["ops"]]

    # QFT: butterflies + pair phases + stars; at most one scaled H per sweep carries the folded scales
    n = 7
    gates = circuits.qft(n)
    plan = dm.plan_json(n, 1, gates)
    c = codes(plan)
    assert RC["HAD"] in c and RC["CP2"] in c and RC["STAR"] in c
    for st in plan["steps"]:
        assert sum(1 for o in st["dev"]["ops"] if o["code"] == RC["RR"]) <= 1
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL

    # phases controlled by one qubit with several partners inside a round -> RC_DIAGP; T/S/Z on top -> RC_DIAGR
    n = 4
    gates = [("H", [q], 0, 0, 0) for q in range(n)]
    gates += [("CU1", [3, q], 0, 0, 0.3 + 0.2 * q) for q in range(3)] + [("H", [3], 0, 0, 0)]
    gates += [("T", [0], 0, 0, 0), ("S", [1], 0, 0, 0), ("CZ", [0, 1], 0, 0, 0), ("U1", [2], 0, 0, 0.7), ("H", [0], 0, 0, 0)]
    plan = dm.plan_json(n, 1, gates)
    c = codes(plan)
    assert RC["DIAGP"] in c
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL

    # 24 non-commuting 2-qubit blocks cycling through the pairs of 3 qubits: their L parts share one register round
    # (3 register bits) with more than 16 device ops, which the encoder splits
    n = 3
    gates = []
    for i in range(24):
        a, b = [(0, 1), (1, 2), (2, 0)][i % 3]
        gates.append(("CU3", [a, b], float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2))))
    plan = dm.plan_json(n, 1, gates)
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL
    split = False
    for st in plan["steps"]:
        rounds = st["dev"]["rounds"]
        for r in rounds:
            assert r["count"] <= 16
        split |= any(a["roff"] == b["roff"] and a["lane_tab"] == b["lane_tab"] and a["count"] == 16 for a, b in zip(rounds, rounds[1:]))
    assert split


@pytest.mark.parametrize("world", [1, 2])
def test_hadamard_cx_fans_become_controlled_phases(dm, oracle_mod, world):
    """rewrite_hadamard_cx: H_t CX(c_i,t).. H_t -> CZ(c_i,t)..  and  H_t CX(c_i,t).. -> CZ(c_i,t).. H_t for fans with
    two or more controls (Bernstein-Vazirani); a monomial 1-qubit gate is kept out of a diagonal 2-qubit block.  Same
    state as the oracle, and the device programs contain controlled phases instead of register permutations."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    n = 6
    dm.set_option("move_h", 1)   # the second identity is off by default (see plan.cpp)
    cases = {
        "bv": circuits.bv(n),
        "closed_fan": [("H", [q], 0, 0, 0) for q in range(n)] + [("CX", [q, n - 1], 0, 0, 0) for q in range(3)] +
                      [("H", [n - 1], 0, 0, 0), ("T", [n - 1], 0, 0, 0), ("H", [0], 0, 0, 0)],
        "fan_then_control": [("H", [2], 0, 0, 0), ("CX", [0, 2], 0, 0, 0), ("CX", [1, 2], 0, 0, 0), ("CX", [2, 3], 0, 0, 0),
                             ("RY", [2], 0.4, 0, 0), ("H", [0], 0, 0, 0)],
        "single_cx_untouched": [("H", [1], 0, 0, 0), ("CU1", [0, 1], 0, 0, 0.7), ("X", [1], 0, 0, 0), ("CZ", [0, 1], 0, 0, 0)],
    }
    for name, gates in cases.items():
        re, im = oracle_mod.Oracle(n).sim(gates).dm()
        ref = to_complex(re, im)
        plan = dm.plan_json(n, world, gates)
        assert np.abs(pe.run_plan(plan, zero_state(n)) - ref).max() < TOL, name
        assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - ref).max() < TOL, name
        dev_codes = [o["code"] for st in plan["steps"] if st["kind"] == "sweep" for o in st["dev"]["ops"]]
        n_cp = sum(1 for st in plan["steps"] if st["kind"] == "sweep" for op in st["ops"] if op["cls"] == 7)
        if name in ("bv", "closed_fan"):
            assert n_cp >= 3 and 6 not in dev_codes, (name, dev_codes)   # 6 = RC_PERM2
    # with the controlled-phase machinery switched off the circuits are left as they are
    dm.set_option("move_h", 0)
    dm.set_option("cphase", 0)
    try:
        plan0 = dm.plan_json(n, 1, cases["bv"])
        assert 6 in [o["code"] for st in plan0["steps"] for o in st["dev"]["ops"]]
        re, im = oracle_mod.Oracle(n).sim(cases["bv"]).dm()
        assert np.abs(ke.run_plan_dev(plan0, zero_state(n)) - to_complex(re, im)).max() < TOL
    finally:
        dm.set_option("cphase", 1)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_hot_bits_move_to_the_low_tile_positions(dm, oracle_mod, opts, world):
    """Option "hot_low": a bit that still has ops needing it inside a tile swaps into one of the always-in-tile low
    physical positions when the bit sitting there is finished (in-place permutation in the sweep's store phase).  Same
    state as the oracle through both interpreters, never more sweeps than without, fewer for a shared-target circuit."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    n = 7
    opts(tile_bits=6, low_bits=2, min_tiles_log2=2)
    rng = np.random.default_rng(3 + world)
    for gates in (circuits.bv(n), random_gates(n, 80, rng), circuits.qft(n) + circuits.bv(n)):
        re, im = oracle_mod.Oracle(n).sim(gates).dm()
        ref = to_complex(re, im)
        sweeps = {}
        for hot in (1, 0):
            dm.set_option("hot_low", hot)
            try:
                plan = dm.plan_json(n, world, gates)
            finally:
                dm.set_option("hot_low", 1)
            sweeps[hot] = plan["n_sweeps"]
            permuting = [st for st in plan["steps"] if st["kind"] == "sweep" and st["in_pos"] != st["out_pos"] and not st["out_of_place"]]
            assert hot or not permuting
            for st in permuting:
                assert sorted(st["in_pos"]) == sorted(st["out_pos"])
            assert np.abs(pe.run_plan(plan, zero_state(n)) - ref).max() < TOL
            assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - ref).max() < TOL
        if world == 1:
            assert sweeps[1] <= sweeps[0]
    dm.set_option("hot_low", 1)
    p1 = dm.plan_json(n, 1, circuits.bv(n))
    dm.set_option("hot_low", 0)
    try:
        p0 = dm.plan_json(n, 1, circuits.bv(n))
    finally:
        dm.set_option("hot_low", 1)
    assert p1["n_sweeps"] < p0["n_sweeps"]


@pytest.mark.parametrize("family,n", [("qft", 7), ("random", 7), ("allops", 6)])
def test_rounds_planned_from_the_back(dm, oracle_mod, family, n):
    """Option heavy_last: a sweep's rounds chosen from the END of its op list (the fullest round runs last) give the same state."""
    import importlib
    circuits = importlib.import_module("dm-sim_b200.circuits")
    rng = np.random.default_rng(123)
    gates = circuits.qft(n) if family == "qft" else (circuits.random_c1c2(n, 60, seed=4) if family == "random" else random_gates(n, 70, rng, exclude=("SRN",)))
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    dm.set_option("heavy_last", 1)
    try:
        plan = dm.plan_json(n, 1, gates)
    finally:
        dm.set_option("heavy_last", 0)
    assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL


@pytest.mark.parametrize("n,world,o", [(7, 4, dict(tile_bits=6, low_bits=2, min_tiles_log2=2)), (8, 4, dict(tile_bits=8, low_bits=3, min_tiles_log2=2)),
                                       (8, 8, dict(tile_bits=7, low_bits=2, min_tiles_log2=2))])
def test_remap_pack_sweep_enumerates_rank_bits_first(dm, oracle_mod, opts, n, world, o):
    """Pack sweep of a qubit remap: the rank-selecting top local bits that are not tile bits are the LOWEST bits of the tile id
    (concurrent CTAs store to different peers); the enumeration is the same on the load and the store side and the result is
    unchanged (kernel mirror vs oracle), with the option on and off."""
    rng = np.random.default_rng(900 + n + world)
    gates = random_gates(n, 60, rng, exclude=("SRN",))
    re, im = oracle_mod.Oracle(n).sim(gates).dm()
    g = world.bit_length() - 1
    M = 2 * n - g
    seen_spread = False
    for spread in (1, 0):
        opts(spread_peers=spread, **o)
        try:
            plan = dm.plan_json(n, world, gates)
        finally:
            dm.set_option("spread_peers", 1)
        assert plan["n_exchanges"] >= 1
        for st in plan["steps"]:
            if st["kind"] != "sweep":
                continue
            d = st["dev"]
            cin, cout = d["cin"][:d["n_comp"]], d["cout"][:d["n_comp"]]
            assert cin == cout and sorted(cin) == sorted(set(range(M)) - set(st["in_pos"]))
            top = [p for p in cin if p >= M - g]
            if st["out_of_place"] and spread and top:
                assert cin[:len(top)] == top and cin[len(top):] == sorted(cin[len(top):])
                seen_spread = seen_spread or cin != sorted(cin)
            else:
                assert cin == sorted(cin) or (st["out_of_place"] and spread)
        assert np.abs(ke.run_plan_dev(plan, zero_state(n)) - to_complex(re, im)).max() < TOL
    if (n, world) == (7, 4):
        assert seen_spread, "this case must exercise a reordered enumeration (a rank-selecting bit outside the tile)"
