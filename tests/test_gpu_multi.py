"""-m gpu, needs >= 2 GPUs (skipped otherwise).  Both multi-GPU forms of the engine against the oracle:
  * one process per GPU (torchrun style): qubit-remap exchange inside dmb_run, collective result calls;
  * ONE process driving all GPUs through one handle (the reference's Simulation(n_qubits, n_gpus)): Python mirror,
    C++ drop-in header (examples/adder_n10.cpp run as `./adder_n10 P`), pybind module."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, q):
    try:
        import torch
        import torch.distributed as dist
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import time
        from helpers import random_gates
        dm = importlib.import_module("dm-sim_b200")
        gates = random_gates(n, 60, np.random.default_rng(seed))
        sim = dm.Simulation(n, world, rank=rank, device=rank)
        rec, mats = dm.pack_gates(gates)
        results = []
        for rep in range(2):  # two runs: the second starts from the remapped layout
            dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                               mats.size // 32))
            sim._uploaded = True
            sim.run()
            if rank == world - 1:
                time.sleep(0.3)  # skew: the other ranks read out (and later reset / remap) while this one lags behind
            data, lay = sim.shard()
            # collective result calls: every rank gets the GLOBAL answer
            diag = sim.diag()
            tr = np.array([sim.trace(), sim.purity()])
            re, im = sim.get_dm()
            shots = sim.sample(np.linspace(0.0, 0.999, 64))[0]
            probe = np.arange(0, 4 ** n, 97, dtype=np.uint64)
            el = sim.elements(probe)
            parts = [torch.empty(2 * data.size, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(data.view(np.float64))).cuda())
            same = torch.tensor([float(np.abs(diag).sum()), float(np.abs(re).sum()), float(shots.sum())], device="cuda", dtype=torch.float64)
            lo, hi = same.clone(), same.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "collective results differ between ranks"
            if rank == 0:
                full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
                results.append((full, lay.copy(), diag, tr, dict(sim.last_stats), re + 1j * im, shots, (probe, el)))
        if rank == 0:
            import oracle
            from plan_emulator import physical_to_logical
            o = oracle.Oracle(n)
            errs = []
            for full, lay, diag, tr, st, dmat, shots, (probe, el) in results:
                o.sim(gates)
                re, im = o.dm()
                ref = (re + 1j * im).reshape(-1)
                res = physical_to_logical(full, lay)
                errs.append(float(np.abs(res - ref).max()))
                errs.append(float(np.abs(dmat.reshape(-1) - ref).max()))  # dmb_get_dm through the communicator
                errs.append(float(np.abs(el - ref[probe.astype(np.int64)]).max()))
                errs.append(float(np.abs(diag - o.diag()).max()))
                errs.append(abs(float(tr[0]) - 1.0))
                errs.append(abs(float(tr[1]) - 1.0) / 10)
                want = np.asarray(o.sample_with_r(np.linspace(0.0, 0.999, 64))).astype(np.int64)
                errs.append(float(np.abs(shots.astype(np.int64) - want).max()))
            q.put((max(errs), results[0][4]["n_exchanges"], results[0][4]["comm_ms"]))
        # reset on a skewed rank right after a run, then a third run from |0><0| (the peers' remap stores must wait)
        if rank == 0:
            time.sleep(0.2)
        sim.reset_dm()
        sim.run()
        d3 = sim.diag()
        if rank == 0:
            import oracle
            errs3 = float(np.abs(d3 - oracle.Oracle(n).sim(gates).diag()).max())
            q.put((errs3, 0, 0))
        dist.barrier()
        del sim
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put(("error", repr(e), 0))
        raise


@pytest.mark.parametrize("world,n", [(2, 8), (2, 11), (4, 10), (8, 11)])
def test_sharded_engine_matches_oracle(world, n):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import __graft_entry__ as ge
    ge.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 5 + n, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, n_exch, comm_ms = q.get(timeout=300)
    assert err != "error", n_exch
    err3 = q.get(timeout=300)[0]
    for p in procs:
        p.join(timeout=60)
    assert err != "error", n_exch
    assert err < 1e-12 and err3 < 1e-12
    assert n_exch >= 1


# ---- ONE process, one handle, P devices (reference Simulation(n_qubits, n_gpus), src/dmsim_nvgpu_omp.cuh:196-271) ----
def _need(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")


@pytest.mark.parametrize("world,n", [(2, 8), (2, 11), (4, 10), (8, 11), (8, 5)])
def test_single_process_group_matches_oracle(dm, oracle_mod, world, n):
    _need(world)
    from helpers import random_gates
    rng = np.random.default_rng(40 + n + world)
    gates = random_gates(n, 60, rng)
    sim = dm.Simulation(n, world)
    assert sim.group
    o = oracle_mod.Oracle(n)
    for rep in range(2):  # the second run continues from the remapped layout
        sim.clear_circuit()
        for g in gates:
            sim.append(dm.Gate(g[0], *(list(g[1]) + [0] * (5 - len(g[1]))), theta=g[2], phi=g[3], lam=g[4],
                               matrix=g[5] if len(g) > 5 else None))
        sim.upload()
        sim.run()
        assert sim.last_stats["n_exchanges"] >= 1
        o.sim(gates)
        re, im = o.dm()
        gre, gim = sim.get_dm()
        assert max(np.abs(gre - re).max(), np.abs(gim - im).max()) < 1e-12
        assert np.abs(sim.diag() - o.diag()).max() < 1e-12
        assert abs(sim.trace() - 1.0) < 1e-12 and abs(sim.purity() - 1.0) < 1e-11
        probe = np.arange(0, 4 ** n, 61, dtype=np.uint64)
        assert np.abs(sim.elements(probe) - (re + 1j * im).reshape(-1)[probe.astype(np.int64)]).max() < 1e-12
        r = np.linspace(0.0, 0.999, 50)
        assert np.array_equal(sim.sample(r)[0].astype(np.int64), np.asarray(o.sample_with_r(r)).astype(np.int64))
    # reset, then load the oracle's state and read it back / continue from it
    sim.reset_dm()
    assert sim.diag()[0] == 1.0 and abs(sim.trace() - 1.0) < 1e-15
    sim.set_dm(re, im)
    bre, bim = sim.get_dm()
    assert np.array_equal(bre, re) and np.array_equal(bim, im)
    del sim


@pytest.mark.parametrize("world,n", [(2, 9), (4, 9), (8, 10)])
def test_group_on_specialised_kernels(dm, oracle_mod, world, n):
    """The run-time specialised kernels (csrc/jit.cu) on a sharded state: local sweeps and the fused remap sweep whose
    stores go to the peers' shards, forced on at a size the oracle checks in full."""
    _need(world)
    from helpers import random_gates
    gates = random_gates(n, 60, np.random.default_rng(90 + n + world))
    dm.set_option("jit", 2); dm.set_option("jit_min_bits", 0)
    try:
        sim = dm.Simulation(n, world)
        for g in gates:
            sim.append(dm.Gate(g[0], *(list(g[1]) + [0] * (5 - len(g[1]))), theta=g[2], phi=g[3], lam=g[4],
                               matrix=g[5] if len(g) > 5 else None))
        sim.upload()
        sim.run()
        assert sim.last_stats["n_exchanges"] >= 1
        assert dm.query("jit_sweeps", sim._h) == world * sim.last_stats["n_sweeps"] and dm.query("jit_pending", sim._h) == 0
        re, im = oracle_mod.Oracle(n).sim(gates).dm()
        gre, gim = sim.get_dm()
        assert max(np.abs(gre - re).max(), np.abs(gim - im).max()) < 1e-12
        del sim
    finally:
        dm.set_option("jit", 1); dm.set_option("jit_min_bits", 16)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_reference_example_runs_on_n_gpus(dm, world, tmp_path):
    """examples/adder_n10.cpp (the reference's example/adder_n10_nvgpu_omp.cu with one include changed) as `./adder P`:
    five times 1000000010 (README.md:235-244), from ONE process."""
    _need(world)
    import subprocess
    exe = tmp_path / "adder_n10"
    lib_dir = os.path.join(ROOT, "dm-sim_b200", "lib")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "adder_n10.cpp"),
                    "-L", lib_dir, "-ldmsim_b200", "-Wl,-rpath," + lib_dir, "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(world)], check=True, capture_output=True, text=True, timeout=300).stdout
    assert out.count("1000000010") == 5 and "OK" in out and "ngpus:%d" % world in out.replace(" ", "")


def test_pybind_module_on_two_gpus(dm):
    """The reference's Python surface: Simulation(10, 2) without torchrun (example/adder_n10_omp.py)."""
    _need(2)
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import libdmsim_py_nvgpu_omp as dmsim\n"
        "sim = dmsim.Simulation(10, 2)\n"
        "for q in (1, 5, 6, 7, 8): sim.append(sim.X(q))\n"
        "def maj(a, b, c):\n    sim.append(sim.CX(c, b)); sim.append(sim.CX(c, a)); sim.append(sim.CCX(a, b, c))\n"
        "def unmaj(a, b, c):\n    sim.append(sim.CCX(a, b, c)); sim.append(sim.CX(c, a)); sim.append(sim.CX(a, b))\n"
        "maj(0, 5, 1); maj(1, 6, 2); maj(2, 7, 3); maj(3, 8, 4)\n"
        "sim.append(sim.CX(4, 9))\n"
        "unmaj(3, 8, 4); unmaj(2, 7, 3); unmaj(1, 6, 2); unmaj(0, 5, 1)\n"
        "sim.upload(); sim.run()\n"
        "print('RES', sim.measure(5))\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True, timeout=300).stdout
    assert "RES [514, 514, 514, 514, 514]" in out
