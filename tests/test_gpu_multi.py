"""-m gpu, needs >= 2 GPUs (skipped otherwise): one process per GPU, NCCL qubit-remap exchange inside dmb_run,
result gathered from the shards and compared with the oracle."""
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
            data, lay = sim.shard()
            diag = torch.from_numpy(sim.diag()).cuda()
            dist.all_reduce(diag)
            tr = torch.tensor([sim.trace(), sim.purity()], device="cuda", dtype=torch.float64)
            dist.all_reduce(tr)
            parts = [torch.empty(2 * data.size, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(data.view(np.float64))).cuda())
            if rank == 0:
                full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
                results.append((full, lay.copy(), diag.cpu().numpy(), tr.cpu().numpy(), dict(sim.last_stats)))
        if rank == 0:
            import oracle
            from plan_emulator import physical_to_logical
            o = oracle.Oracle(n)
            errs = []
            for full, lay, diag, tr, st in results:
                o.sim(gates)
                re, im = o.dm()
                res = physical_to_logical(full, lay)
                errs.append(float(np.abs(res - (re + 1j * im).reshape(-1)).max()))
                errs.append(float(np.abs(diag - o.diag()).max()))
                errs.append(abs(float(tr[0]) - 1.0))
                errs.append(abs(float(tr[1]) - 1.0) / 10)
            q.put((max(errs), results[0][4]["n_exchanges"], results[0][4]["comm_ms"]))
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
    for p in procs:
        p.join(timeout=60)
    assert err != "error", n_exch
    assert err < 1e-12
    assert n_exch >= 1
