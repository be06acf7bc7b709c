"""CPU, world_size 2 and 4 over gloo: the multi-rank path -- every rank plans the same circuit, owns ONE shard of the
state (top log2 P physical bits = rank), runs the sweeps on its shard (numpy mirror of the kernel) and performs the
qubit-remap exchange as a real all-to-all between processes with exactly the chunking dmb_run uses (chunk p of rank r
-> chunk r of rank p).  Rank 0 gathers the shards and checks them against the oracle."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import kernel_emulator as ke
        from helpers import random_gates
        from plan_emulator import logical_to_physical
        dm = importlib.import_module("dm-sim_b200")
        gates = random_gates(n, 50, np.random.default_rng(seed))
        plan = dm.plan_json(n, world, gates)
        g = plan["g"]
        M = 2 * n - g
        v0 = np.zeros(4 ** n, dtype=np.complex128)
        v0[0] = 1.0
        shard = logical_to_physical(v0, plan["start_layout"])[rank << M:(rank + 1) << M].copy()
        n_exch = 0
        for st in plan["steps"]:
            if st["kind"] == "exchange":
                chunk = (1 << M) // world
                send = torch.from_numpy(np.ascontiguousarray(shard.view(np.float64)))
                recv = torch.empty_like(send)
                reqs = []
                for p in range(world):
                    if p == rank:
                        recv[2 * p * chunk:2 * (p + 1) * chunk] = send[2 * p * chunk:2 * (p + 1) * chunk]
                        continue
                    reqs.append(dist.isend(send[2 * p * chunk:2 * (p + 1) * chunk].clone(), dst=p))
                    reqs.append(dist.irecv(recv[2 * p * chunk:2 * (p + 1) * chunk], src=p))
                for r in reqs:
                    r.wait()
                shard = recv.numpy().view(np.complex128).copy()
                n_exch += 1
            else:
                out = np.full_like(shard, np.nan) if st["out_of_place"] else None
                shard = ke.run_sweep(st["dev"], shard, out)
        parts = [torch.empty(2 << M, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(np.ascontiguousarray(shard.view(np.float64))), parts, dst=0)
        if rank == 0:
            from plan_emulator import physical_to_logical
            import oracle
            full = np.concatenate([p.numpy().view(np.complex128) for p in parts])
            res = physical_to_logical(full, plan["end_layout"])
            if plan["conj_end"]:
                res = np.conj(res)
            re, im = oracle.Oracle(n).sim(gates).dm()
            err = float(np.abs(res - (re + 1j * im).reshape(-1)).max())
            q.put((err, n_exch, plan["n_sweeps"]))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put(("error", repr(e), 0))
        raise


@pytest.mark.parametrize("world,n", [(2, 5), (2, 7), (4, 6)])
def test_sharded_run_with_real_exchange(world, n):
    import __graft_entry__ as ge
    ge.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 77 + n, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, n_exch, n_sweeps = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert err != "error", n_exch
    assert err < 1e-12
    assert n_exch >= 1 and n_sweeps >= 2
