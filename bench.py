#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 density-matrix engine (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is ONE sim() of the workload circuit on the resident density matrix (dmb_run through the C-ABI).
N = 1 : workload qft_n15 (benchmark/qft_n15.qasm gate-for-gate, 15 qubits, 16 GiB state) -- the configuration
        BASELINE.json's metric (gates/s, ms/gate at 15 q on one GPU) is quoted on.
N > 1 : one process per GPU under torchrun; workload random_c1c2_n16 (N = 2, 4) / random_c1c2_n17 (N = 8),
        state sharded on the top log2(N) index bits, NCCL qubit-remap exchange inside the timed region.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference CPU backend (oracle/_ref, the
unmodified dmsim_cpu_omp.hpp compiled in place) on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(name):
    circuits = importlib.import_module("dm-sim_b200.circuits")
    fam, _, n = name.rpartition("_n")
    n = int(n)
    if fam == "qft":
        return n, circuits.qft(n)
    if fam == "bv":
        return n, circuits.bv(n)
    if fam == "adder":
        return 10, circuits.adder_n10()
    if fam == "vqe_uccsd":  # benchmark/vqe_uccsd_n8.qasm (10808 gates): the gate list travels inside the golden fixture
        import numpy as np
        z = np.load(os.path.join(ROOT, "tests", "golden", "vqe_uccsd_n8.npz"))
        names = importlib.import_module("dm-sim_b200").OP_NAMES
        return int(z["n"]), [(names[int(g["op"])], [int(q) for q in g["qb"][:2]], float(g["theta"]), float(g["phi"]),
                              float(g["lam"])) for g in z["gates"]]
    if fam == "random_c1c2":
        return n, circuits.random_c1c2(n, 256)
    if fam == "single":   # one gate = one sweep with 2 ops: the memory pipeline of the sweep kernel
        return n, [("H", [5], 0.0, 0.0, 0.0)]
    if fam == "hlayer":   # one dense 1-qubit gate per qubit
        return n, [("H", [q], 0.0, 0.0, 0.0) for q in range(n)]
    raise SystemExit(f"unknown workload {name}")


def reference_arm(args, name):
    """The reference's own CPU implementation (oracle/_ref) on the host cores, bounded sample."""
    import oracle
    oracle.build()
    kind = "reference" if oracle.have_reference() else "port"
    n, gates = workload(name)
    cores = os.cpu_count() or 1
    n_cpus = 1
    while n_cpus * 2 <= cores:
        n_cpus *= 2
    # bounded sample: the same circuit family at a size the CPU finishes in seconds; cost scales 4x per qubit
    n_s = min(n, args.cpu_sample_qubits)
    fam = name.rpartition("_n")[0]
    _, g_s = workload(f"{fam}_n{n_s}") if fam != "adder" else (n, gates)
    n_cpus = min(n_cpus, 1 << n_s)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    times = []
    for i in range(args.warmup + args.steps):
        if kind == "reference":
            r = oracle.reference_run(n_s, g_s, n_cpus=n_cpus, want_dm=False)
            ms = r["sim_ms"] if r["sim_ms"] > 0 else r["wall_ms"]
        else:
            t0 = time.perf_counter()
            oracle.Oracle(n_s).sim(g_s)
            ms = (time.perf_counter() - t0) * 1e3
        if i >= args.warmup:
            times.append(ms)
    ms_sample = sum(times) / len(times)
    scale = 4.0 ** (n - n_s)
    ms_full = ms_sample * scale * (len(gates) / len(g_s))
    value = len(gates) / (ms_full * 1e-3)
    sample = (f"{fam}_n{n_s} ({len(g_s)} gates) via dmsim_cpu_omp n_cpus={n_cpus}: {ms_sample:.1f} ms/sim; scaled x4^{n - n_s}"
              f" and x{len(gates)}/{len(g_s)} gates to {name}")
    line = {"metric": "gates/sec", "value": value, "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_full, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": name, "n_qubits": n, "n_gates": len(gates)},
            "cpu_baseline": {"value": value, "unit": "gates/s", "cores": n_cpus, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample-qubits", type=int, default=13)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = {1: "qft_n15", 2: "random_c1c2_n16", 4: "random_c1c2_n16", 8: "random_c1c2_n17"}.get(args.gpus, "random_c1c2_n16")

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, args.workload)), flush=True)
        return

    import numpy as np
    import torch
    import __graft_entry__ as ge
    if local_rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    else:
        torch.cuda.set_device(0)
    dm = importlib.import_module("dm-sim_b200")
    n, gates = workload(args.workload)
    rec, mats = dm.pack_gates(gates)
    sim = dm.Simulation(n, world, rank=rank, device=local_rank) if world > 1 else dm.Simulation(n, 1)
    if world > 1 and sim.n_gpus > 1 and not hasattr(sim, "_comm_done"):
        pass
    L = dm.lib()

    def set_circuit():
        dm._check(L.dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32))
        sim._uploaded = True

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # `value` / `roofline` are measured on the DENSE resident state: every tile of every sweep is launched (after a reset
    # the engine would otherwise skip the tiles that are still all-zero, see the e2e leg below)
    dm.set_option("sparse", 0)
    sim.reset_dm()
    set_circuit()
    for _ in range(args.warmup):
        sim.run()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, comm_ms, launches, sweeps, exch = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        sim.run()
        st = sim.last_stats
        dev_ms += st["sim_ms"]; comm_ms += st["comm_ms"]
        launches += st["n_launches"]; sweeps += st["n_sweeps"]; exch += st["n_exchanges"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    st = sim.last_stats
    if dist is not None:
        t = torch.tensor([dev_ms, comm_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, comm_ms, wall_ms = (float(x) for x in t.tolist())

    # ---- end to end through the C-ABI with HOST buffers: reset + circuit upload (H2D) + run + diagonal (D2H)
    diag = np.empty(1 << n)
    dm.set_option("sparse", 1)
    e2e_ms, parts = [], [0.0, 0.0, 0.0, 0.0]
    for i in range(2 + min(args.steps, 3)):
        barrier()
        t1 = time.perf_counter()
        sim.reset_dm()
        t2 = time.perf_counter()
        set_circuit()
        t3 = time.perf_counter()
        sim.run()
        t4 = time.perf_counter()
        dm._check(L.dmb_get_diag(sim._h, diag.ctypes.data))
        barrier()
        t5 = time.perf_counter()
        if i >= 2:
            e2e_ms.append((t5 - t1) * 1e3)
            for j, dt in enumerate((t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                parts[j] += dt * 1e3
    e2e = sum(e2e_ms) / len(e2e_ms)
    parts = [x / len(e2e_ms) for x in parts]
    # the same call sequence WITHOUT the reset: the circuit is re-uploaded and applied to the resident (dense) state, so
    # that no tile is skipped as still-zero
    dm.set_option("sparse", 0)
    sim.reset_dm()
    res_ms = []
    for i in range(1 + min(args.steps, 3)):
        barrier()
        t1 = time.perf_counter()
        set_circuit()
        sim.run()
        dm._check(L.dmb_get_diag(sim._h, diag.ctypes.data))
        barrier()
        if i >= 1:
            res_ms.append((time.perf_counter() - t1) * 1e3)
    e2e_res = sum(res_ms) / len(res_ms)
    if dist is not None:
        t = torch.tensor([e2e, e2e_res], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e, e2e_res = (float(x) for x in t.tolist())
        d = torch.from_numpy(diag).cuda()
        dist.all_reduce(d)
        diag = d.cpu().numpy()
    trace = float(diag.sum())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    ms_step = dev_ms / args.steps
    n_gates = len(gates)
    sweep_bytes = st["sweep_bytes"]
    comp_ms = dev_ms - comm_ms
    achieved = (sweeps * sweep_bytes) / (comp_ms * 1e-3) / 1e9 if comp_ms > 0 else 0.0
    line = {
        "metric": "gates/sec", "value": n_gates / (ms_step * 1e-3), "unit": "gates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_gate": ms_step / n_gates,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "n_qubits": n, "n_gates": n_gates, "n_primitives": st["n_primitives"],
                   "fused_blocks_per_side": st["n_blocks"], "sweeps_per_step": st["n_sweeps"],
                   "exchanges_per_step": st["n_exchanges"], "state_bytes": 16 * 4 ** n,
                   "state_bytes_per_gpu": 16 * 4 ** n // world,
                   "scaling_note": "BASELINE.json configs: 15 q on 1 GPU, 16 q on 2/4 GPUs, 17 q on 8 GPUs (16-32 GiB per GPU)",
                   "l2_policy": "state per GPU (>= 16 GiB at the named sizes) is far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"shard on top {world.bit_length() - 1} index bits" if world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "kernel": "sweep_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_8TBs_nominal": achieved / 8000.0, "peak_source": peak_src,
                     "avg_launch_ms": comp_ms / max(1, sweeps), "bytes_per_launch": sweep_bytes,
                     "traffic": _ncu_traffic(args.workload)},
        "e2e": {"value": n_gates / (e2e * 1e-3), "unit": "gates/s", "ms_per_step": e2e,
                "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": 8 * (1 << n),
                "host_call_ms": {"reset_dm": parts[0], "set_circuit": parts[1], "run": parts[2], "get_diag": parts[3]},
                "resident_state": {"value": n_gates / (e2e_res * 1e-3), "ms_per_step": e2e_res,
                                   "what": "same calls without dmb_reset_dm: circuit applied to the resident dense state"},
                "what": "host gate list in, host diagonal out, one circuit from |0><0| as the reference's sim(): dmb_reset_dm + "
                        "dmb_set_circuit (plan + H2D of the device op tables) + dmb_run + dmb_get_diag (D2H of the 2^n "
                        "probabilities), wall clock.  After a reset the leading sweeps launch only the tiles that can be "
                        "non-zero (single GPU; DESIGN.md 'sparse start'); `value` and `resident_state` run on the dense state"},
        # size-independent rate (gates/s shrinks 4x per added qubit): gates x density-matrix elements updated per second
        "work_rate": {"value": n_gates * float(4 ** n) / (ms_step * 1e-3), "unit": "gate x element updates/s"},
        "gpu_launches": int(launches), "wall_ms_per_step": wall_ms / args.steps,
        "trace_after_run": trace, "clocks": clocks,
    }
    if world > 1:
        line["comm"] = {"ms_per_step": comm_ms / args.steps, "bytes_sent_per_rank_per_step": st["exchange_bytes"],
                        "GBps_per_direction": (st["exchange_bytes"] / 1e9) / (comm_ms / args.steps * 1e-3) if comm_ms > 0 else None}
    if world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_arm(argparse.Namespace(gpus=1, steps=1, warmup=0, cpu_sample_qubits=args.cpu_sample_qubits), args.workload)
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:  # the checker must never take the product's number down
            line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per sweep_kernel launch from the committed `ncu --set full` capture
    of this workload (profiles/ncu_traffic.json), or None when it has not been captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


if __name__ == "__main__":
    main()
