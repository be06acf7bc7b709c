#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 density-matrix engine (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is ONE sim() of the workload circuit on the resident density matrix (dmb_run through the C-ABI).
N = 1 : workload qft_n15 (benchmark/qft_n15.qasm gate-for-gate, 15 qubits, 16 GiB state) -- the configuration
        BASELINE.json's metric (gates/s, ms/gate at 15 q on one GPU) is quoted on; the line also carries
        `extra_workloads` (the other single-GPU configs of BASELINE.json, each measured outside the timed region).
N > 1 : one process per GPU under torchrun; workload random_c1c2_n16 (N = 2, 4) / random_c1c2_n17 (N = 8), state
        sharded on the top log2(N) index bits, qubit-remap exchange inside the timed region.  Every N > 1 line carries
        `parity` (an 11-qubit all-op circuit against the oracle on the N ranks, and the TIMED workload itself against a
        2^n state-vector run: all diagonal probabilities, 2^16 elements, purity) and `strong_scaling` (random_c1c2_n16
        on ONE GPU, measured by rank 0 in the same run, against the same circuit on N GPUs).
Prints ONE JSON line (rank 0).  `--impl reference` times the reference CPU backend (oracle/_ref, the unmodified
dmsim_cpu_omp.hpp compiled in place) on the host cores on a bounded sample of the same workload.
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

# FP64 pipe of one B200 as probed by tools/fp64probe.cu (profiles/r1_fp64probe.txt): 62.7 DFMA per clock per SM x 148 SMs
# x 1.965 GHz = 18.2e12 FP64 instructions/s = 36.5 TFLOP/s counting an FMA as two
FP64_TFLOPS, FP64_SRC = 36.5, "profiles/r1_fp64probe.txt (DFMA 62.7 per clock per SM, 148 SMs, 1965 MHz)"
PARITY_TOL = 1e-12


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
    if fam == "vqe_uccsd":  # benchmark/vqe_uccsd_n8.qasm (10808 gates)
        return 8, circuits.vqe_uccsd_n8()
    if fam == "random_c1c2":
        return n, circuits.random_c1c2(n, 256)
    if fam == "allops":   # every op of enum OP + raw C1 / C2 on random qubits (the parity circuit of the N > 1 lines)
        return n, circuits.random_allops(n, 60)
    if fam == "single":   # one gate = one sweep with 2 ops: the memory pipeline of the sweep kernel
        return n, [("H", [5], 0.0, 0.0, 0.0)]
    if fam == "hstride":  # ONE sweep over a strided tile (physical bits 0-2, 6-9, 21-24 at n = 15), 8 butterflies: the
        return n, [("H", [q], 0.0, 0.0, 0.0) for q in (6, 7, 8, 9)]  # memory pipeline with 128-byte runs 1 KiB apart
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
    n_s = n if args.cpu_full_size else min(n, args.cpu_sample_qubits)
    fam = name.rpartition("_n")[0]
    _, g_s = workload(f"{fam}_n{n_s}") if fam not in ("adder", "vqe_uccsd") else (n, gates)
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
    extrapolated = n_s != n or len(g_s) != len(gates)
    sample = (f"{fam}_n{n_s} ({len(g_s)} gates) via dmsim_cpu_omp n_cpus={n_cpus}: {ms_sample:.1f} ms/sim"
              + (f"; EXTRAPOLATED x4^{n - n_s} and x{len(gates)}/{len(g_s)} gates to {name} (the full size needs "
                 f"{8 * 8 * 4 ** n / 2 ** 30:.0f} GiB of host memory and minutes per sim)" if extrapolated else "; measured at full size"))
    line = {"metric": "gates/sec", "value": value, "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_full, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": name, "n_qubits": n, "n_gates": len(gates)},
            "cpu_baseline": {"value": value, "unit": "gates/s", "cores": n_cpus, "kind": kind, "sample": sample,
                             "same_config": not extrapolated, "extrapolated": extrapolated,
                             "measured_at_full_size": not extrapolated},
            "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if extrapolated:
        # the same reference backend measured ONCE at the full size on a box of this pool (committed line; 64 GiB of host
        # memory and minutes per run are outside the bounded sample every bench run takes)
        try:
            f = os.path.join(ROOT, "profiles", "r2_bench_reference_full_size.jsonl")
            d = json.loads(open(f).read().strip().splitlines()[-1])
            if d["config"]["workload"] == name:
                line["cpu_baseline"]["full_size_measurement"] = {"ms_per_step": d["ms_per_step"], "value": d["value"], "cores": d["cpu_baseline"]["cores"],
                                                                 "file": "profiles/r2_bench_reference_full_size.jsonl",
                                                                 "what": "bench.py --impl reference --cpu-full-size --steps 1, round 2, same pool"}
        except Exception:  # noqa: BLE001
            pass
    return line


class Job:
    """Process-group plumbing of one bench run (torch.distributed only for the barrier / max-over-ranks)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None

    def init(self):
        import torch
        self.torch = torch
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            dist.barrier()
            self.dist = dist
        else:
            torch.cuda.set_device(0)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max(self, values):
        if self.dist is None:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


JIT_MODE = 2  # --jit: 2 = wait for the run-time specialised sweep kernels (deterministic timing), 1 = tiered, 0 = interpreter kernels only


def measure(job, dm, name, steps, warmup, world=None, sampler=None, e2e=True, keep=False, first_call=False):
    """Device time of `steps` runs of workload `name` on the DENSE resident state (+ the end-to-end legs).
    world = None: the job's ranks (one shard per process); world = 1 inside a multi-rank job: this process alone."""
    import numpy as np
    L = dm.lib()
    n, gates = workload(name)
    rec, mats = dm.pack_gates(gates)
    solo = world == 1 and job.world > 1
    world = job.world if world is None else world
    sim = dm.Simulation(n, world, rank=job.rank, device=job.local_rank) if world > 1 else dm.Simulation(n, 1)
    barrier = (lambda: job.torch.cuda.synchronize()) if solo or world == 1 else job.barrier

    def set_circuit():
        dm._check(L.dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32))
        sim._uploaded = True

    # `value` / `roofline` are measured on the DENSE resident state: every tile of every sweep is launched (after a reset
    # the engine would otherwise skip the tiles that are still all-zero, see the e2e leg below)
    first_ms = None
    if first_call:
        # the very first call of this circuit in the process, as a user's one-shot run sees it (tiered execution: nothing
        # waits for the run-time compiler -- the interpreter kernels run while the specialised ones are being built)
        dm.set_option("jit", 1 if JIT_MODE else 0)
        dm.set_option("sparse", 1)
        d0 = np.empty(1 << n)
        barrier()
        t1 = time.perf_counter()
        sim.reset_dm()
        set_circuit()
        sim.run()
        dm._check(L.dmb_get_diag(sim._h, d0.ctypes.data))
        barrier()
        first_ms = (time.perf_counter() - t1) * 1e3
    # A step = the circuit applied to a freshly reset state with sparse start OFF (every tile of every sweep is processed: the
    # work of a dense state; FP64 timing does not depend on the values).  Starting every step from the same bit layout lets the
    # engine reuse the circuit's plan, its captured graph and its specialised kernels -- a circuit that CONTINUES from the
    # previous run's state starts from the layout that run left behind and is planned anew (`continued_state` below).
    dm.set_option("jit", JIT_MODE)
    dm.set_option("sparse", 0)
    sim.reset_dm()
    set_circuit()
    for _ in range(warmup):
        sim.reset_dm()
        sim.run()
    if sampler is not None:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, comm_ms, launches, sweeps = 0.0, 0.0, 0, 0
    for _ in range(steps):
        sim.reset_dm()
        sim.run()
        st = sim.last_stats
        dev_ms += st["sim_ms"]; comm_ms += st["comm_ms"]
        launches += st["n_launches"]; sweeps += st["n_sweeps"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler is not None else None
    st = dict(sim.last_stats)
    if not solo and world > 1:
        dev_ms, comm_ms, wall_ms = job.max([dev_ms, comm_ms, wall_ms])
    jit = {"mode": JIT_MODE, "sweeps_specialised": int(dm.query("jit_sweeps", sim._h)), "sweeps_pending": int(dm.query("jit_pending", sim._h)),
           "sweeps_per_step": int(st["n_sweeps"])}
    cont_ms = None
    if first_call:
        # the same circuit CONTINUING from the state (and bit layout) the previous run left: tiered execution, no waiting
        dm.set_option("jit", 1 if JIT_MODE else 0)
        cont = []
        for _ in range(2):
            sim.run()
            cont.append(sim.last_stats["sim_ms"])
        cont_ms = job.max([min(cont)])[0] if (not solo and world > 1) else min(cont)
        dm.set_option("jit", JIT_MODE)
    out = {"name": name, "n": n, "gates": gates, "n_gates": len(gates), "st": st, "ms_step": dev_ms / steps, "dev_ms": dev_ms, "jit": jit,
           "first_ms": first_ms, "cont_ms": cont_ms,
           "comm_ms": comm_ms, "wall_ms_step": wall_ms / steps, "launches": launches, "sweeps": sweeps, "clocks": clocks,
           "steps": steps, "world": world}
    if e2e:
        # ---- end to end through the C-ABI with HOST buffers: reset + circuit upload (H2D) + run + diagonal (D2H)
        # (plan cache OFF: every step plans the circuit and copies its device tables host -> device, like a first call)
        diag = np.empty(1 << n)
        dm.set_option("sparse", 1)
        dm.set_option("plan_cache", 0)
        e2e_ms, parts = [], [0.0, 0.0, 0.0, 0.0]
        for i in range(2 + min(steps, 3)):
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
        e2e_v = sum(e2e_ms) / len(e2e_ms)
        parts = [x / len(e2e_ms) for x in parts]
        # the same call sequence WITHOUT the reset: the circuit is re-uploaded and applied to the resident (dense) state,
        # so that no tile is skipped as still-zero
        dm.set_option("sparse", 0)
        dm.set_option("jit", 1 if JIT_MODE else 0)  # (a continued state is planned anew every time: tiered, never waiting)
        sim.reset_dm()
        res_ms = []
        for i in range(1 + min(steps, 3)):
            barrier()
            t1 = time.perf_counter()
            set_circuit()
            sim.run()
            dm._check(L.dmb_get_diag(sim._h, diag.ctypes.data))
            barrier()
            if i >= 1:
                res_ms.append((time.perf_counter() - t1) * 1e3)
        e2e_res = sum(res_ms) / len(res_ms)
        dm.set_option("jit", JIT_MODE)
        h2d = int(sim.last_stats["h2d_bytes"])
        # the same circuit set again with the plan cache ON (the engine's default): the host pipeline and the H2D of the
        # tables are skipped, the captured graph / parameter list is reused
        dm.set_option("plan_cache", 1)
        dm.set_option("sparse", 1)
        warm_ms = []
        for i in range(2 + min(steps, 3)):
            barrier()
            t1 = time.perf_counter()
            sim.reset_dm()
            set_circuit()
            sim.run()
            dm._check(L.dmb_get_diag(sim._h, diag.ctypes.data))
            barrier()
            if i >= 2:
                warm_ms.append((time.perf_counter() - t1) * 1e3)
        e2e_warm = sum(warm_ms) / len(warm_ms)
        dm.set_option("sparse", 0)
        if not solo and world > 1:
            e2e_v, e2e_res, e2e_warm = job.max([e2e_v, e2e_res, e2e_warm])
        out.update({"e2e_ms": e2e_v, "e2e_parts": parts, "e2e_res_ms": e2e_res, "e2e_warm_ms": e2e_warm, "trace": float(diag.sum()),
                    "h2d": h2d})
    if keep:
        out["sim"] = sim
    else:
        del sim
    return out


def roofline(m, peak, peak_src):
    """The dominant kernel (sweep_kernel) of a measured workload against BOTH floors; `bound` = the higher floor."""
    st = m["st"]
    comp_s = (m["dev_ms"] - m["comm_ms"]) * 1e-3
    sweep_bytes = st["sweep_bytes"]
    hbm = (m["sweeps"] * sweep_bytes) / comp_s / 1e9 if comp_s > 0 else 0.0
    fp64 = 2.0 * st["fp64_ops"] * m["steps"] / comp_s / 1e12 if comp_s > 0 else 0.0  # FMA-equivalent TFLOP/s
    frac_hbm, frac_fp64 = hbm / peak, fp64 / FP64_TFLOPS
    bound = "fp64" if frac_fp64 > frac_hbm else "hbm"
    specialised = m.get("jit", {}).get("sweeps_specialised", 0) > 0
    r = {"bound": bound, "kernel": "dmb_jit_sweep (run-time specialised sweep_kernel)" if specialised else "sweep_kernel",
         "achieved": fp64 if bound == "fp64" else hbm, "peak": FP64_TFLOPS if bound == "fp64" else peak,
         "unit": "TFLOP/s" if bound == "fp64" else "GB/s", "frac": max(frac_hbm, frac_fp64),
         "frac_hbm": frac_hbm, "achieved_GBps": hbm, "peak_GBps": peak, "frac_of_8TBs_nominal": hbm / 8000.0, "peak_source": peak_src,
         "frac_fp64": frac_fp64, "achieved_fp64_TFLOPs": fp64, "peak_fp64_TFLOPs": FP64_TFLOPS, "peak_fp64_source": FP64_SRC,
         "fp64_note": "FP64 pipe slots (DFMA / DMUL / DADD of the register-level ops, dmb_stats.fp64_ops) x 2 = FMA-equivalent flop",
         "avg_launch_ms": comp_s * 1e3 / max(1, m["sweeps"]), "bytes_per_launch": sweep_bytes,
         "fp64_instr_per_element_per_step": st["fp64_ops"] / (sweep_bytes / 32),
         "traffic": _ncu_traffic(m["name"])}
    return r


def brief(m, peak, peak_src):
    """Compact record of an extra workload."""
    r = roofline(m, peak, peak_src)
    d = {"workload": m["name"], "n_qubits": m["n"], "n_gates": m["n_gates"], "n_gpus": m["world"], "steps": m["steps"],
         "ms_per_step": m["ms_step"], "gates_per_s": m["n_gates"] / (m["ms_step"] * 1e-3), "sweeps_per_step": m["st"]["n_sweeps"],
         "gpu_launches_per_step": m["launches"] // max(1, m["steps"]), "jit": m["jit"],
         "roofline": {k: r[k] for k in ("bound", "frac", "frac_hbm", "frac_fp64", "achieved_GBps", "achieved_fp64_TFLOPs", "avg_launch_ms")}}
    if "e2e_ms" in m:
        d["e2e"] = {"ms_per_step": m["e2e_ms"], "gates_per_s": m["n_gates"] / (m["e2e_ms"] * 1e-3),
                    "resident_state_ms_per_step": m["e2e_res_ms"], "repeated_circuit_ms_per_step": m["e2e_warm_ms"], "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 8 * (1 << m["n"])}
        d["trace_after_run"] = m["trace"]
    return d


def parity_small(job, dm):
    """An 11-qubit circuit of every op (two runs: the second starts from the remapped layout) on the job's ranks against
    the oracle (the C restatement, bit-identical to the reference CPU backend): full matrix, diagonal, trace."""
    import numpy as np
    import oracle
    n = 11 if job.world <= 8 else 12
    _, gates = workload(f"allops_n{n}")
    sim = dm.Simulation(n, job.world, rank=job.rank, device=job.local_rank) if job.world > 1 else dm.Simulation(n, 1)
    o = oracle.Oracle(n) if job.rank == 0 else None
    rec, mats = dm.pack_gates(gates)
    err, exch = 0.0, 0
    for rep in range(2):
        dm._check(dm.lib().dmb_set_circuit(sim._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32))
        sim._uploaded = True
        sim.run()
        exch += sim.last_stats["n_exchanges"]
        re, im = sim.get_dm()      # collective for world > 1: every rank receives the full matrix
        d, tr = sim.diag(), sim.trace()
        if job.rank == 0:
            o.sim(gates)
            ore, oim = o.dm()
            err = max(err, float(np.abs(re - ore).max()), float(np.abs(im - oim).max()), float(np.abs(d - o.diag()).max()), abs(tr - 1.0))
    del sim
    return {"world": job.world, "n": n, "circuit": f"allops_n{n} (60 random gates over all 38 ops + C1/C2, run twice)",
            "against": "oracle/dmsim_oracle.c (bit-identical to the reference dmsim_cpu_omp, tests/test_oracle.py)",
            "exchanges": exch, "max_abs_err": err, "tol": PARITY_TOL}


def parity_workload(job, m):
    """The TIMED workload itself: one run from |0..0><0..0| against a 2^n state-vector run of the same gates."""
    from oracle.statevector import check_against_statevector
    sim = m["sim"]
    sim.reset_dm()
    sim.run()
    diag, purity = sim.diag(), sim.purity()
    res = check_against_statevector(m["n"], m["gates"], diag, sim.elements, purity)
    res.update({"workload": m["name"], "n": m["n"], "world": m["world"], "trace_err": abs(float(diag.sum()) - 1.0),
                "against": "oracle/statevector.py (2^n state vector of the same gates; pinned to the oracle at n <= 7): all 2^n "
                           "diagonal probabilities, 2^16 random elements, purity",
                "max_abs_err": max(res["diag"], res["elements"]), "tol": PARITY_TOL})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample-qubits", type=int, default=13)
    ap.add_argument("--cpu-full-size", action="store_true", help="reference arm: run the named size itself (n = 15: 64 GiB, minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads / strong_scaling / parity (quick A/B runs)")
    ap.add_argument("--no-first-call", action="store_true", help="skip the first_call / continued_state legs (profiling runs)")
    ap.add_argument("--jit", type=int, default=int(os.environ.get("DMB_JIT", "2")),
                    help="run-time specialised sweep kernels: 2 wait for the compiler (default), 1 tiered, 0 interpreter kernels only")
    args = ap.parse_args()
    global JIT_MODE
    JIT_MODE = args.jit

    job = Job()
    rank, world = job.rank, job.world
    default = args.workload is None
    if default:
        args.workload = {1: "qft_n15", 2: "random_c1c2_n16", 4: "random_c1c2_n16", 8: "random_c1c2_n17"}.get(args.gpus, "random_c1c2_n16")

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, args.workload)), flush=True)
        return

    import __graft_entry__ as ge
    if job.local_rank == 0:
        ge.build()
    job.init()
    if job.local_rank != 0:
        ge.build()  # (already built: only imports the package)
    dm = importlib.import_module("dm-sim_b200")
    peak, peak_src = _peaks()
    extras = default and not args.no_extra

    parity = {}
    if extras:
        parity["small"] = parity_small(job, dm)

    sampler = ClockSampler(job.local_rank) if rank == 0 else None
    m = measure(job, dm, args.workload, args.steps, args.warmup, sampler=sampler, keep=True, first_call=not args.no_first_call)
    if extras and world > 1:
        parity["workload"] = parity_workload(job, m)
    m.pop("sim", None)

    extra_workloads, strong, interpreted = [], None, None
    if extras and world == 1 and JIT_MODE:
        # the same workload on the ahead-of-time INTERPRETER kernels (what runs until the specialised kernels are compiled)
        JIT_MODE = 0
        try:
            mi = measure(job, dm, args.workload, 3, 1, e2e=False)
            interpreted = {"ms_per_step": mi["ms_step"], "roofline_frac": roofline(mi, peak, peak_src)["frac"]}
        finally:
            JIT_MODE = args.jit
    if extras and world == 1:
        # the other single-GPU configurations of BASELINE.json, each outside the timed region of the headline
        for name, k, w in (("bv_n15", 5, 3), ("adder_n10", 20, 3), ("vqe_uccsd_n8", 5, 3), ("random_c1c2_n15", 2, 1), ("random_c1c2_n16", 2, 1)):
            try:
                extra_workloads.append(brief(measure(job, dm, name, k, w), peak, peak_src))
            except Exception as e:  # noqa: BLE001  (e.g. not enough device memory for the 64 GiB state)
                extra_workloads.append({"workload": name, "failed": repr(e)})
    if extras and world > 1:
        # strong scaling on ONE workload: random_c1c2_n16 (64 GiB) on one GPU (rank 0 alone) vs the same circuit on N
        anchor = "random_c1c2_n16"
        mn = m if args.workload == anchor else measure(job, dm, anchor, 2, 1, e2e=False)
        job.barrier()
        m1 = measure(job, dm, anchor, 2, 1, world=1, e2e=False) if rank == 0 else None
        job.barrier()
        if rank == 0:
            strong = {"workload": anchor, "ms_per_step_1gpu": m1["ms_step"], "ms_per_step_Ngpu": mn["ms_step"], "n_gpus": world,
                      "speedup": m1["ms_step"] / mn["ms_step"], "efficiency": m1["ms_step"] / mn["ms_step"] / world,
                      "sweeps_1gpu": m1["st"]["n_sweeps"], "sweeps_Ngpu": mn["st"]["n_sweeps"], "exchanges_Ngpu": mn["st"]["n_exchanges"],
                      "comm_ms_Ngpu": mn["comm_ms"] / mn["steps"],
                      "what": "same circuit, same total work: rank 0 alone on one GPU (measured in this run) vs all N ranks"}

    if rank != 0:
        job.close()
        return
    st = m["st"]
    n, n_gates, ms_step = m["n"], m["n_gates"], m["ms_step"]
    line = {
        "metric": "gates/sec", "value": n_gates / (ms_step * 1e-3), "unit": "gates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_gate": ms_step / n_gates,
        "higher_is_better": True, "scaling": "strong" if world in (2, 4) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "n_qubits": n, "n_gates": n_gates, "n_primitives": st["n_primitives"],
                   "fused_blocks_per_side": st["n_blocks"], "sweeps_per_step": st["n_sweeps"],
                   "exchanges_per_step": st["n_exchanges"], "state_bytes": 16 * 4 ** n,
                   "state_bytes_per_gpu": 16 * 4 ** n // world,
                   "scaling_note": "BASELINE.json names a different configuration per GPU count (15 q on 1 GPU, 16 q on 2 / 4, 17 q "
                                   "on 8): `value` of different N is NOT one scaling series.  The same-workload series is "
                                   "`strong_scaling` (random_c1c2_n16 on 1 GPU, measured by rank 0 in this run, vs N GPUs)",
                   "step": "dmb_reset_dm + dmb_run with sparse start OFF (all tiles of all sweeps processed, the work of a dense state); "
                           "device time of dmb_run (CUDA events on the engine's stream)",
                   "l2_policy": "state per GPU (>= 16 GiB at the named sizes) is far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"shard on top {world.bit_length() - 1} index bits, one process per GPU" if world > 1 else "single GPU"},
        "roofline": roofline(m, peak, peak_src),
        "e2e": {"value": n_gates / (m["e2e_ms"] * 1e-3), "unit": "gates/s", "ms_per_step": m["e2e_ms"],
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 8 * (1 << n),
                "host_call_ms": dict(zip(("reset_dm", "set_circuit", "run", "get_diag"), m["e2e_parts"])),
                "first_call_ms": m["first_ms"],
                "quoted": "from |0..0><0..0| (dmb_reset_dm inside the timed region; on one GPU the leading sweeps skip the tiles "
                          "that are still all-zero, 'sparse start')",
                "resident_state": {"value": n_gates / (m["e2e_res_ms"] * 1e-3), "ms_per_step": m["e2e_res_ms"],
                                   "what": "same calls without dmb_reset_dm: circuit applied to the resident dense state"},
                "repeated_circuit": {"value": n_gates / (m["e2e_warm_ms"] * 1e-3), "ms_per_step": m["e2e_warm_ms"],
                                     "what": "same calls with the engine's plan cache on (default): a circuit that is set again reuses "
                                             "its plan and the device tables, h2d_bytes_per_step = 0; the headline e2e runs with the cache off"},
                "what": "host gate list in, host diagonal out, one circuit from |0><0| as the reference's sim(): dmb_reset_dm + "
                        "dmb_set_circuit (plan + H2D of the device op tables) + dmb_run + dmb_get_diag (D2H of the 2^n "
                        "probabilities), wall clock"},
        # size-independent rate (gates/s shrinks 4x per added qubit): gates x density-matrix elements updated per second
        "work_rate": {"value": n_gates * float(4 ** n) / (ms_step * 1e-3), "unit": "gate x element updates/s"},
        "jit": dict(m["jit"], compiled_kernels=int(dm.query("jit_compiled")), disk_cache_hits=int(dm.query("jit_disk_hits")),
                    failed=int(dm.query("jit_failed")), compile_ms_total=dm.query("jit_compile_ms"), interpreter_kernels=interpreted,
                    what="sweeps run on kernels specialised at run time (NVRTC, csrc/jit.cu: the sweep's program as straight-line "
                         "code over the same op bodies); compiled on worker threads, cached by program text in memory and on disk; "
                         "`value` waits for them (mode 2), e2e.first_call_ms does not (tiered: interpreter kernels meanwhile)"),
        "continued_state": {"ms_per_step": m["cont_ms"],
                            "what": "dmb_run again WITHOUT the reset (tiered): the circuit continues from the bit layout the previous run "
                                    "left, so it is planned anew and runs on the interpreter kernels until its own kernels are compiled"},
        "gpu_launches": int(m["launches"]), "wall_ms_per_step": m["wall_ms_step"],
        "trace_after_run": m["trace"], "clocks": m["clocks"],
    }
    if parity:
        worst = max(p["max_abs_err"] for p in parity.values())
        line["parity"] = dict(parity, max_abs_err=worst, tol=PARITY_TOL, ok=bool(worst < PARITY_TOL))
    if extra_workloads:
        line["extra_workloads"] = extra_workloads
    if strong:
        line["strong_scaling"] = strong
    if world > 1:
        line["comm"] = {"ms_per_step": m["comm_ms"] / args.steps, "bytes_sent_per_rank_per_step": st["exchange_bytes"],
                        "GBps_per_direction": (st["exchange_bytes"] / 1e9) / (m["comm_ms"] / args.steps * 1e-3) if m["comm_ms"] > 0 else None,
                        "what": "device time between the CUDA events around the qubit remap (barrier + permuting sweep storing into "
                                "the peers' shards over NVLink + barrier)"}
    if world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_arm(argparse.Namespace(gpus=1, steps=1, warmup=0, cpu_sample_qubits=args.cpu_sample_qubits,
                                                   cpu_full_size=False), args.workload)
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:  # the checker must never take the product's number down
            line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    job.close()
    if parity and not line["parity"]["ok"]:
        sys.exit(3)


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
