#!/usr/bin/env python
"""Generates tests/golden/*.npz|txt by RUNNING THE REFERENCE ITSELF (oracle/_ref = the unmodified
src/dmsim_cpu_omp.hpp compiled in place) in the build container, where /root/reference exists.  The GPU box has no
reference tree: its tests read these committed fixtures.  Re-run:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import each_op_once, random_gates  # noqa: E402

dm = importlib.import_module("dm-sim_b200")
circuits = importlib.import_module("dm-sim_b200.circuits")
qasm = importlib.import_module("dm-sim_b200.qasm")


def save(name, n, gates, full=True, **extra):
    rec, mats = dm.pack_gates(gates)
    r = oracle.reference_run(n, gates, n_cpus=min(8, 1 << n), want_dm=full)
    out = dict(n=n, gates=rec, mats=mats, diag=r["diag"], **extra)
    if full:
        out["real"], out["imag"] = r["real"], r["imag"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n =", n, "gates =", len(gates), "trace =", r["diag"].sum())


def main():
    assert oracle.have_reference(), "needs /root/reference (build container)"
    rng = np.random.default_rng(20201115)
    # every op of enum OP (+ raw C1/C2) once after a scrambling prefix, full 32x32 matrix
    prefix = random_gates(5, 10, rng, names=["U3", "CX", "H", "T"], with_raw=False)
    save("all_ops_n5", 5, prefix + each_op_once(5, rng))
    save("random_mix_n6", 6, random_gates(6, 80, rng))
    save("random_c1c2_n6", 6, circuits.random_c1c2(6, 64, seed=7))
    save("srn_n4", 4, [("H", [0], 0, 0, 0), ("U3", [1], 0.3, 0.2, 0.1), ("SRN", [1], 0, 0, 0), ("CX", [1, 2], 0, 0, 0),
                       ("SRN", [0], 0, 0, 0), ("U3", [0], 0.3, 0.2, 0.1), ("CX", [0, 3], 0, 0, 0)])
    save("adder_n10", 10, circuits.adder_n10(), full=False)
    save("qft_n10", 10, circuits.qft(10), full=False)
    n, g = qasm.load_file("/root/reference/benchmark/vqe_uccsd_n8.qasm")
    save("vqe_uccsd_n8", n, g, full=False)
    with open(os.path.join(HERE, "factory_dump.txt"), "w") as f:
        f.write(oracle.reference_factory_dump())
    print("factory_dump.txt written")


if __name__ == "__main__":
    main()
