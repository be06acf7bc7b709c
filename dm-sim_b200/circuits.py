"""Workload generators for the configurations BASELINE.json names (gate lists as (op, qubits, theta, phi, lam
[, matrix]) tuples, the form ``pack_gates`` takes).  They reproduce the gate sequences of the reference's
benchmark files so that the GPU box (which has no /root/reference) can run them; tests/test_circuits.py checks
them gate-for-gate against the parsed .qasm files whenever the reference tree is present."""
from __future__ import annotations

import math

import numpy as np


def _r15(x: float) -> float:
    """benchmark/qft_n15.qasm prints its angles with 15 significant digits."""
    return float("%.15g" % x)


def qft(n: int):
    """benchmark/qft_n15.qasm layout: for j: [cu1(pi/2^(j-i)) j->i expanded to u1,cx,u1,cx,u1 for i<j]; h j."""
    g = []
    for j in range(n):
        for i in range(j):
            lam = _r15(math.pi / (1 << (j - i)) / 2.0)
            g.append(("U1", [j], 0.0, 0.0, lam))
            g.append(("CX", [j, i], 0.0, 0.0, 0.0))
            g.append(("U1", [i], 0.0, 0.0, -lam))
            g.append(("CX", [j, i], 0.0, 0.0, 0.0))
            g.append(("U1", [i], 0.0, 0.0, lam))
        g.append(("H", [j], 0.0, 0.0, 0.0))
    return g


def bv(n: int):
    """benchmark/bv_n15.qasm: Bernstein-Vazirani, hidden string all ones, ancilla = qubit n-1."""
    g = [("H", [q], 0.0, 0.0, 0.0) for q in range(n - 1)]
    g.append(("X", [n - 1], 0.0, 0.0, 0.0))
    g.append(("H", [n - 1], 0.0, 0.0, 0.0))
    g += [("CX", [q, n - 1], 0.0, 0.0, 0.0) for q in range(n - 1)]
    g += [("H", [q], 0.0, 0.0, 0.0) for q in range(n - 1)]
    return g


def adder_n10():
    """example/adder_n10_cpu_omp.cpp:46-61 (Cuccaro adder, 30 Gate objects = 142 primitives)."""
    g = [("X", [q], 0.0, 0.0, 0.0) for q in (1, 5, 6, 7, 8)]

    def maj(a, b, c):
        g.append(("CX", [c, b], 0.0, 0.0, 0.0)); g.append(("CX", [c, a], 0.0, 0.0, 0.0))
        g.append(("CCX", [a, b, c], 0.0, 0.0, 0.0))

    def unmaj(a, b, c):
        g.append(("CCX", [a, b, c], 0.0, 0.0, 0.0)); g.append(("CX", [c, a], 0.0, 0.0, 0.0))
        g.append(("CX", [a, b], 0.0, 0.0, 0.0))

    maj(0, 5, 1); maj(1, 6, 2); maj(2, 7, 3); maj(3, 8, 4)
    g.append(("CX", [4, 9], 0.0, 0.0, 0.0))
    unmaj(3, 8, 4); unmaj(2, 7, 3); unmaj(1, 6, 2); unmaj(0, 5, 1)
    return g


def haar_unitary(d: int, rng) -> np.ndarray:
    """QR of a complex Ginibre matrix with the phase fix (Mezzadri) -- Haar measure on U(d)."""
    z = (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / math.sqrt(2.0)
    q, r = np.linalg.qr(z)
    ph = np.diagonal(r) / np.abs(np.diagonal(r))
    return q * ph


def random_c1c2(n: int, n_gates: int = 256, seed: int = 20201115):
    """Synthetic config of BASELINE.json: alternating Haar-random C1 / C2 gates on uniform random qubits."""
    rng = np.random.default_rng(seed)
    g = []
    for i in range(n_gates):
        if i % 2 == 0 or n < 2:
            g.append(("C1", [int(rng.integers(n))], 0.0, 0.0, 0.0, haar_unitary(2, rng)))
        else:
            a, b = rng.choice(n, size=2, replace=False)
            g.append(("C2", [int(a), int(b)], 0.0, 0.0, 0.0, haar_unitary(4, rng)))
    return g


def vqe_uccsd_n8():
    """benchmark/vqe_uccsd_n8.qasm gate for gate (8 qubits, 10808 gates = 5488 cx + 2352 h + 2352 y + 616 rz): the gate
    list as parsed by the OpenQASM front-end (tests/golden/make_golden.py), stored next to this file so that the GPU box
    (which has no /root/reference) can run the configuration."""
    import os

    names = ["U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S", "SDG", "T", "TDG", "RX", "RY", "RZ"]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "vqe_uccsd_n8_gates.npz"))
    return [(names[int(op)], [int(a), int(b)], float(t), float(p), float(la))
            for op, a, b, t, p, la in zip(z["op"], z["q0"], z["q1"], z["theta"], z["phi"], z["lam"])]


_ARITY = {"CX": 2, "CZ": 2, "CY": 2, "SWAP": 2, "CH": 2, "CRX": 2, "CRY": 2, "CRZ": 2, "CU1": 2, "CU3": 2, "RXX": 2, "RZZ": 2,
          "RYY": 2, "CCX": 3, "CSWAP": 3, "RCCX": 3, "RC3X": 4, "C3X": 4, "C3SQRTX": 4, "C4X": 5, "C2": 2}
_ALL_OPS = ["U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S", "SDG", "T", "TDG", "RX", "RY", "RZ", "CZ", "CY", "SWAP", "CH",
            "CCX", "CSWAP", "CRX", "CRY", "CRZ", "CU1", "CU3", "RXX", "RZZ", "RCCX", "RC3X", "C3X", "C3SQRTX", "C4X", "R", "W", "RYY",
            "C1", "C2"]


def random_allops(n: int, n_gates: int = 60, seed: int = 1115):
    """n_gates random gates over every op of enum OP (except SRN, which is not a quantum gate) plus the raw C1 / C2, on
    random distinct qubits with random angles: the parity circuit bench.py runs on the ranks of a multi-GPU job."""
    rng = np.random.default_rng(seed)
    g = []
    while len(g) < n_gates:
        nm = _ALL_OPS[int(rng.integers(len(_ALL_OPS)))]
        a = _ARITY.get(nm, 1)
        if a > n:
            continue
        q = [int(x) for x in rng.choice(n, size=a, replace=False)]
        th, ph, la = (float(x) for x in rng.uniform(-3.2, 3.2, size=3))
        if nm == "R":
            th = 1.0 if rng.integers(2) else -1.0  # R multiplies by i*theta: |theta| = 1 keeps the state normalised
        if nm == "C1":
            g.append((nm, q, 0.0, 0.0, 0.0, haar_unitary(2, rng)))
        elif nm == "C2":
            g.append((nm, q, 0.0, 0.0, 0.0, haar_unitary(4, rng)))
        else:
            g.append((nm, q, th, ph, la))
    return g
