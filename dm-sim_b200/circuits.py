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
