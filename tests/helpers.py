"""Shared test helpers (test infrastructure)."""
import numpy as np

ARITY = {"CX": 2, "CZ": 2, "CY": 2, "SWAP": 2, "CH": 2, "CRX": 2, "CRY": 2, "CRZ": 2, "CU1": 2, "CU3": 2, "RXX": 2,
         "RZZ": 2, "RYY": 2, "CCX": 3, "CSWAP": 3, "RCCX": 3, "RC3X": 4, "C3X": 4, "C3SQRTX": 4, "C4X": 5, "C2": 2}

OP_NAMES = [
    "U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S",
    "SDG", "T", "TDG", "RX", "RY", "RZ", "CZ", "CY", "SWAP", "CH",
    "CCX", "CSWAP", "CRX", "CRY", "CRZ", "CU1", "CU3", "RXX", "RZZ", "RCCX",
    "RC3X", "C3X", "C3SQRTX", "C4X", "R", "SRN", "W", "RYY",
]


def haar(d, rng):
    z = (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / np.sqrt(2.0)
    q, r = np.linalg.qr(z)
    return q * (np.diagonal(r) / np.abs(np.diagonal(r)))


def random_gates(n, m, rng, names=None, with_raw=True, exclude=("SRN",)):
    """m random gates over the full op set (R gets |theta| = 1 so that states stay normalised)."""
    pool = [x for x in (names or OP_NAMES) if x not in exclude]
    if with_raw and names is None:
        pool = pool + ["C1", "C2"]
    out = []
    while len(out) < m:
        nm = pool[rng.integers(len(pool))]
        a = ARITY.get(nm, 1)
        if a > n:
            continue
        q = [int(x) for x in rng.choice(n, size=a, replace=False)]
        th, ph, la = (float(x) for x in rng.uniform(-3.2, 3.2, size=3))
        if nm == "R":
            th = 1.0 if rng.integers(2) else -1.0
        if nm == "C1":
            out.append((nm, q, 0.0, 0.0, 0.0, haar(2, rng)))
        elif nm == "C2":
            out.append((nm, q, 0.0, 0.0, 0.0, haar(4, rng)))
        else:
            out.append((nm, q, th, ph, la))
    return out


def each_op_once(n, rng):
    """Every op of enum OP (except SRN) once, on random distinct qubits."""
    out = []
    for nm in OP_NAMES + ["C1", "C2"]:
        if nm == "SRN":
            continue
        out += random_gates(n, 1, rng, names=[nm], with_raw=False)
    return out


def to_complex(re, im):
    return (np.asarray(re) + 1j * np.asarray(im)).reshape(-1)


def statevector(n, gates):
    """TEST INFRASTRUCTURE: 2^n state-vector run of a circuit made of H / X / U1 / RZ / CX only (the gates of
    benchmark/qft_n15.qasm and bv_n15.qasm), with the reference's conventions (U1 = diag(1, e^{i lam}), RZ == U1,
    src/dmsim_nvgpu_omp.cuh:1381-1397, :1485-1489).  Used for size-independent checks at n = 15: every circuit of
    BASELINE.json starts pure, so rho = psi psi^dagger and what the engine stores is rho^T."""
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    idx = np.arange(1 << n)
    s2i = 0.70710678118654752440
    for g in gates:
        name, q = g[0], g[1]
        if name == "H":
            b = 1 << q[0]
            lo = idx[(idx & b) == 0]
            a0, a1 = psi[lo].copy(), psi[lo | b].copy()
            psi[lo], psi[lo | b] = s2i * (a0 + a1), s2i * (a0 - a1)
        elif name == "X":
            psi = psi[idx ^ (1 << q[0])]
        elif name in ("U1", "RZ"):
            lam = g[4] if name == "U1" else g[3]
            psi[(idx >> q[0]) & 1 == 1] *= np.cos(lam) + 1j * np.sin(lam)
        elif name == "CX":
            c, t = q[0], q[1]
            psi = psi[np.where((idx >> c) & 1 == 1, idx ^ (1 << t), idx)]
        else:
            raise ValueError(f"statevector(): unsupported gate {name}")
    return psi
