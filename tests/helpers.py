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
    """TEST INFRASTRUCTURE: 2^n state-vector run with the reference's gate conventions (oracle/statevector.py: H X Y Z S
    SDG T TDG U1 RZ RX RY U2 U3 W CX CZ and raw C1 / C2).  Used for size-independent checks at n >= 15: every circuit
    of BASELINE.json starts pure, so rho = psi psi^dagger and what the engine stores is rho^T."""
    from oracle.statevector import statevector as sv
    return sv(n, gates)
