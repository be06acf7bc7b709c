"""oracle.statevector -- TEST INFRASTRUCTURE ONLY (checker for sizes the density-matrix oracle cannot reach).

A 2^n state-vector run of a circuit with the reference's gate conventions.  Every BASELINE.json configuration starts
from the pure state |0..0><0..0|, so rho = psi psi^dagger and what the engine stores (rho^T, [col][row]) is

    dm_real_res[col*dim + row] + i dm_imag_res[col*dim + row] = psi[row] * conj(psi[col])

which gives ALL 2^n diagonal probabilities, any sampled off-diagonal element and the purity (= 1) of a 16- or
17-qubit run for the price of a 2^n vector (SURVEY.md section 8c "Reach" iii).  This is a RESTATEMENT of the
reference's gate matrices (src/dmsim_nvgpu_omp.cuh:1004-1489, Appendix A.1 of SURVEY.md), not the reference itself:
it is pinned to the reference by tests/test_oracle.py (the density matrix it implies equals the oracle's at n <= 7).

Supported: H X Y Z S SDG T TDG U1 RZ RX RY U2 U3 W CX CZ and the raw C1 / C2 gates (matrix index of C2 =
2*bit(qubit1) + bit(qubit2), :1066-1069).
"""
from __future__ import annotations

import numpy as np

S2I = 0.70710678118654752440  # src/config.hpp:57


def _one(name, th, ph, la):
    c, s = np.cos, np.sin
    if name == "H":
        return np.array([[S2I, S2I], [S2I, -S2I]], dtype=np.complex128)
    if name == "X":
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if name == "Y":
        return np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    if name == "Z":
        return np.diag([1, -1]).astype(np.complex128)
    if name == "S":
        return np.diag([1, 1j]).astype(np.complex128)
    if name == "SDG":
        return np.diag([1, -1j]).astype(np.complex128)
    if name == "T":
        return np.diag([1, S2I * (1 + 1j)]).astype(np.complex128)
    if name == "TDG":
        return np.diag([1, S2I * (1 - 1j)]).astype(np.complex128)
    if name == "U1":
        return np.diag([1, c(la) + 1j * s(la)]).astype(np.complex128)
    if name == "RZ":  # == U1(phi), :1485-1489
        return np.diag([1, c(ph) + 1j * s(ph)]).astype(np.complex128)
    if name == "RX":
        return np.array([[c(th / 2), -1j * s(th / 2)], [-1j * s(th / 2), c(th / 2)]], dtype=np.complex128)
    if name == "RY":
        return np.array([[c(th / 2), -s(th / 2)], [s(th / 2), c(th / 2)]], dtype=np.complex128)
    if name == "W":
        return S2I * np.array([[1, -1j], [-1j, 1]], dtype=np.complex128)
    if name == "U2":
        return S2I * np.array([[1, -(c(la) + 1j * s(la))], [c(ph) + 1j * s(ph), c(ph + la) + 1j * s(ph + la)]], dtype=np.complex128)
    if name == "U3":
        return np.array([[c(th / 2), -(c(la) + 1j * s(la)) * s(th / 2)],
                         [(c(ph) + 1j * s(ph)) * s(th / 2), (c(ph + la) + 1j * s(ph + la)) * c(th / 2)]], dtype=np.complex128)
    return None


def apply_1q(psi, n, q, m):
    v = psi.reshape(1 << (n - 1 - q), 2, 1 << q)
    return np.einsum("ab,xby->xay", m, v).reshape(-1)


def apply_2q(psi, n, q1, q2, m):
    """4x4 matrix with index 2*bit(q1) + bit(q2)."""
    hi, lo = max(q1, q2), min(q1, q2)
    v = psi.reshape(1 << (n - 1 - hi), 2, 1 << (hi - lo - 1), 2, 1 << lo)
    m4 = np.asarray(m, dtype=np.complex128).reshape(2, 2, 2, 2)  # [r1, r2, c1, c2] in (q1, q2) order
    if q1 > q2:
        out = np.einsum("abcd,xcydz->xaybz", m4, v)
    else:
        out = np.einsum("abcd,xdycz->xbyaz", m4, v)
    return out.reshape(-1)


def statevector(n, gates):
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    idx = np.arange(1 << n)
    for g in gates:
        name, q = g[0], g[1]
        th, ph, la = (float(x) for x in g[2:5])
        if name == "C1":
            psi = apply_1q(psi, n, q[0], np.asarray(g[5], dtype=np.complex128).reshape(2, 2))
        elif name == "C2":
            psi = apply_2q(psi, n, q[0], q[1], g[5])
        elif name == "CX":
            c, t = q[0], q[1]
            psi = psi[np.where((idx >> c) & 1 == 1, idx ^ (1 << t), idx)]
        elif name == "CZ":
            psi = np.where(((idx >> q[0]) & 1 == 1) & ((idx >> q[1]) & 1 == 1), -psi, psi)
        elif name == "ID":
            pass
        else:
            m = _one(name, th, ph, la)
            if m is None:
                raise ValueError(f"statevector(): unsupported gate {name}")
            psi = apply_1q(psi, n, q[0], m)
    return psi


def check_against_statevector(n, gates, diag, elements_fn, purity, rng_seed=15, n_probe=1 << 16):
    """max |error| of (all 2^n diagonal probabilities, n_probe random elements, purity - 1) of an engine result
    against the state-vector run.  elements_fn(flat_index uint64[]) -> complex[] (dm_real_res + i dm_imag_res)."""
    psi = statevector(n, gates)
    rng = np.random.default_rng(rng_seed)
    col = rng.integers(0, 1 << n, size=n_probe, dtype=np.uint64)
    row = rng.integers(0, 1 << n, size=n_probe, dtype=np.uint64)
    got = elements_fn((col << np.uint64(n)) | row)
    want = psi[row.astype(np.int64)] * np.conj(psi[col.astype(np.int64)])  # res[col][row] = rho[row][col]
    return {"diag": float(np.abs(np.asarray(diag) - np.abs(psi) ** 2).max()), "elements": float(np.abs(got - want).max()),
            "purity": float(abs(purity - 1.0)), "n_probe": int(n_probe)}
