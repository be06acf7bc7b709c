"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the density-matrix gate-application path of pnnl/DM-Sim:

* ``Oracle``      : ctypes binding of ``oracle/_build/liboracle.so`` (``dmsim_oracle.c``), the plain-C
                    restatement of the reference algorithm (src/dmsim_nvgpu_omp.cuh:816-2008).
* ``reference_run``: ctypes binding of ``oracle/_ref/libdmsim_ref.so`` -- the reference CPU backend
                    itself (src/dmsim_cpu_omp.hpp), compiled in place by ``oracle/Makefile``.
* ``superop_numpy``: an independent numpy statement of the same map as a 2n-qubit superoperator
                    (U on bit q, conj(U) on bit q+n), used to cross-check the
                    restatement at small n.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker.  The product (``dm-sim_b200/``) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libdmsim_ref.so")

# enum OP order, src/dmsim_nvgpu_omp.cuh:42-48
OP_NAMES = [
    "U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S",
    "SDG", "T", "TDG", "RX", "RY", "RZ", "CZ", "CY", "SWAP", "CH",
    "CCX", "CSWAP", "CRX", "CRY", "CRZ", "CU1", "CU3", "RXX", "RZZ", "RCCX",
    "RC3X", "C3X", "C3SQRTX", "C4X", "R", "SRN", "W", "RYY",
]
OP = {name: i for i, name in enumerate(OP_NAMES)}
OP_RAW_C1 = 100
OP_RAW_C2 = 101

# POD gate record shared by liboracle.so, libdmsim_ref.so and the product's C-ABI (include/dmsim_b200.h)
GATE_DTYPE = np.dtype(
    [("op", "<i4"), ("qb", "<i4", (5,)), ("theta", "<f8"), ("phi", "<f8"), ("lam", "<f8"), ("mat", "<i8")],
    align=True,
)
assert GATE_DTYPE.itemsize == 56


def build(force: bool = False) -> None:
    """Compile the checkers (``make -C oracle``).  Building the checker is not using it."""
    need = force or not os.path.exists(_ORACLE_SO)
    have_ref_tree = os.path.exists("/root/reference/src/dmsim_cpu_omp.hpp")
    if have_ref_tree and not os.path.exists(_REF_SO):
        need = True
    if need:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


def pack_gates(gates, mats=None):
    """gates: iterable of (op, qubits, theta, phi, lam[, mat]) -> (structured array, mats array).

    ``op`` is a name from OP_NAMES / an int / "C1" / "C2"; raw C1/C2 gates carry a 2x2 / 4x4 complex
    matrix in ``mat`` which is appended to the returned matrix table (32 doubles per slot)."""
    rec = np.zeros(len(gates), dtype=GATE_DTYPE)
    table = [] if mats is None else list(mats)
    for i, g in enumerate(gates):
        op, qubits, theta, phi, lam = g[:5]
        if isinstance(op, str):
            op = {"C1": OP_RAW_C1, "C2": OP_RAW_C2}.get(op, OP.get(op))
        rec[i]["op"] = op
        qb = list(qubits) + [0] * (5 - len(qubits))
        rec[i]["qb"] = qb
        rec[i]["theta"], rec[i]["phi"], rec[i]["lam"] = theta, phi, lam
        if op >= 100:
            m = np.asarray(g[5], dtype=np.complex128)
            slot = np.zeros(32)
            flat = m.reshape(-1)
            slot[0:2 * flat.size:2] = flat.real
            slot[1:2 * flat.size:2] = flat.imag
            rec[i]["mat"] = len(table)
            table.append(slot)
    mats_arr = np.ascontiguousarray(np.array(table, dtype=np.float64).reshape(-1)) if table else np.zeros(32)
    return rec, mats_arr


class Oracle:
    """The C restatement: create -> sim(gates) [-> sim(gates) ...] -> dm / diag / measure."""

    def __init__(self, n_qubits: int):
        build()
        self.lib = ctypes.CDLL(_ORACLE_SO)
        L = self.lib
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_reset.argtypes = [ctypes.c_void_p]
        L.orc_sim.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.orc_sim.restype = ctypes.c_int
        L.orc_get_dm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_get_diag.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_set_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_measure.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
        L.orc_measure.restype = ctypes.c_double
        L.orc_sample_with_r.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
        self.n = n_qubits
        self.dim = 1 << n_qubits
        self.h = L.orc_create(n_qubits)
        if not self.h:
            raise MemoryError("oracle allocation failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def reset(self):
        self.lib.orc_reset(self.h)

    def sim(self, gates, mats=None):
        rec, table = gates if isinstance(gates, tuple) else pack_gates(gates, mats)
        rc = self.lib.orc_sim(self.h, rec.ctypes.data, len(rec), table.ctypes.data)
        if rc:
            raise ValueError("oracle: bad gate")
        return self

    def dm(self):
        """(real, imag) split arrays, each (dim, dim) indexed [col][row] -- what dm_real_res holds (rho^T)."""
        re = np.empty((self.dim, self.dim))
        im = np.empty((self.dim, self.dim))
        self.lib.orc_get_dm(self.h, re.ctypes.data, im.ctypes.data)
        return re, im

    def diag(self):
        d = np.empty(self.dim)
        self.lib.orc_get_diag(self.h, d.ctypes.data)
        return d

    def measure(self, repetition=10, seed=0):
        out = np.zeros(repetition, dtype=np.uint64)
        total = self.lib.orc_measure(self.h, seed, repetition, out.ctypes.data)
        return out, total

    def sample_with_r(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.zeros(len(r), dtype=np.uint64)
        self.lib.orc_sample_with_r(self.h, r.ctypes.data, len(r), out.ctypes.data)
        return out


_ref_lib = None


def _ref():
    global _ref_lib
    if _ref_lib is None:
        build()
        if not os.path.exists(_REF_SO):
            raise FileNotFoundError("oracle/_ref/libdmsim_ref.so not built (reference tree absent)")
        L = ctypes.CDLL(_REF_SO)
        L.ref_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ref_run.restype = ctypes.c_int
        L.ref_factory_dump.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.ref_factory_dump.restype = ctypes.c_int
        L.ref_measure_adder_smoke.argtypes = [ctypes.c_void_p]
        _ref_lib = L
    return _ref_lib


def reference_run(n_qubits, gates, mats=None, n_cpus=1, want_dm=True):
    """ONE sim() of the reference CPU backend from the reset state.

    Returns dict(real, imag [(dim,dim), [col][row]], diag, sim_ms [backend's own figure], wall_ms)."""
    L = _ref()
    rec, table = gates if isinstance(gates, tuple) else pack_gates(gates, mats)
    dim = 1 << n_qubits
    re = np.empty((dim, dim)) if want_dm else None
    im = np.empty((dim, dim)) if want_dm else None
    diag = np.empty(dim)
    times = np.zeros(2)
    rc = L.ref_run(n_qubits, n_cpus, rec.ctypes.data, len(rec), table.ctypes.data,
                   re.ctypes.data if want_dm else None, im.ctypes.data if want_dm else None,
                   diag.ctypes.data, times.ctypes.data)
    if rc:
        raise RuntimeError("reference run failed")
    return {"real": re, "imag": im, "diag": diag, "sim_ms": float(times[0]), "wall_ms": float(times[1])}


def reference_factory_dump() -> str:
    buf = ctypes.create_string_buffer(1 << 16)
    n = _ref().ref_factory_dump(buf, len(buf))
    if n < 0:
        raise RuntimeError("dump buffer too small")
    return buf.value.decode()


def reference_adder_smoke():
    out = np.zeros(5, dtype=np.uint64)
    _ref().ref_measure_adder_smoke(out.ctypes.data)
    return out


# ----------------------------------------------------------------------------------------------
# Independent numpy statement (small n only): primitive matrices of Appendix A.1 applied as
# E on bit q, conj(E) on bit q+n of the flat 2n-bit vector (probe C.4).
# Valid for complex-linear gates on a Hermitian input (everything except SRN).
# ----------------------------------------------------------------------------------------------
_S2I = 0.70710678118654752440
_PI = 3.14159265358979323846


def _u3(t, p, l):
    return np.array([[np.cos(t / 2), -np.exp(1j * l) * np.sin(t / 2)],
                     [np.exp(1j * p) * np.sin(t / 2), np.exp(1j * (p + l)) * np.cos(t / 2)]])


def _u2(p, l):
    return _S2I * np.array([[1, -np.exp(1j * l)], [np.exp(1j * p), np.exp(1j * (p + l))]])


def _u1(l):
    return np.array([[1, 0], [0, np.exp(1j * l)]])


def _rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]])


_H = _S2I * np.array([[1, 1], [1, -1]], dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_T = np.array([[1, 0], [0, _S2I * (1 + 1j)]])
_TDG = _T.conj()
_S = np.array([[1, 0], [0, 1j]])
_SDG = _S.conj()


def primitives(op, qb, theta, phi, lam, mat=None):
    """Expand one Gate into [(matrix, qubits)] primitives in application order (Appendix A.1 / A.3).
    2-qubit matrices use index 2*bit(qubits[0]) + bit(qubits[1])."""
    CXm = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    q = list(qb)
    P = []

    def one(m, a): P.append((np.asarray(m, dtype=complex), (a,)))
    def cx(a, b): P.append((CXm, (a, b)))
    def u1(l, a): one(_u1(l), a)
    def u2(p, l, a): one(_u2(p, l), a)
    def u3(t, p, l, a): one(_u3(t, p, l), a)
    def h(a): one(_H, a)
    def cu1(l, a, b): u1(l / 2, a); cx(a, b); u1(-l / 2, b); cx(a, b); u1(l / 2, b)
    def ccx(a, b, c):
        h(c); cx(b, c); one(_TDG, c); cx(a, c); one(_T, c); cx(b, c); one(_TDG, c); cx(a, c)
        one(_T, b); one(_T, c); h(c); cx(a, b); one(_T, a); one(_TDG, b); cx(a, b)
    def c3x_like(ang, a, b, c, d):
        seq = [(-ang, a), ("cx", a, b), (ang, b), ("cx", a, b), (-ang, b), ("cx", b, c), (ang, c),
               ("cx", a, c), (-ang, c), ("cx", b, c), (ang, c), ("cx", a, c), (-ang, c)]
        for s in seq:
            if s[0] == "cx":
                cx(s[1], s[2])
            else:
                h(d); cu1(s[0], s[1], d); h(d)

    name = op if isinstance(op, str) else ({OP_RAW_C1: "C1", OP_RAW_C2: "C2"}.get(op) or OP_NAMES[op])
    if name == "U3": u3(theta, phi, lam, q[0])
    elif name == "U2": u2(phi, lam, q[0])
    elif name == "U1": u1(lam, q[0])
    elif name == "CX": cx(q[0], q[1])
    elif name == "ID": pass
    elif name == "X": one(_X, q[0])
    elif name == "Y": one([[0, -1j], [1j, 0]], q[0])
    elif name == "Z": one([[1, 0], [0, -1]], q[0])
    elif name == "H": h(q[0])
    elif name == "S": one(_S, q[0])
    elif name == "SDG": one(_SDG, q[0])
    elif name == "T": one(_T, q[0])
    elif name == "TDG": one(_TDG, q[0])
    elif name == "RX": one(_rx(theta), q[0])
    elif name == "RY":
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        one([[c, -s], [s, c]], q[0])
    elif name == "RZ": u1(phi, q[0])
    elif name == "CZ": h(q[1]); cx(q[0], q[1]); h(q[1])
    elif name == "CY": one(_SDG, q[1]); cx(q[0], q[1]); one(_S, q[1])
    elif name == "SWAP": cx(q[0], q[1]); cx(q[1], q[0]); cx(q[0], q[1])
    elif name == "CH":
        a, b = q[0], q[1]
        h(b); one(_SDG, b); cx(a, b); h(b); one(_T, b); cx(a, b); one(_T, b); h(b); one(_S, b)
        one(_X, b); one(_S, a)
    elif name == "CCX": ccx(q[0], q[1], q[2])
    elif name == "CSWAP": cx(q[2], q[1]); ccx(q[0], q[1], q[2]); cx(q[2], q[1])
    elif name == "CRX":
        a, b = q[0], q[1]
        u1(_PI / 2, b); cx(a, b); u3(-lam / 2, 0, 0, b); cx(a, b); u3(lam / 2, -_PI / 2, 0, b)
    elif name == "CRY":
        a, b = q[0], q[1]
        u3(lam / 2, 0, 0, b); cx(a, b); u3(-lam / 2, 0, 0, b); cx(a, b)
    elif name == "CRZ":
        a, b = q[0], q[1]
        u1(lam / 2, b); cx(a, b); u1(-lam / 2, b); cx(a, b)
    elif name == "CU1": cu1(lam, q[0], q[1])
    elif name == "CU3":
        c, t = q[0], q[1]
        t1, t2, t3 = (lam - phi) / 2, theta / 2, -(phi + lam) / 2
        u1(-t3, c); u1(t1, t); cx(c, t); u3(-t2, 0, t3, t); cx(c, t); u3(t2, phi, 0, t)
    elif name == "RXX":
        a, b = q[0], q[1]
        u3(_PI / 2, theta, 0, a); h(b); cx(a, b); u1(-theta, b); cx(a, b); h(b); u2(-_PI, _PI - theta, a)
    elif name == "RZZ":
        a, b = q[0], q[1]
        cx(a, b); u1(theta, b); cx(a, b)
    elif name == "RCCX":
        a, b, c = q[0], q[1], q[2]
        u2(0, _PI, c); u1(_PI / 4, c); cx(b, c); u1(-_PI / 4, c); cx(a, c); u1(_PI / 4, c); cx(b, c)
        u1(-_PI / 4, c); u2(0, _PI, c)
    elif name == "RC3X":
        a, b, c, d = q[0], q[1], q[2], q[3]
        u2(0, _PI, d); u1(_PI / 4, d); cx(c, d); u1(-_PI / 4, d); u2(0, _PI, d); cx(a, d); u1(_PI / 4, d)
        cx(b, d); u1(-_PI / 4, d); cx(a, d); u1(_PI / 4, d); cx(b, d); u1(-_PI / 4, d); u2(0, _PI, d)
        u1(_PI / 4, d); cx(c, d); u1(-_PI / 4, d); u2(0, _PI, d)
    elif name == "C3X": c3x_like(_PI / 4, *q[:4])
    elif name == "C3SQRTX": c3x_like(_PI / 8, *q[:4])
    elif name == "C4X":
        a, b, c, d, e = q
        h(e); cu1(-_PI / 2, d, e); h(e); c3x_like(_PI / 4, a, b, c, d)
        h(d); cu1(_PI / 4, d, e); h(d); c3x_like(_PI / 4, a, b, c, d); c3x_like(_PI / 8, a, b, c, e)
    elif name == "R": one([[1, 0], [0, 1j * theta]], q[0])
    elif name == "W": one(_S2I * np.array([[1, -1j], [-1j, 1]]), q[0])
    elif name == "RYY":
        a, b = q[0], q[1]
        one(_rx(_PI / 2), a); one(_rx(_PI / 2), b); cx(a, b); u1(theta, b); cx(a, b)
        one(_rx(-_PI / 2), a); one(_rx(-_PI / 2), b)
    elif name == "C1": one(np.asarray(mat).reshape(2, 2), q[0])
    elif name == "C2": P.append((np.asarray(mat, dtype=complex).reshape(4, 4), (q[0], q[1])))
    elif name == "SRN":
        raise ValueError("SRN is not complex-linear; no superoperator form")
    else:
        raise ValueError(name)
    return P


def _apply(vec, nbits, m, bits):
    """Apply matrix m on `bits` (first listed = most significant matrix index bit) of a 2^nbits vector."""
    k = len(bits)
    t = vec.reshape([2] * nbits)  # axis j <-> bit (nbits-1-j)
    axes = [nbits - 1 - b for b in bits]
    t = np.moveaxis(t, axes, list(range(k)))
    shp = t.shape
    t = (m @ t.reshape(1 << k, -1)).reshape(shp)
    t = np.moveaxis(t, list(range(k)), axes)
    return np.ascontiguousarray(t).reshape(-1)


def superop_numpy(n_qubits, gates, state=None):
    """Returns the flat complex vector v[col*dim+row] (what dm_real_res + i*dm_imag_res hold)."""
    dim = 1 << n_qubits
    v = np.zeros(dim * dim, dtype=complex)
    if state is None:
        v[0] = 1.0
    else:
        v[:] = state.reshape(-1)
    for g in gates:
        op, qb, theta, phi, lam = g[:5]
        mat = g[5] if len(g) > 5 else None
        for m, qs in primitives(op, qb, theta, phi, lam, mat):
            v = _apply(v, 2 * n_qubits, m, list(qs))
            v = _apply(v, 2 * n_qubits, m.conj(), [b + n_qubits for b in qs])
    return v
