"""dm-sim_b200 -- Python host mirror of DM-Sim's GPU-backend interface over the B200-native C-ABI.

The directory name carries a hyphen (it is the name the build was given), so import it with
``importlib.import_module("dm-sim_b200")`` or through the alias module ``dmsim_b200`` at the repo root.

Surface (same names / argument order as the reference's pybind module, src/py_nvgpu_omp_wrapper.cu:29-87):
``Gate``, ``Simulation(n_qubits, n_gpus)`` with ``append / upload / clear_circuit / run / reset / measure``
and the 38 static gate factories ``Simulation.U3(theta, phi, lam, m)`` ... ``Simulation.RYY(theta, m, n)``
(reference src/dmsim_nvgpu_omp.cuh:580-767), plus non-breaking extras (``C1``, ``C2``, ``get_dm``, ``diag``,
``trace``, ``purity``, ``sample``, ``stats``).

Everything executes in ``lib/libdmsim_b200.so`` (hand-written sm_100a kernels).  There is NO fallback: if
the library is missing or no GPU is usable the calls raise.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libdmsim_b200.so")

OP_NAMES = [
    "U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S",
    "SDG", "T", "TDG", "RX", "RY", "RZ", "CZ", "CY", "SWAP", "CH",
    "CCX", "CSWAP", "CRX", "CRY", "CRZ", "CU1", "CU3", "RXX", "RZZ", "RCCX",
    "RC3X", "C3X", "C3SQRTX", "C4X", "R", "SRN", "W", "RYY",
]
OP = {name: i for i, name in enumerate(OP_NAMES)}
OP_C1, OP_C2 = 100, 101
ALL_RANKS = -1  # DMB_ALL_RANKS: one handle drives all n_gpus devices of this process

GATE_DTYPE = np.dtype(
    [("op", "<i4"), ("qb", "<i4", (5,)), ("theta", "<f8"), ("phi", "<f8"), ("lam", "<f8"), ("mat", "<i8")],
    align=True,
)
assert GATE_DTYPE.itemsize == 56  # sizeof(dmb_gate)


class dmb_stats(ctypes.Structure):
    _fields_ = [("sim_ms", ctypes.c_double), ("comm_ms", ctypes.c_double), ("comp_ms", ctypes.c_double),
                ("n_gates", ctypes.c_uint64), ("n_primitives", ctypes.c_uint64), ("n_blocks", ctypes.c_uint64),
                ("n_sweeps", ctypes.c_uint64), ("n_exchanges", ctypes.c_uint64), ("n_launches", ctypes.c_uint64),
                ("sweep_bytes", ctypes.c_uint64), ("exchange_bytes", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64),
                ("fp64_ops", ctypes.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DMSimError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads the C-ABI library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise DMSimError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU / PyTorch fallback for this engine)")
    L = ctypes.CDLL(_LIB_PATH)
    vp, i32, u64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t
    L.dmb_create.argtypes = [i32, i32, i32, i32, ctypes.POINTER(vp)]
    L.dmb_destroy.argtypes = [vp]
    L.dmb_reset_dm.argtypes = [vp]
    L.dmb_set_dm.argtypes = [vp, vp, vp]
    L.dmb_set_circuit.argtypes = [vp, vp, sz, vp, sz]
    L.dmb_clear_circuit.argtypes = [vp]
    L.dmb_run.argtypes = [vp, ctypes.POINTER(dmb_stats)]
    L.dmb_get_dm.argtypes = [vp, vp, vp]
    L.dmb_get_diag.argtypes = [vp, vp]
    L.dmb_get_elements.argtypes = [vp, vp, sz, vp, vp]
    L.dmb_trace.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.dmb_purity.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.dmb_sample.argtypes = [vp, vp, sz, vp, ctypes.POINTER(ctypes.c_double)]
    L.dmb_measure.argtypes = [vp, ctypes.c_uint, sz, vp, ctypes.POINTER(ctypes.c_double)]
    L.dmb_comm_unique_id.argtypes = [vp]
    L.dmb_comm_init.argtypes = [vp, vp]
    L.dmb_comm_export.argtypes = [vp, vp]
    L.dmb_comm_import.argtypes = [vp, vp]
    L.dmb_comm_p2p.argtypes = [vp, i32]
    L.dmb_get_shard.argtypes = [vp, vp, vp]
    L.dmb_plan_json.argtypes = [i32, i32, vp, sz, vp, sz, vp, i32, ctypes.c_char_p, sz]
    L.dmb_plan_json.restype = ctypes.c_int64
    L.dmb_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
    L.dmb_jit_source.argtypes = [i32, i32, vp, sz, vp, sz, i32, i32, ctypes.c_char_p, sz]
    L.dmb_jit_source.restype = ctypes.c_int64
    L.dmb_query.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]
    L.dmb_jit_compile.argtypes = [i32, i32, vp, sz, vp, sz, i32, i32, i32]
    L.dmb_last_error.restype = ctypes.c_char_p
    L.dmb_version.restype = ctypes.c_char_p
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise DMSimError(f"dmsim_b200 error {rc}: {lib().dmb_last_error().decode()}")


def set_option(name: str, value: int):
    _check(lib().dmb_set_option(name.encode(), int(value)))


def query(name: str, handle=None) -> float:
    """Counters of the run-time compiler ("jit_compiled", "jit_compile_ms", ...; with a handle: "jit_sweeps", "jit_pending")."""
    v = ctypes.c_double(0)
    _check(lib().dmb_query(handle, name.encode(), ctypes.byref(v)))
    return v.value


class Gate:
    """POD mirror of DMSim::Gate (reference src/dmsim_nvgpu_omp.cuh:99-191)."""

    __slots__ = ("op_name", "qb0", "qb1", "qb2", "qb3", "qb4", "theta", "phi", "lam", "matrix")

    def __init__(self, op_name, qb0=0, qb1=0, qb2=0, qb3=0, qb4=0, theta=0.0, phi=0.0, lam=0.0, matrix=None):
        self.op_name = OP[op_name] if isinstance(op_name, str) and op_name in OP else op_name
        if isinstance(self.op_name, str):
            self.op_name = {"C1": OP_C1, "C2": OP_C2}[self.op_name]
        self.qb0, self.qb1, self.qb2, self.qb3, self.qb4 = int(qb0), int(qb1), int(qb2), int(qb3), int(qb4)
        self.theta, self.phi, self.lam = float(theta), float(phi), float(lam)
        self.matrix = None if matrix is None else np.asarray(matrix, dtype=np.complex128)

    def dump(self) -> str:
        # reference Gate::dump :170-176 ("<<" of a double = %g)
        name = OP_NAMES[self.op_name] if self.op_name < 100 else ("C1" if self.op_name == OP_C1 else "C2")
        return "%s(%d,%d,%d,%d,%d,%g,%g,%g);\n" % (name, self.qb0, self.qb1, self.qb2, self.qb3, self.qb4,
                                                   self.theta, self.phi, self.lam)


def pack_gates(gates):
    """[Gate | (op, qubits, theta, phi, lam[, matrix])] -> (structured array, matrix table)."""
    rec = np.zeros(len(gates), dtype=GATE_DTYPE)
    table = []
    for i, g in enumerate(gates):
        if isinstance(g, Gate):
            op, qb, th, ph, la, mat = g.op_name, [g.qb0, g.qb1, g.qb2, g.qb3, g.qb4], g.theta, g.phi, g.lam, g.matrix
        else:
            op, qb, th, ph, la = g[:5]
            mat = g[5] if len(g) > 5 else None
            if isinstance(op, str):
                op = {"C1": OP_C1, "C2": OP_C2}.get(op, OP.get(op))
                if op is None:
                    raise DMSimError(f"unknown op {g[0]!r}")
            qb = list(qb) + [0] * (5 - len(qb))
        rec[i]["op"] = op
        rec[i]["qb"] = qb
        rec[i]["theta"], rec[i]["phi"], rec[i]["lam"] = th, ph, la
        if op >= 100:
            flat = np.asarray(mat, dtype=np.complex128).reshape(-1)
            slot = np.zeros(32)
            slot[0:2 * flat.size:2] = flat.real
            slot[1:2 * flat.size:2] = flat.imag
            rec[i]["mat"] = len(table)
            table.append(slot)
    mats = np.ascontiguousarray(np.concatenate(table)) if table else np.zeros(0)
    return rec, mats


def plan_json(n_qubits, world_size, gates, start_layout=None, conj_state=False, non_hermitian=False) -> dict:
    """The schedule the engine would run (host-only planner; works without a GPU)."""
    rec, mats = gates if isinstance(gates, tuple) else pack_gates(gates)
    L = lib()
    lay = None if start_layout is None else np.ascontiguousarray(start_layout, dtype=np.int32)
    args = (n_qubits, world_size, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32,
            lay.ctypes.data if lay is not None else None, int(bool(conj_state)) | (int(bool(non_hermitian)) << 1))
    need = L.dmb_plan_json(*args, None, 0)
    if need < 0:
        _check(int(need))
    buf = ctypes.create_string_buffer(int(need))
    _check(min(0, int(L.dmb_plan_json(*args, buf, int(need)))))
    return json.loads(buf.value.decode())


def jit_source(n_qubits, world_size, gates, sweep_index, peer=False):
    """(defines, program): the CUDA text of the run-time specialised kernel of one sweep of the plan (host only), or None
    when there is no such sweep / the generator does not cover it."""
    rec, mats = gates if isinstance(gates, tuple) else pack_gates(gates)
    L = lib()
    args = (n_qubits, world_size, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32,
            int(sweep_index), int(bool(peer)))
    need = L.dmb_jit_source(*args, None, 0)
    if need < 0:
        _check(int(need))
    if need == 0:
        return None
    buf = ctypes.create_string_buffer(int(need))
    _check(min(0, int(L.dmb_jit_source(*args, buf, int(need)))))
    defines, program = buf.value.decode().split("//---- program\n", 1)
    return defines, program


def jit_compile(n_qubits, world_size, gates, sweep_index, peer=False, wait=True) -> int:
    """Builds the specialised kernel of one sweep with the library's run-time compiler (no GPU needed): 1 built, 0 queued,
    -1 failed, -2 no such sweep."""
    rec, mats = gates if isinstance(gates, tuple) else pack_gates(gates)
    rc = lib().dmb_jit_compile(n_qubits, world_size, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None, mats.size // 32,
                               int(sweep_index), int(bool(peer)), int(bool(wait)))
    if rc < -2:
        _check(rc)
    return rc


class Simulation:
    """Drop-in for DMSim::Simulation (reference src/dmsim_nvgpu_omp.cuh:193-814) as seen from Python.

    ``Simulation(n_qubits, n_gpus)``: with n_gpus == 1 everything runs on the current CUDA device.  With
    n_gpus > 1 and no ``rank`` the object drives devices 0 .. n_gpus-1 from THIS process, like the reference
    (src/py_nvgpu_omp_wrapper.cu:36-37) -- unless a ``torch.distributed`` job with world_size == n_gpus is
    running (torchrun, one process per GPU): then, or when ``rank`` is given, the object is ONE RANK of that job
    (see ``attach_communicator``) and its result calls are collective.
    """

    def __init__(self, n_qubits, n_gpus=1, rank=None, device=-1):
        self.n_qubits, self.n_gpus = int(n_qubits), int(n_gpus)
        self.dim = 1 << self.n_qubits
        self.circuit = []
        self._uploaded = False
        self.last_stats = None
        self.group = False
        if rank is None:
            rank = 0
            if self.n_gpus > 1:
                dist = getattr(sys.modules.get("torch"), "distributed", None)
                if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() == self.n_gpus:
                    rank = dist.get_rank()
                else:
                    rank, self.group = ALL_RANKS, True
        self.rank = rank
        h = ctypes.c_void_p()
        _check(lib().dmb_create(self.n_qubits, self.n_gpus, self.rank, device, ctypes.byref(h)))
        self._h = h
        if self.n_gpus > 1 and not self.group:
            self.attach_communicator()

    def attach_communicator(self):
        """NCCL bootstrap: rank 0 makes the unique id, torch.distributed (plumbing) broadcasts it."""
        import torch
        import torch.distributed as dist
        ident = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            _check(lib().dmb_comm_unique_id(ident.ctypes.data))
        obj = [ident.tobytes()]
        dist.broadcast_object_list(obj, src=0)
        ident = np.frombuffer(obj[0], dtype=np.uint8).copy()
        _check(lib().dmb_comm_init(self._h, ident.ctypes.data))
        # peer-memory exchange (fused pack + all-to-all over NVLink); DMB_P2P=0 keeps the NCCL send/recv path.  Every
        # rank must end up with the same form: the outcome of the import is agreed on collectively.
        self.p2p = False
        if os.environ.get("DMB_P2P", "1") != "0":
            mine = np.zeros(128, dtype=np.uint8)
            ok = lib().dmb_comm_export(self._h, mine.ctypes.data) == 0
            everyone = [None] * self.n_gpus
            dist.all_gather_object(everyone, mine.tobytes())
            blob = np.frombuffer(b"".join(everyone), dtype=np.uint8).copy()
            ok = ok and lib().dmb_comm_import(self._h, blob.ctypes.data) == 0
            why = "" if ok else lib().dmb_last_error().decode()
            flag = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            self.p2p = bool(int(flag.item()))
            _check(lib().dmb_comm_p2p(self._h, 1 if self.p2p else 0))
            if not self.p2p and self.rank == 0:
                print(f"dmsim_b200: peer-memory remap unavailable ({why or 'another rank failed'}); using NCCL send/recv",
                      file=sys.stderr)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and _lib is not None:
            _lib.dmb_destroy(h)
            self._h = None

    # ---- circuit building (reference :331-377, :496-520) ----
    def append(self, g):
        for q in (g.qb0, g.qb1, g.qb2, g.qb3, g.qb4):  # append()'s asserts :334-338
            if not (0 <= q < self.n_qubits):
                raise DMSimError(f"qubit index {q} out of range for {self.n_qubits} qubits")
        self.circuit.append(g)

    @property
    def n_gates(self):
        return len(self.circuit)

    def upload(self):
        if self._uploaded:
            raise DMSimError("upload() called twice without clear_circuit()")  # reference asserts :349-350
        rec, mats = pack_gates(self.circuit)
        _check(lib().dmb_set_circuit(self._h, rec.ctypes.data, len(rec), mats.ctypes.data if mats.size else None,
                                     mats.size // 32))
        self._uploaded = True
        return self

    def clear_circuit(self):
        self.circuit = []
        self._uploaded = False
        _check(lib().dmb_clear_circuit(self._h))

    def reset_dm(self):
        _check(lib().dmb_reset_dm(self._h))

    def reset(self):
        self.clear_circuit()
        self.reset_dm()

    def dump(self) -> str:
        return "".join(g.dump() for g in self.circuit)

    # ---- execution ----
    def run(self):
        if not self._uploaded:
            raise DMSimError("run() before upload()")
        st = dmb_stats()
        _check(lib().dmb_run(self._h, ctypes.byref(st)))
        self.last_stats = st.asdict()

    sim = run

    # ---- results ----
    def measure(self, repetition=10, seed=None):
        """List of basis-state indices, as the reference's measure(); seed defaults to time(0) like RAND_SEED."""
        import time
        out = np.zeros(repetition, dtype=np.uint64)
        total = ctypes.c_double()
        seed = int(time.time()) if seed is None else int(seed)
        _check(lib().dmb_measure(self._h, seed & 0xFFFFFFFF, repetition, out.ctypes.data, ctypes.byref(total)))
        if abs(total.value - 1.0) > 1e-3:  # ERROR_BAR, reference :544-545
            print("Sum of probability along diag is far from 1.0 with %f" % total.value)
        return [int(x) for x in out]

    def sample(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.zeros(len(r), dtype=np.uint64)
        total = ctypes.c_double()
        _check(lib().dmb_sample(self._h, r.ctypes.data, len(r), out.ctypes.data, ctypes.byref(total)))
        return out, total.value

    def get_dm(self):
        """(real, imag), each (dim, dim) indexed [col][row]: what dm_real_res / dm_imag_res hold (rho^T)."""
        re = np.empty((self.dim, self.dim))
        im = np.empty((self.dim, self.dim))
        _check(lib().dmb_get_dm(self._h, re.ctypes.data, im.ctypes.data))
        return re, im

    def set_dm(self, real, imag):
        real = np.ascontiguousarray(real, dtype=np.float64)
        imag = np.ascontiguousarray(imag, dtype=np.float64)
        assert real.size == self.dim * self.dim and imag.size == self.dim * self.dim
        _check(lib().dmb_set_dm(self._h, real.ctypes.data, imag.ctypes.data))

    def elements(self, flat_index):
        """dm_real_res[f] + 1j * dm_imag_res[f] for f = col*dim + row (spot checks without the full copy-back)."""
        idx = np.ascontiguousarray(flat_index, dtype=np.uint64)
        re, im = np.empty(idx.size), np.empty(idx.size)
        _check(lib().dmb_get_elements(self._h, idx.ctypes.data, idx.size, re.ctypes.data, im.ctypes.data))
        return re + 1j * im

    def diag(self):
        d = np.empty(self.dim)
        _check(lib().dmb_get_diag(self._h, d.ctypes.data))
        return d

    def trace(self):
        v = ctypes.c_double()
        _check(lib().dmb_trace(self._h, ctypes.byref(v)))
        return v.value

    def purity(self):
        v = ctypes.c_double()
        _check(lib().dmb_purity(self._h, ctypes.byref(v)))
        return v.value

    def shard(self):
        """(interleaved complex shard in PHYSICAL order, phys_of_logical[2n]); a single-process group returns all its
        shards back to back (the whole state in physical order)."""
        elems = (self.dim * self.dim) // (1 if self.group else self.n_gpus)
        data = np.empty(elems, dtype=np.complex128)
        lay = np.zeros(2 * self.n_qubits, dtype=np.int32)
        _check(lib().dmb_get_shard(self._h, data.ctypes.data, lay.ctypes.data))
        return data, lay

    # ---- the 38 factories, parameter-first then qubits (reference :580-767) + C1/C2 ----
    @staticmethod
    def U3(theta, phi, lam, m): return Gate(OP["U3"], m, theta=theta, phi=phi, lam=lam)
    @staticmethod
    def U2(phi, lam, m): return Gate(OP["U2"], m, phi=phi, lam=lam)
    @staticmethod
    def U1(lam, m): return Gate(OP["U1"], m, lam=lam)
    @staticmethod
    def CX(m, n): return Gate(OP["CX"], m, n)
    @staticmethod
    def ID(m): return Gate(OP["ID"], m)
    @staticmethod
    def X(m): return Gate(OP["X"], m)
    @staticmethod
    def Y(m): return Gate(OP["Y"], m)
    @staticmethod
    def Z(m): return Gate(OP["Z"], m)
    @staticmethod
    def H(m): return Gate(OP["H"], m)
    @staticmethod
    def S(m): return Gate(OP["S"], m)
    @staticmethod
    def SDG(m): return Gate(OP["SDG"], m)
    @staticmethod
    def T(m): return Gate(OP["T"], m)
    @staticmethod
    def TDG(m): return Gate(OP["TDG"], m)
    @staticmethod
    def RX(theta, m): return Gate(OP["RX"], m, theta=theta)
    @staticmethod
    def RY(theta, m): return Gate(OP["RY"], m, theta=theta)
    @staticmethod
    def RZ(phi, m): return Gate(OP["RZ"], m, phi=phi)
    @staticmethod
    def CZ(m, n): return Gate(OP["CZ"], m, n)
    @staticmethod
    def CY(m, n): return Gate(OP["CY"], m, n)
    @staticmethod
    def SWAP(m, n): return Gate(OP["SWAP"], m, n)
    @staticmethod
    def CH(m, n): return Gate(OP["CH"], m, n)
    @staticmethod
    def CCX(l, m, n): return Gate(OP["CCX"], l, m, n)
    @staticmethod
    def CSWAP(l, m, n): return Gate(OP["CSWAP"], l, m, n)
    @staticmethod
    def CRX(lam, m, n): return Gate(OP["CRX"], m, n, lam=lam)
    @staticmethod
    def CRY(lam, m, n): return Gate(OP["CRY"], m, n, lam=lam)
    @staticmethod
    def CRZ(lam, m, n): return Gate(OP["CRZ"], m, n, lam=lam)
    @staticmethod
    def CU1(lam, m, n): return Gate(OP["CU1"], m, n, lam=lam)
    @staticmethod
    def CU3(theta, phi, lam, m, n): return Gate(OP["CU3"], m, n, theta=theta, phi=phi, lam=lam)
    @staticmethod
    def RXX(theta, m, n): return Gate(OP["RXX"], m, n, theta=theta)
    @staticmethod
    def RZZ(theta, m, n): return Gate(OP["RZZ"], m, n, theta=theta)
    @staticmethod
    def RCCX(l, m, n): return Gate(OP["RCCX"], l, m, n)
    @staticmethod
    def RC3X(l, m, n, o): return Gate(OP["RC3X"], l, m, n, o)
    @staticmethod
    def C3X(l, m, n, o): return Gate(OP["C3X"], l, m, n, o)
    @staticmethod
    def C3SQRTX(l, m, n, o): return Gate(OP["C3SQRTX"], l, m, n, o)
    @staticmethod
    def C4X(l, m, n, o, p): return Gate(OP["C4X"], l, m, n, o, p)
    @staticmethod
    def R(theta, m): return Gate(OP["R"], m, theta=theta)
    @staticmethod
    def SRN(m): return Gate(OP["SRN"], m)
    @staticmethod
    def W(m): return Gate(OP["W"], m)
    @staticmethod
    def RYY(theta, m, n): return Gate(OP["RYY"], m, n, theta=theta)
    # the reference's dead-code generic gates (C1_GATE :1004-1025, C2_GATE :1028-1122), made reachable
    @staticmethod
    def C1(matrix, m): return Gate(OP_C1, m, matrix=np.asarray(matrix, dtype=np.complex128).reshape(2, 2))
    @staticmethod
    def C2(matrix, m, n): return Gate(OP_C2, m, n, matrix=np.asarray(matrix, dtype=np.complex128).reshape(4, 4))


def print_measurement(res_state, n_qubits, repetition):
    """reference src/util_nvgpu.cuh:145-155 (MSB-first bit strings)."""
    print("\n===============  Measurement (tests=%d) ================" % repetition)
    for i in range(repetition):
        print("Test-%d: %s" % (i, format(int(res_state[i]), "0%db" % n_qubits)))
