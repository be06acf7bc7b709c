"""TEST INFRASTRUCTURE: a numpy interpreter for the schedules the C++ planner emits (dmb_plan_json).

It executes a plan step by step on the full 2n-bit vector (all ranks at once: the rank is the top g
physical bits), using the FULL op matrices the planner reports -- independent of the device op encodings.
Lets the CPU test-suite check expansion + fusion + scheduling + qubit-remap bookkeeping against the oracle
without a GPU.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np


def _apply_matrix(vec, nbits, m, bits):
    """m acts on `bits` (first = most significant matrix index bit) of a 2^nbits vector."""
    k = len(bits)
    t = vec.reshape([2] * nbits)
    axes = [nbits - 1 - b for b in bits]
    t = np.moveaxis(t, axes, list(range(k)))
    shp = t.shape
    t = (m @ t.reshape(1 << k, -1)).reshape(shp)
    t = np.moveaxis(t, list(range(k)), axes)
    return np.ascontiguousarray(t).reshape(-1)


def _apply_srn(vec, nbits, bit):
    """reference SRN_GATE (src/dmsim_nvgpu_omp.cuh:1253-1266): v0' = (v0 + conj(v1))/2, v1' = (conj(v0) + v1)/2."""
    t = vec.reshape([2] * nbits)
    ax = nbits - 1 - bit
    t = np.moveaxis(t, ax, 0)
    v0, v1 = t[0].copy(), t[1].copy()
    out = np.empty_like(t)
    out[0] = 0.5 * (v0 + np.conj(v1))
    out[1] = 0.5 * (np.conj(v0) + v1)
    return np.ascontiguousarray(np.moveaxis(out, 0, ax)).reshape(-1)


def _permute_bits(vec, nbits, src_to_dst):
    """new[index with bit dst] = old[index with bit src] for a bijection src->dst on bit positions."""
    t = vec.reshape([2] * nbits)
    # result axis for bit d must come from source axis of bit s where src_to_dst[s] = d
    inv = {d: s for s, d in src_to_dst.items()}
    axes = [nbits - 1 - inv.get(nbits - 1 - i, nbits - 1 - i) for i in range(nbits)]
    return np.ascontiguousarray(np.transpose(t, axes)).reshape(-1)


def logical_to_physical(vec_logical, layout):
    """vec_phys[P(L)] = vec_logical[L], P moves logical bit l to physical bit layout[l]."""
    n = len(layout)
    return _permute_bits(vec_logical, n, {l: int(layout[l]) for l in range(n)})


def physical_to_logical(vec_phys, layout):
    n = len(layout)
    return _permute_bits(vec_phys, n, {int(layout[l]): l for l in range(n)})


def run_plan(plan: dict, vec_logical: np.ndarray) -> np.ndarray:
    """Runs the plan on a state given in LOGICAL order; returns the result in LOGICAL order."""
    n, g = plan["n"], plan["g"]
    N, M = 2 * n, 2 * n - g
    v = np.asarray(vec_logical, dtype=np.complex128).reshape(-1)
    if plan.get("conj_start"):
        v = np.conj(v)
    v = logical_to_physical(v, plan["start_layout"])
    for st in plan["steps"]:
        if st["kind"] == "exchange":
            mp = {}
            for i in range(g):
                mp[M - g + i] = M + i
                mp[M + i] = M - g + i
            v = _permute_bits(v, N, mp)
            continue
        in_pos, out_pos = st["in_pos"], st["out_pos"]
        assert sorted(in_pos) == sorted(out_pos), "a sweep may only permute bits inside its tile"
        assert all(p < M for p in in_pos), "tile bits must be shard-local"
        assert in_pos == sorted(in_pos)
        assert len(in_pos) == st["k"]
        for op in st["ops"]:
            nb = op["nb"]
            if op["cls"] == 3:
                v = _apply_srn(v, N, in_pos[op["j0"]])
                continue
            m = np.array(op["m"], dtype=np.float64)
            m = (m[0::2] + 1j * m[1::2]).reshape(1 << nb, 1 << nb)
            bits = [in_pos[op["j0"]]]
            if nb == 2:  # a controlled phase may have its second bit outside the tile (even on a rank bit)
                bits.append(in_pos[op["j1"]] if op["j1"] >= 0 else op["p1"])
            v = _apply_matrix(v, N, m, bits)
        if in_pos != out_pos:  # out of place (remap pack) or a permutation inside the tile done in place
            v = _permute_bits(v, N, {in_pos[j]: out_pos[j] for j in range(len(in_pos))})
    v = physical_to_logical(v, plan["end_layout"])
    return np.conj(v) if plan.get("conj_end") else v


def run_circuit(n_qubits, world_size, gates, plan_fn, state=None):
    """Plans `gates` with plan_fn(n, world, gates) and runs it from |0..0><0..0| (or `state`, flat logical)."""
    dim = 1 << n_qubits
    v = np.zeros(dim * dim, dtype=np.complex128)
    if state is None:
        v[0] = 1.0
    else:
        v[:] = np.asarray(state).reshape(-1)
    plan = plan_fn(n_qubits, world_size, gates)
    return run_plan(plan, v), plan
