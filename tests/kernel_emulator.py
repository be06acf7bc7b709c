"""TEST INFRASTRUCTURE: a numpy mirror of sweep_kernel's INDEXING (dm-sim_b200/csrc/kernels.cu).

It consumes the device tables the host encoder produces ("dev" objects of dmb_plan_json: pre-swizzled lane /
iteration / warp / member tables, class payloads, load/store address tables) and executes a sweep exactly the
way the CUDA kernel walks it (tile id -> base address, thread -> element, warp group -> sub-tile, lane/iter ->
work item), so that the encoder and the kernel's addressing scheme are checked against the oracle on CPU.
Never imported by the product.
"""
from __future__ import annotations

import numpy as np

CLS_DENSE1, CLS_DIAG1, CLS_MONO1, CLS_SRN1, CLS_DENSE2, CLS_DIAG2, CLS_MONO2 = range(7)
NT = 256


def swz(e):
    return e ^ ((e >> 3) & 7)


def _dep(v, pos):
    r = np.zeros_like(v, dtype=np.int64)
    for i, p in enumerate(pos):
        r |= ((v >> i) & 1) << int(p)
    return r


def run_sweep(dev: dict, shard_in: np.ndarray, shard_out: np.ndarray | None = None) -> np.ndarray:
    """One sweep_kernel launch over a shard given in PHYSICAL order; returns the output shard."""
    k, n_comp = dev["k"], dev["n_comp"]
    tile_elems, n_tiles = 1 << k, 1 << n_comp
    assert shard_in.size == tile_elems * n_tiles
    out = shard_in.copy() if shard_out is None else shard_out
    t = np.arange(NT, dtype=np.int64)
    klo = min(k, 8)
    n_it = 1 if k <= 8 else 1 << (k - 8)
    act = t < tile_elems
    g_in_lo = _dep(t, dev["gin"][:klo])
    g_out_lo = _dep(t, dev["gout"][:klo])
    s_out_lo = swz(_dep(t, dev["sout"][:klo]))
    s_in = swz(t)
    tid = np.arange(n_tiles, dtype=np.int64)
    base_in = _dep(tid, dev["cin"])
    base_out = _dep(tid, dev["cout"])

    tiles = np.full((n_tiles, tile_elems), np.nan + 0j, dtype=np.complex128)
    for it in range(n_it):
        src = (base_in[:, None] | g_in_lo[None, act]) + int(dev["hin"][it])
        tiles[:, (it << 8) | s_in[act]] = shard_in[src]
    assert not np.isnan(tiles.real).any(), "load did not fill the tile"

    for grp in dev["groups"]:
        for w in range(grp["n_warps"]):
            wpart = grp["wtab"][w]
            for o in range(grp["first"], grp["first"] + grp["count"]):
                _apply_op(dev["ops"][o], tiles, wpart)

    seen = np.zeros(shard_in.size, dtype=bool)
    for it in range(n_it):
        dst = (base_out[:, None] | g_out_lo[None, act]) + int(dev["hout"][it])
        assert not seen[dst].any()
        seen[dst] = True
        out[dst] = tiles[:, s_out_lo[act] ^ int(dev["hs"][it])]
    assert seen.all(), "store did not cover the shard"
    return out


def _apply_op(op, tiles, wpart):
    cls, aux, n_iter, n_active = op["cls"], op["aux"], op["n_iter"], op["n_active"]
    m = np.array(op["m"], dtype=np.float64)
    m = m[0::2] + 1j * m[1::2]
    lane = np.array(op["lane_tab"][:n_active], dtype=np.int64)
    itab = np.array(op["iter_tab"][:n_iter], dtype=np.int64)
    x = ((lane[:, None] ^ wpart) ^ itab[None, :]).reshape(-1)  # member-0 index of every work item
    off = [int(v) for v in op["off"]]
    nmem = 4 if cls >= CLS_DENSE2 else 2
    idx = [x ^ off[c] for c in range(nmem)]
    allidx = np.concatenate(idx)
    assert len(np.unique(allidx)) == allidx.size, "work items of one warp overlap"
    v = [tiles[:, i].copy() for i in idx]
    skip = (aux >> 8) & 15
    unit = (aux >> 12) & 1
    if cls == CLS_DENSE2:
        for r in range(4):
            tiles[:, idx[r]] = sum(m[4 * r + c] * v[c] for c in range(4))
    elif cls == CLS_DENSE1:
        tiles[:, idx[0]] = m[0] * v[0] + m[1] * v[1]
        tiles[:, idx[1]] = m[2] * v[0] + m[3] * v[1]
    elif cls in (CLS_DIAG2, CLS_DIAG1):
        for r in range(nmem):
            if not (skip >> r) & 1:
                tiles[:, idx[r]] = m[r] * v[r]
    elif cls == CLS_MONO2:
        for r in range(4):
            if not (skip >> r) & 1:
                s = (aux >> (2 * r)) & 3
                tiles[:, idx[r]] = v[s] if unit else m[r] * v[s]
    elif cls == CLS_MONO1:
        tiles[:, idx[0]] = v[1] if unit else m[0] * v[1]
        tiles[:, idx[1]] = v[0] if unit else m[1] * v[0]
    elif cls == CLS_SRN1:
        re = 0.5 * (v[0].real + v[1].real)
        tiles[:, idx[0]] = re + 1j * 0.5 * (v[0].imag - v[1].imag)
        tiles[:, idx[1]] = re + 1j * 0.5 * (-v[0].imag + v[1].imag)
    else:
        raise ValueError(cls)


def check_group_partition(dev: dict):
    """Every group's warps x lanes x iterations x members must tile the 2^k elements exactly once per op."""
    k = dev["k"]
    for grp in dev["groups"]:
        for o in range(grp["first"], grp["first"] + grp["count"]):
            op = dev["ops"][o]
            nmem = 4 if op["cls"] >= CLS_DENSE2 else 2
            lane = np.array(op["lane_tab"][:op["n_active"]], dtype=np.int64)
            itab = np.array(op["iter_tab"][:op["n_iter"]], dtype=np.int64)
            wt = np.array(grp["wtab"][:grp["n_warps"]], dtype=np.int64)
            x = (wt[:, None, None] ^ lane[None, :, None] ^ itab[None, None, :]).reshape(-1)
            allidx = np.concatenate([x ^ int(op["off"][c]) for c in range(nmem)])
            assert allidx.size == 1 << k and len(np.unique(allidx)) == 1 << k, (grp, op["cls"])


def run_plan_dev(plan: dict, vec_logical: np.ndarray) -> np.ndarray:
    """Whole plan through the kernel mirror: per-rank shards, exchanges as block transposes of shards."""
    from plan_emulator import logical_to_physical, physical_to_logical
    n, g = plan["n"], plan["g"]
    N, M = 2 * n, 2 * n - g
    P = 1 << g
    v = np.asarray(vec_logical, dtype=np.complex128).reshape(-1)
    if plan.get("conj_start"):
        v = np.conj(v)
    v = logical_to_physical(v, plan["start_layout"])
    shards = [v[r << M:(r + 1) << M].copy() for r in range(P)]
    for st in plan["steps"]:
        if st["kind"] == "exchange":
            chunk = (1 << M) // P
            new = [np.empty_like(s) for s in shards]
            for r in range(P):
                for p in range(P):  # rank r sends its chunk p to rank p, which stores it as chunk r
                    new[p][r * chunk:(r + 1) * chunk] = shards[r][p * chunk:(p + 1) * chunk]
            shards = new
            continue
        check_group_partition(st["dev"])
        for r in range(P):
            if st["out_of_place"]:
                shards[r] = run_sweep(st["dev"], shards[r], np.full_like(shards[r], np.nan))
            else:
                shards[r] = run_sweep(st["dev"], shards[r])
    v = physical_to_logical(np.concatenate(shards), plan["end_layout"])
    return np.conj(v) if plan.get("conj_end") else v
