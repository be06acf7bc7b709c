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
# CTA geometry of the build (devop.hpp DMB_THREAD_BITS / DMB_REG_BITS), taken from the "geom" object of every sweep
NT = 256
TB = 8  # log2(NT)
E = 8   # elements a lane keeps in registers per round (2^kRegBits)
WB = 3  # log2(warps per CTA)
RB = 3  # register bits
SW = 16 # star table entries per slot: (warp << iteration bits) | iteration


def _set_geom(dev):
    global NT, TB, E, WB, RB, SW
    g = dev.get("geom")
    if g:
        TB, RB, WB, SW = g["thread_bits"], g["reg_bits"], g["warp_bits"], g["star_w"]
        NT, E = 1 << TB, 1 << RB


def swz(e, mode=0):
    """mode 0: the 3-term XOR swizzle; mode 1 (TMA tiles): the hardware's 128-byte swizzle (devop.hpp swz_host)."""
    if mode == 1:
        return e ^ ((e >> 3) & 7)
    return e ^ ((e >> 3) & 7) ^ ((e >> 6) & 7) ^ ((e >> 9) & 7)


def _tma_load(dev, shard_in, base_in):
    """The tile as the TMA path fills it: copy j = one box at byte offset j * box_bytes; inside a box the dimensions are
    laid out dimension 0 fastest; the 16-byte chunk of every 128-byte row is XORed with (row & 7)."""
    g = dev["tma"]
    k = dev["k"]
    tiles = np.full((base_in.size, 1 << k), np.nan + 0j, dtype=np.complex128)
    box_elems = g["box_bytes"] // 16
    e = np.arange(box_elems, dtype=np.int64)
    off = np.zeros_like(e)  # element offset of box-local index e in the shard
    bit = 0
    for d in range(5):
        for b in range(g["box_log2"][d]):
            off |= ((e >> bit) & 1) << (g["start"][d] + b)
            bit += 1
    assert (1 << bit) == box_elems
    for j in range(g["n_copies"]):
        src = (base_in[:, None] | int(g["enum_off"][j])) + off[None, :]
        local = j * box_elems + e
        tiles[:, swz(local, 1)] = shard_in[src]
    return tiles


def _dep(v, pos):
    r = np.zeros_like(v, dtype=np.int64)
    for i, p in enumerate(pos):
        r |= ((v >> i) & 1) << int(p)
    return r


def run_sweep(dev: dict, shard_in: np.ndarray, shard_out: np.ndarray | None = None, rank: int = 0) -> np.ndarray:
    """One sweep_kernel launch over a shard given in PHYSICAL order; returns the output shard."""
    _set_geom(dev)
    k, n_comp = dev["k"], dev["n_comp"]
    tile_elems, n_tiles = 1 << k, 1 << n_comp
    assert shard_in.size == tile_elems * n_tiles
    out = shard_in.copy() if shard_out is None else shard_out
    t = np.arange(NT, dtype=np.int64)
    klo = min(k, TB)
    n_it = 1 if k <= TB else 1 << (k - TB)
    act = t < tile_elems
    g_in_lo = _dep(t, dev["gin"][:klo])
    g_out_lo = _dep(t, dev["gout"][:klo])
    mode = dev.get("swz", 0)
    s_out_lo = swz(_dep(t, dev["sout"][:klo]), mode)
    s_in = swz(t, mode)
    tid = np.arange(n_tiles, dtype=np.int64)
    base_in = _dep(tid, dev["cin"])
    base_out = _dep(tid, dev["cout"])

    if dev.get("tma_load"):
        tiles = _tma_load(dev, shard_in, base_in)
    else:
        tiles = np.full((n_tiles, tile_elems), np.nan + 0j, dtype=np.complex128)
        for it in range(n_it):
            src = (base_in[:, None] | g_in_lo[None, act]) + int(dev["hin"][it])
            tiles[:, swz(it << TB, mode) ^ s_in[act]] = shard_in[src]
    assert not np.isnan(tiles.real).any(), "load did not fill the tile"

    full = base_in | (rank << (k + n_comp))  # the full physical index of the tile's first element (rank bits on top)
    direct = dev.get("direct")
    seen = np.zeros(shard_in.size, dtype=bool)
    n_rounds = len(dev["rounds"])
    for grp in dev["groups"]:
        for w in range(grp["n_warps"]):
            wpart = grp["wtab"][w]
            for ri in range(grp["first"], grp["first"] + grp["count"]):
                store = (direct, out, base_out, seen) if direct and ri == n_rounds - 1 else None
                _run_round(dev, dev["rounds"][ri], tiles, wpart, w, full, store)
    if direct:  # the last round wrote its results straight to the shard (DevDirect)
        assert seen.all(), "direct store did not cover the shard"
        return out

    for it in range(n_it):
        dst = (base_out[:, None] | g_out_lo[None, act]) + int(dev["hout"][it])
        assert not seen[dst].any()
        seen[dst] = True
        out[dst] = tiles[:, s_out_lo[act] ^ int(dev["hs"][it])]
    assert seen.all(), "store did not cover the shard"
    return out


(RC_DENSE1, RC_DIAG1, RC_MONO1, RC_SRN1, RC_DENSE2, RC_DIAG2, RC_PERM2, RC_DIAGR, RC_DENSE1_RR, RC_DENSE1_RI, RC_STAR,
 RC_HAD, RC_DIAGP, RC_CP2, RC_QFT2, RC_DENSE2_LU) = range(16)
_POS2 = {0: (1, 0), 1: (2, 0), 2: (2, 1), 3: (3, 0), 4: (3, 1), 5: (3, 2)}
_PERMS = {0: [0, 1, 3, 2], 1: [0, 3, 2, 1], 2: [0, 2, 1, 3]}


def _round_items(rd, wpart):
    lane = np.array(rd["lane_tab"][:rd["n_active"]], dtype=np.int64)
    itab = np.array(rd["iter_tab"][:rd["n_iter"]], dtype=np.int64)
    return ((lane[:, None] ^ wpart) ^ itab[None, :]).reshape(-1)


def _run_round(dev, rd, tiles, wpart, warp, full, store=None):
    """One register round of one warp: load 16 elements per work item, apply the round's ops, store back."""
    base = _round_items(rd, wpart)
    roff = [int(x) for x in rd["roff"]]
    idx = [base ^ roff[c] for c in range(E)]
    v = [tiles[:, i].copy() for i in idx]  # v[c]: (n_tiles, n_items)
    # work item -> (lane, iteration), as _round_items flattens them; iw = (warp << iteration bits) | iteration
    n_iter = rd["n_iter"]
    lane = np.repeat(np.arange(rd["n_active"]), n_iter)
    iw = (warp << (n_iter.bit_length() - 1)) | np.tile(np.arange(n_iter), rd["n_active"])
    first = next(i for i, op in enumerate(dev["ops"]) if op["off"] == rd["first"]) if rd["count"] else 0
    for o in range(first, first + rd["count"]):
        op = dev["ops"][o]
        if op["code"] == RC_STAR:
            _apply_star(dev, op, v, lane, iw, full)
        else:
            _apply_reg_op(op, v)
    if store is not None:
        # direct store: global element = tile base | lane bits | warp bits | iteration bits | register bits, each at the
        # physical position the encoder recorded for the last round
        direct, out, base_out, seen = store
        g = _dep(lane, direct["lane_pos"]) | _dep(np.full_like(lane, warp), direct["warp_pos"]) | \
            _dep(np.tile(np.arange(n_iter), rd["n_active"]), direct["iter_pos"])
        for c in range(E):
            dst = base_out[:, None] | (g | int(_dep(np.array([c]), direct["reg_pos"])[0]))[None, :]
            assert not seen[dst].any()
            seen[dst] = True
            out[dst] = v[c]
        return
    for c in range(E):  # duplicates (tiny tiles) carry identical values
        tiles[:, idx[c]] = v[c]


def _cplx(a):
    a = np.asarray(a, dtype=np.float64)
    return a[0::2] + 1j * a[1::2]


def _apply_star(dev, op, v, lane, iw, full):
    """RC_STAR: elements with register bit p set get L_p[lane] * WO_p[tile, iw]; WO folds the outside partner bits."""
    slot = op["star"]
    for p in range(RB):
        if not (op["aux"] >> p) & 1:
            continue
        st = dev["stars"][slot]
        slot += 1
        w, phi = _cplx(st["w"]), _cplx(st["phi"])
        l = _cplx(st["la"])[np.arange(32) & 7] * _cplx(st["lb"])[np.arange(32) >> 3]
        wo = np.tile(w[None, :], (full.size, 1))  # (n_tiles, 8)
        for j, b in enumerate(st["bit"]):
            sel = ((full >> int(b)) & 1).astype(bool)
            wo[sel, :] *= phi[j]
        ph = l[lane][None, :] * wo[:, iw]  # (n_tiles, n_items)
        for c in range(E):
            if c & (1 << p):
                v[c] = ph * v[c]


def _apply_reg_op(op, v):
    code, aux, pos = op["code"], op["aux"], op["pos"]
    m = np.array(op["m"], dtype=np.float64)
    m = m[0::2] + 1j * m[1::2]
    skip, unit = (aux >> 8) & 15, (aux >> 12) & 1
    if code == RC_DIAGR:
        for c in range(E):
            if not (aux >> c) & 1:
                v[c] = m[c] * v[c]
        return
    if code == RC_DIAGP:  # 8 entries over the other three register bits, for the elements with register bit `pos` set
        for j in range(E // 2):
            c = ((j >> pos) << (pos + 1)) | (1 << pos) | (j & ((1 << pos) - 1))
            if not (aux >> j) & 1:
                v[c] = m[j] * v[c]
        return
    if code == RC_HAD:  # unscaled butterflies on every register bit of the mask; the scale sits in a dense op of the sweep
        for pb in range(RB):
            if not (aux >> pb) & 1:
                continue
            b = 1 << pb
            for q in range(E):
                if q & b:
                    continue
                a0, a1 = v[q].copy(), v[q | b].copy()
                v[q] = a0 + a1
                v[q | b] = v[q] - 2.0 * a1
        return
    if code in (RC_DENSE1, RC_DIAG1, RC_MONO1, RC_SRN1, RC_DENSE1_RR, RC_DENSE1_RI, RC_HAD):
        b = 1 << pos
        d = op["m"]
        for q in range(E):
            if q & b:
                continue
            a0, a1 = v[q].copy(), v[q | b].copy()
            if code == RC_DENSE1:
                v[q], v[q | b] = m[0] * a0 + m[1] * a1, m[2] * a0 + m[3] * a1
            elif code == RC_HAD:  # unscaled butterfly; the scale sits in a later op of the round
                v[q] = a0 + a1
                v[q | b] = v[q] - 2.0 * a1
            elif code == RC_DENSE1_RR:  # pivoted in-place form {d0, d1, d2/d0, det/d0}
                v[q] = d[0] * a0 + d[1] * a1
                v[q | b] = d[2] * v[q] + d[3] * a1
            elif code == RC_DENSE1_RI:
                v[q] = d[0] * a0 + 1j * d[1] * a1
                v[q | b] = 1j * d[2] * v[q] + d[3] * a1
            elif code == RC_DIAG1:
                if not skip & 1:
                    v[q] = m[0] * a0
                if not skip & 2:
                    v[q | b] = m[1] * a1
            elif code == RC_MONO1:
                v[q], v[q | b] = (a1, a0) if unit else (m[0] * a1, m[1] * a0)
            else:
                re = 0.5 * (a0.real + a1.real)
                v[q] = re + 1j * 0.5 * (a0.imag - a1.imag)
                v[q | b] = re + 1j * 0.5 * (-a0.imag + a1.imag)
        return
    ph_, pl_ = _POS2[pos]
    bh, bl = 1 << ph_, 1 << pl_
    if code == RC_QFT2:  # butterfly on the lower bit, controlled phase, butterfly on the higher bit
        for b in (bl, None, bh):
            if b is None:
                for q in range(E):
                    if (q & (bh | bl)) == (bh | bl):
                        v[q] = m[0] * v[q]
                continue
            for q in range(E):
                if q & b:
                    continue
                a0, a1 = v[q].copy(), v[q | b].copy()
                v[q] = a0 + a1
                v[q | b] = v[q] - 2.0 * a1
        return
    for q in range(E):
        if q & (bh | bl):
            continue
        ids = [q, q | bl, q | bh, q | bh | bl]
        a = [v[i].copy() for i in ids]
        if code == RC_DENSE2:
            for r in range(4):
                v[ids[r]] = sum(m[4 * r + c] * a[c] for c in range(4))
        elif code == RC_DENSE2_LU:  # in place: U top-down (m[0..9], row by row), then L bottom-up (m[10..15])
            at = 0
            for i in range(4):
                v[ids[i]] = m[at] * v[ids[i]]
                at += 1
                for j in range(i + 1, 4):
                    v[ids[i]] = v[ids[i]] + m[at] * v[ids[j]]
                    at += 1
            for i in (3, 2, 1):
                for j in range(i):
                    v[ids[i]] = v[ids[i]] + m[10 + i * (i - 1) // 2 + j] * v[ids[j]]
        elif code == RC_DIAG2:
            for r in range(4):
                if not (skip >> r) & 1:
                    v[ids[r]] = m[r] * a[r]
        elif code == RC_CP2:
            v[ids[3]] = m[0] * a[3]
        elif code == RC_PERM2:
            src = _PERMS[aux & 3]
            for r in range(4):
                v[ids[r]] = a[src[r]] if unit else m[r] * a[src[r]]
        else:
            raise ValueError(code)


def check_group_partition(dev: dict):
    """Every round's warps x lanes x iterations x 16 registers must cover the 2^k tile elements exactly once
    (for tiles with fewer than 2^kRegBits register bits, duplicates of the same element are allowed)."""
    _set_geom(dev)
    k = dev["k"]
    for grp in dev["groups"]:
        for ri in range(grp["first"], grp["first"] + grp["count"]):
            rd = dev["rounds"][ri]
            wt = [int(x) for x in grp["wtab"][:grp["n_warps"]]]
            allidx = []
            for w in wt:
                base = _round_items(rd, w)
                for c in range(E):
                    allidx.append(base ^ int(rd["roff"][c]))
            allidx = np.concatenate(allidx)
            uniq = np.unique(allidx)
            assert len(uniq) == 1 << k, (k, grp, rd)
            if k - (WB if grp["n_warps"] == (1 << WB) else 0) >= RB:
                assert allidx.size == 1 << k


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
                shards[r] = run_sweep(st["dev"], shards[r], np.full_like(shards[r], np.nan), rank=r)
            else:
                shards[r] = run_sweep(st["dev"], shards[r], rank=r)
    v = physical_to_logical(np.concatenate(shards), plan["end_layout"])
    return np.conj(v) if plan.get("conj_end") else v
