// dm-sim_b200/csrc/sweep_device.cuh -- device code of the sweep kernel (sm_100a): tile I/O, register-op bodies and the
// per-tile driver.  Compiled twice: ahead of time by nvcc into the INTERPRETER kernels of sweep_kernel.cu (a round's ops
// are dispatched from the device tables), and at run time by NVRTC (jit.cpp) with DMB_JIT defined: the text of this file
// is embedded in the library and the sweep's program arrives as generated straight-line code ("dmb_jit_program.inc").
//
//
// sweep_kernel replaces the reference's per-gate grid-stride loops (OP_HEAD/OP_TAIL + *_GATE bodies,
// src/dmsim_nvgpu_omp.cuh:989-1813), its per-gate grid.sync (:1001) and its block_transpose (:825-855):
// ONE HBM pass applies a whole fused block of 1-/2-bit ops.
//
//   * persistent CTAs of 128 threads, 3 resident per SM (they overlap each other's load / compute / store
//     phases); tile = 2^k complex FP64 (k <= 12, 64 KiB) staged in shared memory with 128-bit cp.async (LDGSTS)
//     into an XOR-swizzled layout, results streamed back with evict-first 128-bit stores -- optionally to
//     permuted bit positions (in place for a permutation inside the tile; the pack step of the multi-GPU qubit remap,
//     reference packing :858-882, optionally straight into the peers' shards).  HBM runs are >= 2^low_bits * 16 B.
//   * ops run in warp-local GROUPS (each of the 4 warps owns the sub-tile selected by 2 tile bits no op of the
//     group touches, so only __syncwarp() separates rounds; CTA barriers only between groups) made of register
//     ROUNDS (a lane keeps 16 elements = 4 tile bits in registers, applies every op of the round there: one
//     shared-memory round trip per round instead of one per gate).
//   * all index arithmetic is pre-computed on the host as pre-swizzled XOR tables (encode.cpp); the sweep's
//     program (op stream / rounds / groups / star tables) is staged once per CTA in shared memory; a round's ops are
//     dispatched from a packed list of jump-table indices held in registers.
#pragma once
#include "devop.hpp"

namespace dmb
{
__device__ __forceinline__ unsigned swz(unsigned e, int mode)
{
    return mode == kSwzTma ? e ^ ((e >> 3) & 7u) : e ^ ((e >> 3) & 7u) ^ ((e >> 6) & 7u) ^ ((e >> 9) & 7u);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) // a*b + c
{
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// v *= p, in place and without a temporary copy of v: both products with p.y are formed first, then each component is
// updated by ONE read-modify-write FMA (ptxas otherwise parks the new real part in a temporary and moves it back)
__device__ __forceinline__ void cmul_ip(double2& v, const double2 p)
{
    const double t1 = v.y * p.y, t2 = v.x * p.y;
    v.x = fma(v.x, p.x, -t1);
    v.y = fma(v.y, p.x, t2);
}

// a <-> b in place as three XORs per 64-bit half.  A plain register swap is "free" only inside one op body: at the
// dispatch's merge point ptxas has to restore ONE register assignment for the 16 resident elements and pays for every
// renamed register with moves (measured on bv_n15: 46 % of the executed instructions were moves).  XORs leave it nothing
// to rename.
__device__ __forceinline__ void xswap(double& a, double& b)
{
    unsigned long long x = (unsigned long long)__double_as_longlong(a), y = (unsigned long long)__double_as_longlong(b);
    asm("xor.b64 %0, %0, %1;" : "+l"(x) : "l"(y));
    asm("xor.b64 %0, %0, %1;" : "+l"(y) : "l"(x));
    asm("xor.b64 %0, %0, %1;" : "+l"(x) : "l"(y));
    a = __longlong_as_double((long long)x);
    b = __longlong_as_double((long long)y);
}
__device__ __forceinline__ void xswap(double2& a, double2& b)
{
    xswap(a.x, b.x);
    xswap(a.y, b.y);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_u32(unsigned smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- TMA tile I/O (cp.async.bulk.tensor + mbarrier; SASS: UTMALDG / UTMASTG / SYNCS) ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do
    {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const TmaDesc* map, unsigned bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
                 : "memory");
}
// the box into L2 only (the next tile of this CTA: its real load then hits L2 instead of waiting for DRAM)
__device__ __forceinline__ void tma_prefetch_5d(const TmaDesc* map, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];\n"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const TmaDesc* map, unsigned src, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// until the committed stores have been READ out of shared memory (the tile buffer may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// until the committed stores are complete (before the CTA exits)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// start coordinates of the box that holds element `full` (its box bits are zero): dimension d takes the index bits
// [start[d], start[d] + span[d]); dimension 0 counts doubles (two per element)
__device__ __forceinline__ void tma_coords(const TmaGeom& g, unsigned long long full, int (&c)[5])
{
#pragma unroll
    for (int d = 0; d < 5; d++) c[d] = (int)((full >> g.start[d]) & ((1ull << g.span[d]) - 1ull));
    c[0] <<= 1;
}

// explicit shared-state-space accesses with a 32-bit address: the tile's window offset folds into the instruction's
// immediate (through generic pointers ptxas adds the shared window base to every address: one extra instruction each)
__device__ __forceinline__ double2 lds128(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// ... with a literal byte offset folded into the instruction (generated code: the part of a register offset that no lane /
// warp / iteration contribution can touch is an ADD, not an XOR)
template <unsigned OFF>
__device__ __forceinline__ double2 lds128o(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
template <unsigned OFF>
__device__ __forceinline__ void sts128o(unsigned addr, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};\n" ::"r"(addr), "n"(OFF), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void st_stream(double2* p, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------------
// register-level op bodies.  A lane holds the 16 tile elements of its work item in v[0..15] (register index bit r
// <-> round register bit r).  P / (PH, PL) are compile-time register bits, so every v[] index is static.
// ------------------------------------------------------------------------------------------------
constexpr int E = kRegElems;
constexpr unsigned kDynSmemBase = 1024; // shared-window offset of the dynamic shared memory (no static shared memory in the
                                        // kernel; the first KiB is the system's) -- checked at kernel start
// an op in the shared-memory op stream: DevOpHdr (16 bytes) + payload
// (the kernel never reads vid / size16 from the stream: it dispatches from DevRound::vids and every body knows its size)
struct Op
{
    const unsigned char* p;
    __device__ __forceinline__ const DevOpHdr* hdr() const { return reinterpret_cast<const DevOpHdr*>(p); }
    __device__ __forceinline__ const double* m() const { return reinterpret_cast<const double*>(p + 16); }
    __device__ __forceinline__ int aux() const { return hdr()->aux; }
    __device__ __forceinline__ int star() const { return hdr()->star[0]; }
};
__device__ __forceinline__ const double2* op_m(Op op) { return reinterpret_cast<const double2*>(op.p + 16); }

template <int P>
__device__ __forceinline__ void r_dense1(double2 (&v)[E], Op op)
{
    const double2 m0 = op_m(op)[0], m1 = op_m(op)[1], m2 = op_m(op)[2], m3 = op_m(op)[3];
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (1 << P)))
        {
            const double2 a = v[q], b = v[q | (1 << P)];
            v[q] = cfma(m1, b, cmul(m0, a));
            v[q | (1 << P)] = cfma(m3, b, cmul(m2, a));
        }
}
// four real entries [[d0, d1], [d2, d3]] in pivoted IN-PLACE form (no temporaries to copy back, half the FP64
// instructions of the generic 2x2): a' = d0 a + d1 b;  b' = (d2/d0) a' + (det/d0) b.  The encoder stores
// e = {d0, d1, d2/d0, det/d0} and only emits this code when |d0| is not small.
template <int P>
__device__ __forceinline__ void r_dense1_rr(double2 (&v)[E], Op op)
{
    const double e0 = op.m()[0], e1 = op.m()[1], e2 = op.m()[2], e3 = op.m()[3];
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (1 << P)))
        {
            double2& a = v[q];
            double2& b = v[q | (1 << P)];
            a.x = fma(e0, a.x, e1 * b.x);
            a.y = fma(e0, a.y, e1 * b.y);
            b.x = fma(e2, a.x, e3 * b.x);
            b.y = fma(e2, a.y, e3 * b.y);
        }
}
// unscaled Hadamard butterfly, in place (the scale lives in another op of the round, see RC_HAD)
template <int P>
__device__ __forceinline__ void r_had(double2 (&v)[E])
{
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (1 << P)))
        {
            double2& a = v[q];
            double2& b = v[q | (1 << P)];
            a.x = a.x + b.x;
            a.y = a.y + b.y;
            b.x = fma(-2.0, b.x, a.x);
            b.y = fma(-2.0, b.y, a.y);
        }
}
// [[d0, i d1], [i d2, d3]] with real d, same in-place form: a' = d0 a + i d1 b;  b' = i (d2/d0) a' + (det/d0) b with
// det = d0 d3 + d1 d2.  e = {d0, d1, d2/d0, det/d0}
template <int P>
__device__ __forceinline__ void r_dense1_ri(double2 (&v)[E], Op op)
{
    const double e0 = op.m()[0], e1 = op.m()[1], e2 = op.m()[2], e3 = op.m()[3];
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (1 << P)))
        {
            double2& a = v[q];
            double2& b = v[q | (1 << P)];
            a.x = fma(e0, a.x, -e1 * b.y);
            a.y = fma(e0, a.y, e1 * b.x);
            const double bx = fma(-e2, a.y, e3 * b.x);
            b.y = fma(e2, a.x, e3 * b.y);
            b.x = bx;
        }
}
template <int P>
__device__ __forceinline__ void r_mono1(double2 (&v)[E], Op op, int aux)
{
    if ((aux >> 12) & 1) // unit phases (X): a register renaming
    {
#pragma unroll
        for (int q = 0; q < E; q++)
            if (!(q & (1 << P))) xswap(v[q], v[q | (1 << P)]);
        return;
    }
    const double2 m0 = op_m(op)[0], m1 = op_m(op)[1];
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (1 << P)))
        {
            const double2 a = v[q], b = v[q | (1 << P)];
            v[q] = cmul(m0, b);
            v[q | (1 << P)] = cmul(m1, a);
        }
}
// reference SRN_GATE (:1253-1266): re0'=re1'=(re0+re1)/2, im0'=(im0-im1)/2, im1'=(-im0+im1)/2
template <int P>
__device__ __forceinline__ void r_srn1(double2 (&v)[E], Op)
{
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (1 << P)))
        {
            const double2 a = v[q], b = v[q | (1 << P)];
            const double re = 0.5 * (a.x + b.x);
            v[q] = make_double2(re, 0.5 * (a.y - b.y));
            v[q | (1 << P)] = make_double2(re, 0.5 * (-a.y + b.y));
        }
}
// 4x4 dense: two quads at a time, the matrix streamed row by row from shared memory (broadcast LDS.128): 32 matrix
// loads per 256 DFMA, 32 temporaries -- the whole 4x4 in registers would not leave room for the 16 resident elements
template <int PH, int PL>
__device__ __forceinline__ void r_dense2(double2 (&v)[E], Op op)
{
    constexpr int bh = 1 << PH, bl = 1 << PL;
    constexpr int rest = (E - 1) & ~(bh | bl);   // the register bits the op does not touch (two for E = 16, one for E = 8)
    constexpr int r0 = rest & -rest, r1 = rest & ~r0;
#pragma unroll
    for (int half = 0; half < E / 8; half++)
    {
        const int qa = half ? r1 : 0, qb = qa | r0;
        const double2 a0 = v[qa], a1 = v[qa | bl], a2 = v[qa | bh], a3 = v[qa | bh | bl];
        const double2 b0 = v[qb], b1 = v[qb | bl], b2 = v[qb | bh], b3 = v[qb | bh | bl];
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const double2 m0 = op_m(op)[4 * r], m1 = op_m(op)[4 * r + 1], m2 = op_m(op)[4 * r + 2], m3 = op_m(op)[4 * r + 3];
            const int o = ((r & 2) ? bh : 0) | ((r & 1) ? bl : 0);
            v[qa | o] = cfma(m3, a3, cfma(m2, a2, cfma(m1, a1, cmul(m0, a0))));
            v[qb | o] = cfma(m3, b3, cfma(m2, b2, cfma(m1, b1, cmul(m0, b0))));
        }
    }
}
// 4x4 dense as L U, in place (RC_DENSE2_LU): every matrix entry is loaded once (broadcast LDS.128) and applied to the four
// quads of the lane (two for 8 resident elements) -- 16 loads per op, independent chains per entry, no temporaries
__device__ __forceinline__ void cfma_ip(double2& c, const double2 a, const double2 b) // c += a * b
{
    c.x = fma(a.x, b.x, fma(-a.y, b.y, c.x));
    c.y = fma(a.x, b.y, fma(a.y, b.x, c.y));
}
template <int PH, int PL>
__device__ __forceinline__ void r_dense2_lu(double2 (&v)[E], Op op)
{
    constexpr int bh = 1 << PH, bl = 1 << PL;
    constexpr int rest = (E - 1) & ~(bh | bl); // the two register bits the op does not touch
    constexpr int r0 = rest & -rest, r1 = rest & ~r0;
    const double2* m = op_m(op);
#define DMB_QUAD(qi) (((qi) & 1 ? r0 : 0) | ((qi) & 2 ? r1 : 0))
#define DMB_ELEM(i) (((i) & 1 ? bl : 0) | ((i) & 2 ? bh : 0))
    int at = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) // x_i = u_ii x_i + sum_{j > i} u_ij x_j   (x_j for j > i still hold the inputs)
    {
        const double2 d = m[at++];
#pragma unroll
        for (int q = 0; q < E / 4; q++) cmul_ip(v[DMB_QUAD(q) | DMB_ELEM(i)], d);
#pragma unroll
        for (int j = i + 1; j < 4; j++)
        {
            const double2 e = m[at++];
#pragma unroll
            for (int q = 0; q < E / 4; q++) cfma_ip(v[DMB_QUAD(q) | DMB_ELEM(i)], e, v[DMB_QUAD(q) | DMB_ELEM(j)]);
        }
    }
#pragma unroll
    for (int i = 3; i >= 1; i--) // x_i += sum_{j < i} l_ij x_j, bottom-up (x_j for j < i still hold U x)
#pragma unroll
        for (int j = 0; j < i; j++)
        {
            const double2 e = m[10 + (i * (i - 1)) / 2 + j];
#pragma unroll
            for (int q = 0; q < E / 4; q++) cfma_ip(v[DMB_QUAD(q) | DMB_ELEM(i)], e, v[DMB_QUAD(q) | DMB_ELEM(j)]);
        }
#undef DMB_QUAD
#undef DMB_ELEM
}
// monomial ops with one of three row permutations: W = 0: CX (MSB control): rows 2<->3; 1: CX (LSB control): rows
// 1<->3; 2: SWAP: rows 1<->2.  out[r] = ph[r] * in[src[r]].  With unit phases the op is a pure register renaming.
template <int PH, int PL, int W>
__device__ __forceinline__ void r_perm2w(double2 (&v)[E], Op op, int aux)
{
    constexpr int bh = 1 << PH, bl = 1 << PL;
    constexpr int x = W == 0 ? bh : (W == 1 ? bl : bl), y = W == 0 ? (bh | bl) : (W == 1 ? (bh | bl) : bh);
#pragma unroll
    for (int q = 0; q < E; q++)
        if (!(q & (bh | bl))) xswap(v[q | x], v[q | y]);
    if (!((aux >> 12) & 1))
    {
        const double2 p0 = op_m(op)[0], p1 = op_m(op)[1], p2 = op_m(op)[2], p3 = op_m(op)[3];
#pragma unroll
        for (int q = 0; q < E; q++)
            if (!(q & (bh | bl)))
            {
                cmul_ip(v[q], p0);
                cmul_ip(v[q | bl], p1);
                cmul_ip(v[q | bh], p2);
                cmul_ip(v[q | bh | bl], p3);
            }
    }
}
// one controlled phase between register bits PH and PL
template <int PH, int PL>
__device__ __forceinline__ void r_cp2(double2 (&v)[E], Op op)
{
    constexpr int both = (1 << PH) | (1 << PL);
    const double2 phi = op_m(op)[0];
#pragma unroll
    for (int c = 0; c < E; c++)
        if ((c & both) == both) cmul_ip(v[c], phi);
}
// radix-4 step of a QFT round: butterfly on PL, controlled phase between PH and PL, butterfly on PH
template <int PH, int PL>
__device__ __forceinline__ void r_qft2(double2 (&v)[E], Op op)
{
    const double2 phi = op_m(op)[0]; // (in flight while the first butterfly runs)
    r_had<PL>(v);
    constexpr int both = (1 << PH) | (1 << PL);
#pragma unroll
    for (int c = 0; c < E; c++)
        if ((c & both) == both) cmul_ip(v[c], phi);
    r_had<PH>(v);
}
template <int P>
__device__ __forceinline__ void r_mono1(double2 (&v)[E], Op op) { r_mono1<P>(v, op, op.aux()); }
template <int PH, int PL>
__device__ __forceinline__ void r_perm2(double2 (&v)[E], Op op, int aux)
{
    switch (aux & 3)
    {
    case 0: r_perm2w<PH, PL, 0>(v, op, aux); break;
    case 1: r_perm2w<PH, PL, 1>(v, op, aux); break;
    default: r_perm2w<PH, PL, 2>(v, op, aux); break;
    }
}
template <int PH, int PL>
__device__ __forceinline__ void r_perm2(double2 (&v)[E], Op op) { r_perm2<PH, PL>(v, op, op.aux()); }

__device__ __forceinline__ void r_diagr(double2 (&v)[E], Op op, int aux)
{
    const int skip = aux & 0xffff;
#pragma unroll
    for (int c = 0; c < E; c++)
        if (!((skip >> c) & 1)) cmul_ip(v[c], op_m(op)[c]);
}
__device__ __forceinline__ void r_diagr(double2 (&v)[E], Op op) { r_diagr(v, op, op.aux()); }
// diagonal whose non-unit entries all have register bit P set: 8 entries over the other three register bits
template <int P>
__device__ __forceinline__ void r_diagp(double2 (&v)[E], Op op, int aux)
{
    const int skip = aux;
#pragma unroll
    for (int j = 0; j < E / 2; j++)
    {
        const int c = ((j >> P) << (P + 1)) | (1 << P) | (j & ((1 << P) - 1));
        if (!((skip >> j) & 1)) cmul_ip(v[c], op_m(op)[j]);
    }
}

template <int P>
__device__ __forceinline__ void r_diagp(double2 (&v)[E], Op op) { r_diagp<P>(v, op, op.aux()); }

// controlled-phase star: the elements whose register bit p is set get the phase  L_p[lane] * WO_p[iw]
struct StarCtx
{
    // shared, per slot: WO[kStarW] (rebuilt per tile) | L[32] = la x lb per lane (built once); the two pointers
    // address this lane's entry of slot 0 (lane: fixed for the kernel; wo: per warp and iteration)
    const double2* lane_p;
    const double2* wo_p;
};
constexpr int kStarEntries = kStarSmemBytes / 16;
// mask = the register bits with a star; their DevStar slots are consecutive from `slot` (advanced past them)
__device__ __forceinline__ void r_star(double2 (&v)[E], int mask, int& slot, const StarCtx& sc)
{
    // two register bits at a time: their table reads first (independent addresses), then the multiplications
#pragma unroll
    for (int h = 0; h < kRegBits; h += 2)
    {
        double2 ph[2];
#pragma unroll
        for (int q = 0; q < 2; q++)
            if ((mask >> (h + q)) & 1)
            {
                ph[q] = cmul(sc.lane_p[slot * kStarEntries], sc.wo_p[slot * kStarEntries]);
                slot++;
            }
#pragma unroll
        for (int q = 0; q < 2; q++)
            if ((mask >> (h + q)) & 1)
            {
#pragma unroll
                for (int c = 0; c < E; c++)
                    if (c & (1 << (h + q))) cmul_ip(v[c], ph[q]);
            }
    }
}
// butterflies on every register bit of the mask
__device__ __forceinline__ void r_hadm(double2 (&v)[E], int mask)
{
    if (mask & 1) r_had<0>(v);
    if (mask & 2) r_had<1>(v);
    if (mask & 4) r_had<2>(v);
    if constexpr (kRegBits > 3)
        if (mask & 8) r_had<3>(v);
}

// MASK = the register-op codes compiled into this instantiation of the kernel (bit c <-> RegOpCode c).  ptxas keeps
// the 16 resident elements in ONE register assignment across the dispatch only when few bodies meet there; with all
// bodies in one kernel it copies all 64 registers before and after every op (measured: 135 moves per op).
// vid = dev_vid(code, pos, aux) (devop.hpp): one dense jump table for op kind and register position.
#define DMB_HAS(c) ((MASK >> (c)) & 1u)
#if DMB_REG_BITS > 3
#define DMB_IF4(...) __VA_ARGS__
#else
#define DMB_IF4(...)
#endif
#define DMB_SZ(c) (16 + dev_op_payload_bytes(c))
// (NI = iterations of the round a lane holds in registers at once: with NI == 2 every op is applied to both halves from
// ONE dispatch -- two independent instruction streams for ptxas to interleave)
#define DMB_DO(c, CALL0, CALL1) if (DMB_HAS(c)) { CALL0; if (NI == 2) { CALL1; } p += DMB_SZ(c); } break;
// (positions on register bit 3 exist only with 16 resident elements; the jump table keeps its numbering)
#define DMB_CASE1(base, c, FN)                                                     \
    case (base) + 0: DMB_DO(c, FN<0>(v[0], op), FN<0>(v[NI - 1], op))              \
    case (base) + 1: DMB_DO(c, FN<1>(v[0], op), FN<1>(v[NI - 1], op))              \
    case (base) + 2: DMB_DO(c, FN<2>(v[0], op), FN<2>(v[NI - 1], op))              \
    DMB_IF4(case (base) + 3: DMB_DO(c, FN<kRegBits - 1>(v[0], op), FN<kRegBits - 1>(v[NI - 1], op)))
#define DMB_CASE2(base, c, FN)                                                           \
    case (base) + 0: DMB_DO(c, (FN<1, 0>(v[0], op)), (FN<1, 0>(v[NI - 1], op)))          \
    case (base) + 1: DMB_DO(c, (FN<2, 0>(v[0], op)), (FN<2, 0>(v[NI - 1], op)))          \
    case (base) + 2: DMB_DO(c, (FN<2, 1>(v[0], op)), (FN<2, 1>(v[NI - 1], op)))          \
    DMB_IF4(case (base) + 3: DMB_DO(c, (FN<kRegBits - 1, 0>(v[0], op)), (FN<kRegBits - 1, 0>(v[NI - 1], op))))  \
    DMB_IF4(case (base) + 4: DMB_DO(c, (FN<kRegBits - 1, 1>(v[0], op)), (FN<kRegBits - 1, 1>(v[NI - 1], op))))  \
    DMB_IF4(case (base) + 5: DMB_DO(c, (FN<kRegBits - 1, kRegBits - 2>(v[0], op)), (FN<kRegBits - 1, kRegBits - 2>(v[NI - 1], op))))

// applies the op at stream position p and advances p past it (header + payload: a compile-time size per op code).
// vid = dev_vid(): ONE dense jump table; RC_HAD / RC_STAR carry their register-bit mask in the vid (no header read).
template <unsigned MASK, int NI>
__device__ __forceinline__ void apply_reg_op(double2 (&v)[NI][E], const unsigned char*& p, int vid, int& slot, const StarCtx (&sc)[NI])
{
    const Op op = {p};
    // the two mask-carrying ops first (two compares), everything else through one dense jump table
    if (DMB_HAS(RC_STAR) && vid >= kVidStar)
    {
        int slot1 = slot;
        r_star(v[0], vid - kVidStar + 1, slot, sc[0]);
        if (NI == 2) r_star(v[NI - 1], vid - kVidStar + 1, slot1, sc[NI - 1]);
        p += DMB_SZ(RC_STAR);
        return;
    }
    if (DMB_HAS(RC_HAD) && vid >= kVidHad)
    {
        r_hadm(v[0], vid - kVidHad + 1);
        if (NI == 2) r_hadm(v[NI - 1], vid - kVidHad + 1);
        p += DMB_SZ(RC_HAD);
        return;
    }
    switch (vid)
    {
        DMB_CASE2(kVidDense2, RC_DENSE2, r_dense2)
        DMB_CASE2(kVidPerm2, RC_PERM2, r_perm2)
        DMB_CASE2(kVidCp2, RC_CP2, r_cp2)
        DMB_CASE2(kVidQft2, RC_QFT2, r_qft2)
        DMB_CASE2(kVidLu2, RC_DENSE2_LU, r_dense2_lu)
        DMB_CASE1(kVidDense1, RC_DENSE1, r_dense1)
        DMB_CASE1(kVidRR, RC_DENSE1_RR, r_dense1_rr)
        DMB_CASE1(kVidRI, RC_DENSE1_RI, r_dense1_ri)
        DMB_CASE1(kVidMono1, RC_MONO1, r_mono1)
        DMB_CASE1(kVidSrn1, RC_SRN1, r_srn1)
        DMB_CASE1(kVidDiagP, RC_DIAGP, r_diagp)
    case kVidDiagR: DMB_DO(RC_DIAGR, r_diagr(v[0], op), r_diagr(v[NI - 1], op))
    default: __builtin_unreachable();
    }
}

// ------------------------------------------------------------------------------------------------
// the sweep kernel.  Shared memory = [tile | ops | rounds | groups]
// ------------------------------------------------------------------------------------------------
// NI = 1: three CTAs per SM, one iteration of a round in registers.  NI = 2 (full-size tiles only: every round has two
// iterations): two CTAs per SM with up to 255 registers, both iterations resident -- half the dispatches, two
// independent instruction streams per warp.
template <unsigned MASK, int NI>
__device__ __forceinline__ void sweep_body(const SweepArgs& a, unsigned char* smem_raw)
{
    // DMB_JIT (run-time compiled, jit.cpp): the structural parameters of the sweep are literals and its program is generated
    // straight-line code; otherwise they come from the parameter block and the program is interpreted from the device tables
#ifdef DMB_JIT
    constexpr int k = DMB_J_K, A_swz_mode = DMB_J_SWZ, A_tma_load = DMB_J_TMA_LOAD, A_tma_store = DMB_J_TMA_STORE, A_n_stars = DMB_J_N_STARS,
                  A_n_rounds = DMB_J_N_ROUNDS, A_n_groups = DMB_J_N_GROUPS, A_ops_bytes = DMB_J_OPS_BYTES, A_direct = DMB_J_DIRECT,
                  A_half_enum = DMB_J_HALF_ENUM, A_peer = DMB_J_PEER, A_tma_prefetch = DMB_J_TMA_PREFETCH;
#else
    const int k = a.k, A_swz_mode = a.swz_mode, A_tma_load = a.tma_load, A_tma_store = a.tma_store, A_n_stars = a.n_stars,
              A_n_rounds = a.n_rounds, A_n_groups = a.n_groups, A_ops_bytes = a.ops_bytes, A_direct = a.direct.enabled,
              A_half_enum = a.direct.half_enum, A_peer = a.peer_shift >= 0, A_tma_prefetch = a.tma_prefetch;
#endif
    const unsigned tile_elems = 1u << k;
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    // [tile | mbarrier (16 bytes) | op stream | rounds | groups | star tables]
    unsigned char* s_ops = smem_raw + (size_t)16 * tile_elems + 16;
    DevRound* s_rounds = reinterpret_cast<DevRound*>(s_ops + A_ops_bytes);
    DevGroup* s_groups = reinterpret_cast<DevGroup*>(s_rounds + A_n_rounds);
    double2* s_star = reinterpret_cast<double2*>(s_groups + A_n_groups); // [n_stars][WO[8] | la[8] | lb[4]]
    constexpr int NT = kTileThreads;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;

    // stage the sweep's program once per CTA (every tile runs the same program)
    {
        auto stage = [&](const void* gsrc, void* sdst, int bytes) {
            const int4* src = reinterpret_cast<const int4*>(gsrc);
            int4* dst = reinterpret_cast<int4*>(sdst);
            for (int i = t; i < bytes / 16; i += NT) dst[i] = __ldg(src + i);
        };
        stage(a.ops, s_ops, A_ops_bytes);
        stage(a.rounds, s_rounds, A_n_rounds * (int)sizeof(DevRound));
        stage(a.groups, s_groups, A_n_groups * (int)sizeof(DevGroup));
        if (DMB_HAS(RC_STAR))
            for (int i = t; i < A_n_stars * 32; i += NT) // the lane part L[lane] = la[lane & 7] * lb[lane >> 3], once per CTA
            {
                const DevStar* st = a.stars + (i >> 5);
                s_star[(i >> 5) * kStarEntries + kStarW + (i & 31)] = cmul(__ldg(reinterpret_cast<const double2*>(st->la) + (i & 7)),
                                                                      __ldg(reinterpret_cast<const double2*>(st->lb) + ((i & 31) >> 3)));
            }
    }

    // per-thread part of the address maps (the low kThreadBits loop bits come from the thread index): recomputed per
    // tile by the legacy load / store paths (a handful of instructions next to 32 copies) instead of living in registers
    // across the compute phase, where the 16 resident elements need every register
    const int klo = k < kThreadBits ? k : kThreadBits;
    const int n_it = k <= kThreadBits ? 1 : (1 << (k - kThreadBits));
    const bool t_active = (unsigned)t < tile_elems;
    const int mode = A_swz_mode;
    auto thread_in = [&]() {
        unsigned long long g = 0;
        for (int i = 0; i < klo; i++) g |= (unsigned long long)((t >> i) & 1) << a.gin[i];
        return g;
    };
    auto thread_out = [&](unsigned& s_lo) {
        unsigned long long g = 0;
        unsigned sl = 0;
        for (int i = 0; i < klo; i++)
        {
            const unsigned long long bit = (t >> i) & 1;
            g |= bit << a.gout[i];
            sl |= (unsigned)bit << a.sout[i];
        }
        s_lo = swz(sl, mode);
        return g;
    };
    const unsigned tile_u32 = (unsigned)__cvta_generic_to_shared(tile);
    // the round trips address the tile as LITERAL window offset + byte offset (the literal folds into the instruction's
    // immediate; the generic-to-shared conversion would cost an add per access)
    if (tile_u32 != kDynSmemBase) __trap();
    const unsigned bar_u32 = tile_u32 + 16u * tile_elems;
    unsigned tma_phase = 0;
    const bool direct = A_direct != 0;
    // direct store: this lane's part of the global element index of the last round's elements (lane and warp bits)
    unsigned long long d_lane = 0;
    if (direct)
    {
#pragma unroll
        for (int i = 0; i < 5; i++) d_lane |= (unsigned long long)((lane >> i) & 1) << a.direct.lane_pos[i];
#pragma unroll
        for (int i = 0; i < kWarpBits; i++) d_lane |= (unsigned long long)((warp >> i) & 1) << a.direct.warp_pos[i];
    }
    if (A_tma_load)
    {
        if (tile_u32 & 1023u) __trap(); // the hardware swizzle pattern is a function of the shared-memory ADDRESS
        // arrivals per tile: thread 0 (with the byte count of the TMA loads) + one per warp once its share of the star
        // prologue is in shared memory
        if (t == 0) mbar_init(bar_u32, 1 + NT / 32);
    }
    const double2* __restrict__ gin = reinterpret_cast<const double2*>(a.in);
    double2* __restrict__ gout = reinterpret_cast<double2*>(a.out);
    // star prologue inputs of thread t (8 lanes per star: its share of the partner phases and of the w table).  The specialised
    // kernels have registers to spare and keep them across the whole sweep when one pass of the CTA covers every star; the
    // interpreter kernels re-read them per tile (L1 hits, but a dependent global-load latency in every tile's prologue)
    constexpr int kStarQ = (kMaxStarOut + 7) / 8;
#ifdef DMB_JIT
    constexpr bool star_hoist = DMB_HAS(RC_STAR) && A_n_stars * 8 <= NT;
#else
    constexpr bool star_hoist = false;
#endif
    double2 h_wv[kStarW / 8], h_fj[kStarQ];
    int h_bj[kStarQ];
    auto star_inputs = [&](int i, bool on, double2 (&wv)[kStarW / 8], int (&bj)[kStarQ], double2 (&fj)[kStarQ]) {
        const DevStar* st = a.stars + (on ? (i >> 3) : 0);
        // (bit[] is padded with 63 and phi[] with 1 up to kMaxStarOut: every load is independent)
#pragma unroll
        for (int q = 0; q < kStarW / 8; q++) wv[q] = on ? __ldg(reinterpret_cast<const double2*>(st->w) + (i & 7) + 8 * q) : make_double2(1.0, 0.0);
#pragma unroll
        for (int q = 0; q < kStarQ; q++)
        {
            const int j = (i & 7) + 8 * q;
            bj[q] = on && j < kMaxStarOut ? __ldg(st->bit + j) : 63;
            fj[q] = on && j < kMaxStarOut ? __ldg(reinterpret_cast<const double2*>(st->phi) + j) : make_double2(1.0, 0.0);
        }
    };
    if (star_hoist) star_inputs(t, t < A_n_stars * 8, h_wv, h_bj, h_fj);
    __syncthreads(); // program tables visible

    for (unsigned long long tile_id = blockIdx.x; tile_id < a.n_tiles; tile_id += gridDim.x)
    {
        // element offset of the tile: the id's bits deposited at the positions outside the tile (three 7-bit lookups)
        const unsigned i0 = (unsigned)tile_id & 127u, i1 = (unsigned)(tile_id >> 7) & 127u, i2 = (unsigned)(tile_id >> 14) & 127u;
        const unsigned long long base_in = a.base_in[0][i0] | a.base_in[1][i1] | a.base_in[2][i2];
        const unsigned long long base_out = a.base_out[0][i0] | a.base_out[1][i1] | a.base_out[2][i2];
        // ---- load ----
        if (A_tma_load)
        {
            // TMA: one thread issues the tile's boxes (128-byte rows, hardware 128-byte swizzle); everybody waits on the
            // mbarrier after the star prologue below
            // (direct store of the last round: every tile but the CTA's first was requested during the previous tile's last round)
            if (t == 0 && (!direct || tile_id == blockIdx.x))
            {
                if (A_tma_store) tma_store_wait_read(); // the previous tile has left the buffer (nobody else waits for it)
                mbar_expect_tx(bar_u32, 16u * tile_elems);
                for (int j = 0; j < a.tma.n_copies; j++)
                {
                    int c[5];
                    tma_coords(a.tma, base_in | a.tma.enum_off[j], c);
                    tma_load_5d(tile_u32 + (unsigned)j * (unsigned)a.tma.box_bytes, &a.tmap_in, bar_u32, c[0], c[1], c[2], c[3], c[4]);
                }
                // L2 prefetch of this CTA's NEXT tile: it is consumed one tile time from now
                const unsigned long long next_id = tile_id + gridDim.x;
                if (A_tma_prefetch && next_id < a.n_tiles)
                {
                    const unsigned long long nb = a.base_in[0][(unsigned)next_id & 127u] | a.base_in[1][(unsigned)(next_id >> 7) & 127u] |
                                                  a.base_in[2][(unsigned)(next_id >> 14) & 127u];
                    for (int j = 0; j < a.tma.n_copies; j++)
                    {
                        int c[5];
                        tma_coords(a.tma, nb | a.tma.enum_off[j], c);
                        tma_prefetch_5d(&a.tmap_in, c[0], c[1], c[2], c[3], c[4]);
                    }
                }
            }
        }
        // legacy: 128-bit async copies straight into the swizzled tile; runs of >= 2^low_bits * 16 B
        else if (t_active)
        {
            const char* src = reinterpret_cast<const char*>(gin + (base_in | thread_in()));
            const unsigned s_in = swz((unsigned)t, mode);
            if (n_it == kMaxIter)
            {
                // full-size tile: no per-iteration predicates.  swz(it << kThreadBits) = (it << kThreadBits) | l3(it) with a 3-bit
                // l3, and s_in < kTileThreads: the shared address is  tile + 16 * (s_in ^ l3(it)) + 16 * (it << kThreadBits)
#pragma unroll
                for (int it = 0; it < kMaxIter; it++)
                {
                    const unsigned l3 = swz((unsigned)it << kThreadBits, mode) & 7u;
                    cp_async16_u32(tile_u32 + ((s_in ^ l3) << 4) + ((unsigned)it << (kThreadBits + 4)), src + a.hin[it]);
                }
            }
            else
            {
#pragma unroll 1 // (small tiles: a rolled loop keeps the kernel's instruction footprint down)
                for (int it = 0; it < n_it; it++) cp_async16(&tile[swz((unsigned)(it << kThreadBits), mode) ^ s_in], src + a.hin[it]);
            }
        }
        cp_async_commit();
        // controlled-phase stars: fold the partner bits OUTSIDE the tile (fixed for this tile) into the per-warp /
        // per-iteration table while the tile is in flight
        if (DMB_HAS(RC_STAR))
        {
            // WO[iw] = w[iw] * X,  X = product of phi[j] over the outside partner bits set in this tile's index: 8
            // lanes per star, each multiplies every 8th partner, then a 3-step shuffle product
            const unsigned long long full = base_in | a.rank_bits;
            for (int i0 = 0; i0 < A_n_stars * 8; i0 += NT)
            {
                const int i = i0 + t;
                const bool on = i < A_n_stars * 8;
                double2 acc = make_double2(1.0, 0.0);
                double2 wv[kStarW / 8], fj[kStarQ];
                int bj[kStarQ];
                if (star_hoist)
                {
#pragma unroll
                    for (int q = 0; q < kStarW / 8; q++) wv[q] = h_wv[q];
#pragma unroll
                    for (int q = 0; q < kStarQ; q++) { bj[q] = h_bj[q]; fj[q] = h_fj[q]; }
                }
                else star_inputs(i, on, wv, bj, fj);
                // (a lane that is not `on` holds bit 63 / phase 1 everywhere: the full index never has bit 63 set)
#pragma unroll
                for (int q = 0; q < kStarQ; q++)
                    if ((full >> bj[q]) & 1ull) acc = cmul(acc, fj[q]);
#pragma unroll
                for (int m = 1; m < 8; m <<= 1)
                {
                    const double ox = __shfl_xor_sync(0xffffffffu, acc.x, m), oy = __shfl_xor_sync(0xffffffffu, acc.y, m);
                    acc = cmul(acc, make_double2(ox, oy));
                }
                if (on)
#pragma unroll
                    for (int q = 0; q < kStarW / 8; q++) s_star[(i >> 3) * kStarEntries + (i & 7) + 8 * q] = cmul(acc, wv[q]);
            }
        }
        if (A_tma_load)
        {
            // the mbarrier completes when the tile has landed AND every warp has published its star tables
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_u32);
            mbar_wait(bar_u32, tma_phase);
            tma_phase ^= 1u;
        }
        else
        {
            cp_async_wait<0>();
            __syncthreads();
        }

        // ---- apply the sweep's ops: warp-local groups (CTA barrier only between groups) of register rounds
        //      (one shared-memory round trip per round, all its ops applied in registers) ----
        // direct store of the last round, step 1 (after the round's loads): every warp has its elements of this iteration in
        // registers: their half of the tile buffer (the whole buffer after the last iteration) is dead -- request the CTA's next
        // tile into it
        auto direct_request_next = [&](int it, bool last_it) {
            __syncthreads();
            const unsigned long long next_id = tile_id + gridDim.x;
            if (t == 0 && next_id < a.n_tiles)
            {
                const unsigned long long nb = a.base_in[0][(unsigned)next_id & 127u] | a.base_in[1][(unsigned)(next_id >> 7) & 127u] |
                                              a.base_in[2][(unsigned)(next_id >> 14) & 127u];
                const int he = A_half_enum;
                if (it == 0)
                {
                    mbar_expect_tx(bar_u32, 16u * tile_elems);
                    // L2 prefetch of the tile AFTER the next one: its load (one tile time from now) then finds it in L2 -- the
                    // last round alone is too short to cover the DRAM latency of the load it shadows
                    const unsigned long long pf_id = next_id + gridDim.x;
                    if (A_tma_prefetch && pf_id < a.n_tiles)
                    {
                        const unsigned long long pb = a.base_in[0][(unsigned)pf_id & 127u] | a.base_in[1][(unsigned)(pf_id >> 7) & 127u] |
                                                      a.base_in[2][(unsigned)(pf_id >> 14) & 127u];
                        for (int j = 0; j < a.tma.n_copies; j++)
                        {
                            int c[5];
                            tma_coords(a.tma, pb | a.tma.enum_off[j], c);
                            tma_prefetch_5d(&a.tmap_in, c[0], c[1], c[2], c[3], c[4]);
                        }
                    }
                }
                for (int j = 0; j < a.tma.n_copies; j++)
                {
                    if (he >= 0 ? ((j >> he) & 1) != it : !last_it) continue;
                    int c[5];
                    tma_coords(a.tma, nb | a.tma.enum_off[j], c);
                    tma_load_5d(tile_u32 + (unsigned)j * (unsigned)a.tma.box_bytes, &a.tmap_in, bar_u32, c[0], c[1], c[2], c[3], c[4]);
                }
            }
        };
        // ... step 2 (after the round's ops): streaming 128-bit stores straight to the state: quarter warps write whole 128-byte
        // lines (the register bits' byte offsets come from the constant bank)
        auto direct_store = [&](const double2 (&v)[E], int it) {
            const unsigned long long e0 = base_out | d_lane | a.direct.iter_off[it];
            const char* const dst = reinterpret_cast<const char*>(gout + e0);
#pragma unroll
            for (int c = 0; c < E; c++) st_stream(reinterpret_cast<double2*>(const_cast<char*>(dst) + a.direct.reg_off[c]), v[c]);
        };
#ifdef DMB_JIT
#include "dmb_jit_program.inc"
#else
        for (int gi = 0; gi < A_n_groups; gi++)
        {
            const DevGroup* grp = s_groups + gi;
            if (warp < grp->n_warps)
            {
                const unsigned wpart = grp->wtab[warp];
                const int rlast = grp->first + grp->count;
                for (int ri = grp->first; ri < rlast; ri++)
                {
                    const DevRound* rd = s_rounds + ri;
                    if (lane < rd->n_active)
                    {
                        const unsigned lbase = rd->lane_tab[lane] ^ wpart;
                        const int n_iter = rd->n_iter;
                        if (NI == 2 && (n_iter & 1)) __trap(); // (the launcher only picks the dual kernel for k == 12)
                        const unsigned char* ops = s_ops + (size_t)rd->first * 16;
                        const int n_ops = rd->count;
                        int nib = 0;
                        while ((1 << nib) < n_iter) nib++;
                        const ulonglong2 vids = *reinterpret_cast<const ulonglong2*>(rd->vids);
                        const int star0 = rd->star0;
                        // the 16 register offsets, packed two per word (kept in 8 registers: re-reading them from
                        // shared memory at store time would serialise every STS behind an LDS)
                        unsigned rw[E / 2];
                        {
                            const uint4 r0 = *reinterpret_cast<const uint4*>(rd->roff);
                            rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w;
                            if constexpr (E > 8)
                            {
                                const uint4 r1 = *(reinterpret_cast<const uint4*>(rd->roff) + 1);
                                rw[E / 2 - 4] = r1.x; rw[E / 2 - 3] = r1.y; rw[E / 2 - 2] = r1.z; rw[E / 2 - 1] = r1.w;
                            }
                        }
                        // NI iterations at a time live in registers (the dual kernel: both iterations of a full-size tile)
                        for (int it = 0; it < n_iter; it += NI)
                        {
                            unsigned base[NI];
                            double2 v[NI][E];
                            StarCtx sc[NI];
#pragma unroll
                            for (int h = 0; h < NI; h++)
                            {
                                base[h] = lbase ^ rd->iter_tab[it + h];
                                sc[h] = StarCtx{s_star + kStarW + lane, s_star + ((warp << nib) | (it + h))};
                            }
#pragma unroll
                            for (int h = 0; h < NI; h++)
#pragma unroll
                                for (int c = 0; c < E; c++)
                                    v[h][c] = lds128(kDynSmemBase + (base[h] ^ ((rw[c >> 1] >> ((c & 1) * 16)) & 0xffffu)));
                            const bool last_round = direct && ri + 1 == A_n_rounds;
                            if (last_round) direct_request_next(it, it + NI == n_iter);
                            // dispatch from the round's packed vid list (one byte per op, two registers pairs):
                            // no shared-memory load on the dispatch path
                            const unsigned char* p = ops;
                            int slot = star0;
                            unsigned long long v0 = vids.x, v1 = vids.y;
                            for (int o = 0; o < n_ops; o++)
                            {
                                const int vid = (int)(v0 & 0xffull);
                                v0 = (v0 >> 8) | (v1 << 56);
                                v1 >>= 8;
                                apply_reg_op<MASK, NI>(v, p, vid, slot, sc);
                            }
                            if (last_round) direct_store(v[0], it);
                            else
                            {
#pragma unroll
                                for (int h = 0; h < NI; h++)
#pragma unroll
                                    for (int c = 0; c < E; c++)
                                        sts128(kDynSmemBase + (base[h] ^ ((rw[c >> 1] >> ((c & 1) * 16)) & 0xffffu)), v[h][c]);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // (TMA store: the generic-proxy writes of the rounds are fenced for the async proxy before the last barrier)
            if (A_tma_store && gi + 1 == A_n_groups) fence_proxy_async();
            __syncthreads();
        }
#endif // DMB_JIT

        // ---- store ----
        if (direct) continue; // (the last round stored its results itself)
        if (A_tma_store)
        {
            // TMA: one thread issues the boxes.  Nobody waits here: thread 0 waits for the boxes to be READ out of
            // shared memory just before it issues the next tile's load, the other threads go on to the next tile's star
            // prologue (which does not touch the tile) and then block on the load's mbarrier
            if (t == 0)
            {
                for (int j = 0; j < a.tma.n_copies; j++)
                {
                    int c[5];
                    tma_coords(a.tma, base_out | a.tma.enum_off[j], c);
                    tma_store_5d(&a.tmap_out, tile_u32 + (unsigned)j * (unsigned)a.tma.box_bytes, c[0], c[1], c[2], c[3], c[4]);
                }
                tma_store_commit();
            }
            continue;
        }
        // legacy (streaming, evict-first)
        else if (t_active)
        {
            unsigned s_out_lo;
            const unsigned long long g_out_lo = thread_out(s_out_lo);
            if (!A_peer)
            {
                char* dst = reinterpret_cast<char*>(gout + (base_out | g_out_lo));
                if (n_it == kMaxIter)
                {
                    // full-size tile: batches of 8 shared-memory loads, then their 8 streaming stores
#pragma unroll
                    for (int b = 0; b < kMaxIter; b += 8)
                    {
                        double2 r[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) r[j] = tile[s_out_lo ^ a.hs[b + j]];
#pragma unroll
                        for (int j = 0; j < 8; j++) st_stream(reinterpret_cast<double2*>(dst + a.hout[b + j]), r[j]);
                    }
                }
                else
                {
#pragma unroll 1
                    for (int it = 0; it < n_it; it++) st_stream(reinterpret_cast<double2*>(dst + a.hout[it]), tile[s_out_lo ^ a.hs[it]]);
                }
            }
            else
            {
                // fused remap: every 128-byte run goes straight into its destination rank's shard (peer memory)
                const unsigned long long o_lo = base_out | g_out_lo;
                const unsigned long long low_mask = (1ull << a.peer_shift) - 1ull;
                const unsigned long long mine = (unsigned long long)a.peer_rank << a.peer_shift;
#pragma unroll 1
                for (int it = 0; it < n_it; it++)
                {
                    const unsigned long long off = o_lo | (a.hout[it] >> 4);
                    double2* dst = reinterpret_cast<double2*>(a.peer_out[off >> a.peer_shift]) + (mine | (off & low_mask));
                    st_stream(dst, tile[s_out_lo ^ a.hs[it]]);
                }
            }
        }
        __syncthreads(); // every thread is done with the tile before the next load overwrites it
    }
    if (A_tma_store && t == 0) tma_store_wait_all();
}
} // namespace dmb
