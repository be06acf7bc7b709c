// dm-sim_b200/csrc/sweep_kernel.cu -- the hot path (sm_100a).
//
// sweep_kernel replaces the reference's per-gate grid-stride loops (OP_HEAD/OP_TAIL + *_GATE bodies,
// src/dmsim_nvgpu_omp.cuh:989-1813), its per-gate grid.sync (:1001) and its block_transpose (:825-855):
// ONE HBM pass applies a whole fused block of 1-/2-bit ops.
//
//   * persistent CTAs of 256 threads, 2-3 resident per SM (they overlap each other's load / compute / store
//     phases); tile = 2^k complex FP64 (k <= 12, 64 KiB) staged in shared memory with 128-bit cp.async (LDGSTS)
//     into an XOR-swizzled layout, results streamed back with evict-first 128-bit stores -- optionally to
//     permuted bit positions (the pack step of the multi-GPU qubit remap, reference packing :858-882).
//     HBM runs are >= 2^low_bits * 16 B contiguous.
//   * ops run in warp-local GROUPS (each of the 8 warps owns the sub-tile selected by 3 tile bits no op of the
//     group touches, so only __syncwarp() separates ops; CTA barriers only between groups) made of register
//     ROUNDS (a lane keeps 8 elements = 3 tile bits in registers, applies every op of the round there: one
//     shared-memory round trip per round instead of one per gate).
//   * all index arithmetic is pre-computed on the host as pre-swizzled XOR tables (encode.cpp); the sweep's
//     program (ops / rounds / groups) is staged once per CTA in shared memory.
#include "kernels.cuh"

namespace dmb
{
__device__ __forceinline__ unsigned swz(unsigned e) { return e ^ ((e >> 3) & 7u); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) // a*b + c
{
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void st_stream(double2* p, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------------
// register-level op bodies.  A lane holds the 8 tile elements of its work item in v[0..7] (register index bit r
// <-> round register bit r).  P / (PH, PL) are compile-time register bits, so every v[] index is static.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const double2* op_m(const DevOp* op) { return reinterpret_cast<const double2*>(op->m); }

template <int P>
__device__ __forceinline__ void r_dense1(double2 (&v)[8], const DevOp* op)
{
    const double2 m0 = op_m(op)[0], m1 = op_m(op)[1], m2 = op_m(op)[2], m3 = op_m(op)[3];
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (!(q & (1 << P)))
        {
            const double2 a = v[q], b = v[q | (1 << P)];
            v[q] = cfma(m1, b, cmul(m0, a));
            v[q | (1 << P)] = cfma(m3, b, cmul(m2, a));
        }
}
template <int P>
__device__ __forceinline__ void r_diag1(double2 (&v)[8], const DevOp* op)
{
    const int skip = (op->aux >> 8) & 3;
    if (!(skip & 1))
    {
        const double2 d = op_m(op)[0];
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (!(q & (1 << P))) v[q] = cmul(d, v[q]);
    }
    if (!(skip & 2))
    {
        const double2 d = op_m(op)[1];
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (q & (1 << P)) v[q] = cmul(d, v[q]);
    }
}
template <int P>
__device__ __forceinline__ void r_mono1(double2 (&v)[8], const DevOp* op)
{
    const bool unit = (op->aux >> 12) & 1;
    const double2 m0 = op_m(op)[0], m1 = op_m(op)[1];
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (!(q & (1 << P)))
        {
            const double2 a = v[q], b = v[q | (1 << P)];
            v[q] = unit ? b : cmul(m0, b);
            v[q | (1 << P)] = unit ? a : cmul(m1, a);
        }
}
// reference SRN_GATE (:1253-1266): re0'=re1'=(re0+re1)/2, im0'=(im0-im1)/2, im1'=(-im0+im1)/2
template <int P>
__device__ __forceinline__ void r_srn1(double2 (&v)[8])
{
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (!(q & (1 << P)))
        {
            const double2 a = v[q], b = v[q | (1 << P)];
            const double re = 0.5 * (a.x + b.x);
            v[q] = make_double2(re, 0.5 * (a.y - b.y));
            v[q | (1 << P)] = make_double2(re, 0.5 * (-a.y + b.y));
        }
}
// register-light: the 4x4 matrix is streamed row by row from shared memory (broadcast LDS) instead of being held in
// 64 registers, so that the kernel fits 3 CTAs per SM
template <int PH, int PL>
__device__ __forceinline__ void r_dense2(double2 (&v)[8], const DevOp* op)
{
    constexpr int bh = 1 << PH, bl = 1 << PL;
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (!(q & (bh | bl)))
        {
            const double2 v0 = v[q], v1 = v[q | bl], v2 = v[q | bh], v3 = v[q | bh | bl];
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                const double2 m0 = op_m(op)[4 * r], m1 = op_m(op)[4 * r + 1], m2 = op_m(op)[4 * r + 2], m3 = op_m(op)[4 * r + 3];
                v[q | ((r & 2) ? bh : 0) | ((r & 1) ? bl : 0)] = cfma(m3, v3, cfma(m2, v2, cfma(m1, v1, cmul(m0, v0))));
            }
        }
}
template <int PH, int PL>
__device__ __forceinline__ void r_diag2(double2 (&v)[8], const DevOp* op)
{
    const int skip = (op->aux >> 8) & 15;
    constexpr int bh = 1 << PH, bl = 1 << PL;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        if ((skip >> r) & 1) continue;
        const double2 d = op_m(op)[r];
#pragma unroll
        for (int q = 0; q < 8; q++)
            if ((q & (bh | bl)) == (((r & 2) ? bh : 0) | ((r & 1) ? bl : 0))) v[q] = cmul(d, v[q]);
    }
}
// monomial ops with one of three row permutations: 0 = CX (MSB control): rows 2<->3; 1 = CX (LSB control): rows 1<->3;
// 2 = SWAP: rows 1<->2.  out[r] = ph[r] * in[src[r]].
template <int PH, int PL>
__device__ __forceinline__ void r_perm2(double2 (&v)[8], const DevOp* op)
{
    const int which = op->aux & 3;
    const bool unit = (op->aux >> 12) & 1;
    constexpr int bh = 1 << PH, bl = 1 << PL;
    const double2 p0 = op_m(op)[0], p1 = op_m(op)[1], p2 = op_m(op)[2], p3 = op_m(op)[3];
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (!(q & (bh | bl)))
        {
            double2 a0 = v[q], a1 = v[q | bl], a2 = v[q | bh], a3 = v[q | bh | bl];
            double2 t;
            if (which == 0) { t = a2; a2 = a3; a3 = t; }
            else if (which == 1) { t = a1; a1 = a3; a3 = t; }
            else { t = a1; a1 = a2; a2 = t; }
            if (!unit) { a0 = cmul(p0, a0); a1 = cmul(p1, a1); a2 = cmul(p2, a2); a3 = cmul(p3, a3); }
            v[q] = a0; v[q | bl] = a1; v[q | bh] = a2; v[q | bh | bl] = a3;
        }
}

__device__ __forceinline__ void r_diag3(double2 (&v)[8], const DevOp* op)
{
    const int skip = op->aux & 255;
#pragma unroll
    for (int c = 0; c < 8; c++)
        if (!((skip >> c) & 1)) v[c] = cmul(op_m(op)[c], v[c]);
}

#define DMB_DISPATCH1(FN, ...)                      \
    switch (op->pos)                                \
    {                                               \
    case 0: FN<0>(__VA_ARGS__); break;              \
    case 1: FN<1>(__VA_ARGS__); break;              \
    default: FN<2>(__VA_ARGS__); break;             \
    }
#define DMB_DISPATCH2(FN, ...)                      \
    switch (op->pos)                                \
    {                                               \
    case 0: FN<1, 0>(__VA_ARGS__); break;           \
    case 1: FN<2, 0>(__VA_ARGS__); break;           \
    default: FN<2, 1>(__VA_ARGS__); break;          \
    }

__device__ __forceinline__ void apply_reg_op(double2 (&v)[8], const DevOp* op)
{
    switch (op->code)
    {
    case RC_DIAG3: r_diag3(v, op); break;
    case RC_DENSE2: DMB_DISPATCH2(r_dense2, v, op); break;
    case RC_DIAG2: DMB_DISPATCH2(r_diag2, v, op); break;
    case RC_PERM2: DMB_DISPATCH2(r_perm2, v, op); break;
    case RC_DENSE1: DMB_DISPATCH1(r_dense1, v, op); break;
    case RC_DIAG1: DMB_DISPATCH1(r_diag1, v, op); break;
    case RC_MONO1: DMB_DISPATCH1(r_mono1, v, op); break;
    case RC_SRN1: DMB_DISPATCH1(r_srn1, v); break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// the sweep kernel.  Shared memory = [tile | ops | rounds | groups]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads, 3) sweep_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.k;
    const unsigned tile_elems = 1u << k;
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    DevOp* s_ops = reinterpret_cast<DevOp*>(smem_raw + (size_t)16 * tile_elems);
    DevRound* s_rounds = reinterpret_cast<DevRound*>(s_ops + a.n_ops);
    DevGroup* s_groups = reinterpret_cast<DevGroup*>(s_rounds + a.n_rounds);
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
        stage(a.ops, s_ops, a.n_ops * (int)sizeof(DevOp));
        stage(a.rounds, s_rounds, a.n_rounds * (int)sizeof(DevRound));
        stage(a.groups, s_groups, a.n_groups * (int)sizeof(DevGroup));
    }

    // per-thread part of the address maps (the low kThreadBits loop bits come from the thread index)
    const int klo = k < kThreadBits ? k : kThreadBits;
    const int n_it = k <= kThreadBits ? 1 : (1 << (k - kThreadBits));
    const bool t_active = (unsigned)t < tile_elems;
    unsigned long long g_in_lo = 0, g_out_lo = 0;
    unsigned s_out_lo = 0;
    for (int i = 0; i < klo; i++)
    {
        const unsigned long long bit = (t >> i) & 1;
        g_in_lo |= bit << a.gin[i];
        g_out_lo |= bit << a.gout[i];
        s_out_lo |= (unsigned)bit << a.sout[i];
    }
    s_out_lo = swz(s_out_lo);
    const unsigned s_in = swz((unsigned)t);
    const double2* __restrict__ gin = reinterpret_cast<const double2*>(a.in);
    double2* __restrict__ gout = reinterpret_cast<double2*>(a.out);
    __syncthreads(); // program tables visible

    for (unsigned long long tile_id = blockIdx.x; tile_id < a.n_tiles; tile_id += gridDim.x)
    {
        unsigned long long base_in = 0, base_out = 0;
        for (int i = 0; i < a.n_comp; i++)
        {
            const unsigned long long bit = (tile_id >> i) & 1ull;
            base_in |= bit << a.cin[i];
            base_out |= bit << a.cout[i];
        }
        // ---- load: 128-bit async copies straight into the swizzled tile; runs of >= 2^low_bits * 16 B ----
        if (t_active)
        {
            const double2* src = gin + (base_in | g_in_lo);
#pragma unroll
            for (int it = 0; it < 16; it++)
                if (it < n_it) cp_async16(&tile[(it << kThreadBits) | s_in], src + a.hin[it]);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();

        // ---- apply the sweep's ops: warp-local groups (CTA barrier only between groups) of register rounds
        //      (one shared-memory round trip per round, all its ops applied in registers) ----
        for (int gi = 0; gi < a.n_groups; gi++)
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
                        const DevOp* ops = s_ops + rd->first;
                        const int n_ops = rd->count;
                        unsigned roff[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) roff[c] = rd->roff[c];
                        for (int it = 0; it < n_iter; it++)
                        {
                            const unsigned base = lbase ^ rd->iter_tab[it];
                            double2 v[8];
#pragma unroll
                            for (int c = 0; c < 8; c++) v[c] = tile[base ^ roff[c]];
                            for (int o = 0; o < n_ops; o++) apply_reg_op(v, ops + o);
#pragma unroll
                            for (int c = 0; c < 8; c++) tile[base ^ roff[c]] = v[c];
                        }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }

        // ---- store (streaming, evict-first) ----
        if (t_active)
        {
            double2* dst = gout + (base_out | g_out_lo);
#pragma unroll
            for (int it = 0; it < 16; it++)
                if (it < n_it) st_stream(dst + a.hout[it], tile[s_out_lo ^ a.hs[it]]);
        }
        __syncthreads(); // every thread is done with the tile before the next load overwrites it
    }
}

static int g_num_sms = 0;

void sweep_setup()
{
    if (g_num_sms) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    const int max_smem = (16 << kMaxTileBits) + kMaxOpsPerSweep * (int)(sizeof(DevOp) + sizeof(DevRound) + sizeof(DevGroup));
    cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

size_t sweep_smem_bytes(const SweepArgs& a)
{
    return ((size_t)16 << a.k) + (size_t)a.n_ops * sizeof(DevOp) + (size_t)a.n_rounds * sizeof(DevRound) +
           (size_t)a.n_groups * sizeof(DevGroup);
}

int sweep_max_grid(const SweepArgs& a)
{
    sweep_setup();
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel, kTileThreads, sweep_smem_bytes(a));
    if (occ < 1) occ = 1;
    return g_num_sms * occ;
}

void launch_sweep(const SweepArgs& a, int grid, cudaStream_t s)
{
    sweep_kernel<<<grid, kTileThreads, sweep_smem_bytes(a), s>>>(a);
}
} // namespace dmb
