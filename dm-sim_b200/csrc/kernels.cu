// dm-sim_b200/csrc/kernels.cu -- hand-written sm_100a kernels of the density-matrix engine.
//
// sweep_kernel   : the hot path.  Replaces the reference's per-gate grid-stride loops
//                  (OP_HEAD/OP_TAIL + *_GATE bodies, src/dmsim_nvgpu_omp.cuh:989-1813) and its
//                  block_transpose (:825-855): ONE HBM pass stages 2^k complex-FP64 elements per CTA in
//                  shared memory (128-bit cp.async, >=128-byte contiguous runs), applies every fused
//                  1-/2-bit op of the sweep to the tile, and streams it back (optionally to permuted
//                  bit positions = the pack step of the qubit remap, reference packing :858-882).
// diag/trace/purity/scan/sample : measurement path, replaces the host loops of measure() (:521-549).
// gather/scatter_split          : layout conversion to the reference's split dm_real_res / dm_imag_res.
#include "kernels.cuh"
#include "plan.hpp"

namespace dmb
{
// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void st_stream(double2* p, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------------
// op bodies.  A warp owns the sub-tile selected by its group's warp bits and walks the op's work items
// (pairs / quads): item = lane + 32*iter, tile index = lane_tab[lane] ^ iter_tab[iter] ^ wpart ^ off[member]
// (every term pre-swizzled by the host encoder).  Ops, tables and matrices are read from shared memory.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const double2* op_m(const DevOp* op) { return reinterpret_cast<const double2*>(op->m); }

__device__ __forceinline__ void w_dense2(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    double2 m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = op_m(op)[i];
    const unsigned o1 = op->off[1], o2 = op->off[2], o3 = op->off[3];
    for (int it = 0; it < n_iter; it++)
    {
        const unsigned i0 = base ^ op->iter_tab[it];
        const unsigned i1 = i0 ^ o1, i2 = i0 ^ o2, i3 = i0 ^ o3;
        const double2 v0 = tile[i0], v1 = tile[i1], v2 = tile[i2], v3 = tile[i3];
        tile[i0] = cfma(m[3], v3, cfma(m[2], v2, cfma(m[1], v1, cmul(m[0], v0))));
        tile[i1] = cfma(m[7], v3, cfma(m[6], v2, cfma(m[5], v1, cmul(m[4], v0))));
        tile[i2] = cfma(m[11], v3, cfma(m[10], v2, cfma(m[9], v1, cmul(m[8], v0))));
        tile[i3] = cfma(m[15], v3, cfma(m[14], v2, cfma(m[13], v1, cmul(m[12], v0))));
    }
}

__device__ __forceinline__ void w_mono2(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    const int aux = op->aux;
    const int skip = (aux >> 8) & 15;
    const bool unit = (aux >> 12) & 1;
    unsigned off[4], soff[4];
    double2 ph[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        off[r] = op->off[r];
        soff[r] = op->off[(aux >> (2 * r)) & 3];
        ph[r] = op_m(op)[r];
    }
    for (int it = 0; it < n_iter; it++)
    {
        const unsigned x = base ^ op->iter_tab[it];
        double2 v[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (!((skip >> r) & 1)) v[r] = tile[x ^ soff[r]];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (!((skip >> r) & 1)) tile[x ^ off[r]] = unit ? v[r] : cmul(ph[r], v[r]);
    }
}

__device__ __forceinline__ void w_diag2(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    const int skip = (op->aux >> 8) & 15;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        if ((skip >> r) & 1) continue;
        const double2 d = op_m(op)[r];
        const unsigned b = base ^ op->off[r];
        for (int it = 0; it < n_iter; it++)
        {
            const unsigned i = b ^ op->iter_tab[it];
            tile[i] = cmul(d, tile[i]);
        }
    }
}

__device__ __forceinline__ void w_dense1(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    const double2 m0 = op_m(op)[0], m1 = op_m(op)[1], m2 = op_m(op)[2], m3 = op_m(op)[3];
    const unsigned o1 = op->off[1];
    for (int it = 0; it < n_iter; it++)
    {
        const unsigned i0 = base ^ op->iter_tab[it], i1 = i0 ^ o1;
        const double2 v0 = tile[i0], v1 = tile[i1];
        tile[i0] = cfma(m1, v1, cmul(m0, v0));
        tile[i1] = cfma(m3, v1, cmul(m2, v0));
    }
}

__device__ __forceinline__ void w_diag1(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    const int skip = (op->aux >> 8) & 3;
#pragma unroll
    for (int r = 0; r < 2; r++)
    {
        if ((skip >> r) & 1) continue;
        const double2 d = op_m(op)[r];
        const unsigned b = base ^ op->off[r];
        for (int it = 0; it < n_iter; it++)
        {
            const unsigned i = b ^ op->iter_tab[it];
            tile[i] = cmul(d, tile[i]);
        }
    }
}

__device__ __forceinline__ void w_mono1(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    const double2 m0 = op_m(op)[0], m1 = op_m(op)[1];
    const bool unit = (op->aux >> 12) & 1;
    const unsigned o1 = op->off[1];
    for (int it = 0; it < n_iter; it++)
    {
        const unsigned i0 = base ^ op->iter_tab[it], i1 = i0 ^ o1;
        const double2 v0 = tile[i0], v1 = tile[i1];
        tile[i0] = unit ? v1 : cmul(m0, v1);
        tile[i1] = unit ? v0 : cmul(m1, v0);
    }
}

// reference SRN_GATE (:1253-1266): re0'=re1'=(re0+re1)/2, im0'=(im0-im1)/2, im1'=(-im0+im1)/2
__device__ __forceinline__ void w_srn1(double2* tile, const DevOp* op, unsigned base, int n_iter)
{
    const unsigned o1 = op->off[1];
    for (int it = 0; it < n_iter; it++)
    {
        const unsigned i0 = base ^ op->iter_tab[it], i1 = i0 ^ o1;
        const double2 v0 = tile[i0], v1 = tile[i1];
        const double re = 0.5 * (v0.x + v1.x);
        tile[i0] = make_double2(re, 0.5 * (v0.y - v1.y));
        tile[i1] = make_double2(re, 0.5 * (-v0.y + v1.y));
    }
}

// ------------------------------------------------------------------------------------------------
// the sweep kernel: persistent CTAs, 2 per SM; shared memory = [tile | op table | group table]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads, 2) sweep_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = a.k;
    const unsigned tile_elems = 1u << k;
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    DevOp* s_ops = reinterpret_cast<DevOp*>(smem_raw + (size_t)16 * tile_elems);
    DevGroup* s_groups = reinterpret_cast<DevGroup*>(s_ops + a.n_ops);
    constexpr int NT = kTileThreads;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;

    // stage the op / group tables once per CTA (every tile runs the same program)
    {
        const int4* src = reinterpret_cast<const int4*>(a.ops);
        int4* dst = reinterpret_cast<int4*>(s_ops);
        const int n16 = a.n_ops * (int)(sizeof(DevOp) / 16);
        for (int i = t; i < n16; i += NT) dst[i] = __ldg(src + i);
        const int4* gsrc = reinterpret_cast<const int4*>(a.groups);
        int4* gdst = reinterpret_cast<int4*>(s_groups);
        const int g16 = a.n_groups * (int)(sizeof(DevGroup) / 16);
        for (int i = t; i < g16; i += NT) gdst[i] = __ldg(gsrc + i);
    }

    // per-thread part of the address maps (the low 8 loop bits come from the thread index)
    const int klo = k < 8 ? k : 8;
    const int n_it = k <= 8 ? 1 : (1 << (k - 8));
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
    __syncthreads();

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
                if (it < n_it) cp_async16(&tile[(it << 8) | s_in], src + a.hin[it]);
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- apply the sweep's ops: warp-local groups, CTA barrier only between groups ----
        for (int gi = 0; gi < a.n_groups; gi++)
        {
            const DevGroup* grp = s_groups + gi;
            if (warp < grp->n_warps)
            {
                const unsigned wpart = grp->wtab[warp];
                const int last = grp->first + grp->count;
                for (int o = grp->first; o < last; o++)
                {
                    const DevOp* op = s_ops + o;
                    if (lane < op->n_active)
                    {
                        const unsigned base = op->lane_tab[lane] ^ wpart;
                        const int n_iter = op->n_iter;
                        switch (op->cls)
                        {
                        case CLS_DENSE2: w_dense2(tile, op, base, n_iter); break;
                        case CLS_MONO2: w_mono2(tile, op, base, n_iter); break;
                        case CLS_DIAG2: w_diag2(tile, op, base, n_iter); break;
                        case CLS_DENSE1: w_dense1(tile, op, base, n_iter); break;
                        case CLS_DIAG1: w_diag1(tile, op, base, n_iter); break;
                        case CLS_MONO1: w_mono1(tile, op, base, n_iter); break;
                        case CLS_SRN1: w_srn1(tile, op, base, n_iter); break;
                        default: break;
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
        __syncthreads();
    }
}

static int g_num_sms = 0;

void sweep_setup()
{
    if (g_num_sms) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    const int max_smem = (16 << kMaxTileBits) + kMaxOpsPerSweep * (int)(sizeof(DevOp) + sizeof(DevGroup));
    cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

size_t sweep_smem_bytes(const SweepArgs& a)
{
    return ((size_t)16 << a.k) + (size_t)a.n_ops * sizeof(DevOp) + (size_t)a.n_groups * sizeof(DevGroup);
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

// ------------------------------------------------------------------------------------------------
// state init / layout conversion
// ------------------------------------------------------------------------------------------------
__global__ void init_state_kernel(double2* buf, size_t n, bool owns_origin)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        buf[i] = make_double2((i == 0 && owns_origin) ? 1.0 : 0.0, 0.0);
}
void launch_init_state(double2* buf, size_t n, bool owns_origin, cudaStream_t s)
{
    const int grid = (int)min((size_t)148 * 16, (n + 255) / 256);
    init_state_kernel<<<grid, 256, 0, s>>>(buf, n, owns_origin);
}

__device__ __forceinline__ unsigned long long to_phys(unsigned long long logical, const LayoutArgs& L)
{
    unsigned long long p = 0;
    const int N = 2 * L.n;
    for (int l = 0; l < N; l++) p |= ((logical >> l) & 1ull) << L.phys[l];
    return p;
}

__global__ void diag_kernel(const double2* __restrict__ buf, const __grid_constant__ LayoutArgs L,
                            double* __restrict__ out_real, double* __restrict__ out_abs)
{
    const unsigned long long dim = 1ull << L.n;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    const unsigned long long p = to_phys(i * dim + i, L);
    double v = 0.0;
    if ((p >> L.M) == (unsigned long long)L.rank) v = buf[p & ((1ull << L.M) - 1ull)].x;
    if (out_real) out_real[i] = v;
    if (out_abs) out_abs[i] = fabs(v);
}
void launch_diag(const double2* buf, const LayoutArgs& L, double* out_real, double* out_abs, cudaStream_t s)
{
    const unsigned long long dim = 1ull << L.n;
    diag_kernel<<<(unsigned)((dim + 255) / 256), 256, 0, s>>>(buf, L, out_real, out_abs);
}

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double warp_part[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_part[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? warp_part[threadIdx.x] : 0.0;
    if (w == 0)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v; // valid in thread 0
}

__global__ void trace_kernel(const double2* __restrict__ buf, const __grid_constant__ LayoutArgs L, double* out)
{
    const unsigned long long dim = 1ull << L.n;
    double acc = 0.0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < dim;
         i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long p = to_phys(i * dim + i, L);
        if ((p >> L.M) == (unsigned long long)L.rank) acc += buf[p & ((1ull << L.M) - 1ull)].x;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
void launch_trace(const double2* buf, const LayoutArgs& L, double* out, cudaStream_t s)
{
    const unsigned long long dim = 1ull << L.n;
    const unsigned grid = (unsigned)min((unsigned long long)148, (dim + 255) / 256);
    trace_kernel<<<grid, 256, 0, s>>>(buf, L, out);
}

__global__ void purity_kernel(const double2* __restrict__ buf, size_t n, double* out)
{
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const double2 v = __ldcs(&buf[i]);
        acc = fma(v.x, v.x, fma(v.y, v.y, acc));
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
void launch_purity(const double2* buf, size_t n, double* out, cudaStream_t s)
{
    const unsigned grid = (unsigned)min((size_t)148 * 8, (n + 255) / 256);
    purity_kernel<<<grid, 256, 0, s>>>(buf, n, out);
}

// inclusive prefix sum into scan[1..dim], scan[0] = 0.  One CTA of 1024 threads: each thread owns a
// contiguous chunk, chunk totals are scanned with warp shuffles.  dim <= 2^20.
__global__ void __launch_bounds__(1024) scan_kernel(const double* __restrict__ p, double* __restrict__ scan, size_t dim)
{
    __shared__ double warp_tot[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const size_t per = (dim + 1023) / 1024;
    const size_t lo = (size_t)t * per, hi = min(dim, lo + per);
    double sum = 0.0;
    for (size_t i = lo; i < hi; i++) sum += p[i];
    double inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const double n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        double x = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const double n = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += n;
        }
        warp_tot[lane] = x;
    }
    __syncthreads();
    double run = (inc - sum) + (w > 0 ? warp_tot[w - 1] : 0.0);
    if (t == 0) scan[0] = 0.0;
    for (size_t i = lo; i < hi; i++)
    {
        run += p[i];
        scan[i + 1] = run;
    }
}
void launch_scan(const double* p, double* scan, size_t dim, cudaStream_t s) { scan_kernel<<<1, 1024, 0, s>>>(p, scan, dim); }

// reference rule (:539-543): the j with scan[j] <= r < scan[j+1]; none -> 0
__global__ void sample_kernel(const double* __restrict__ scan, size_t dim, const double* __restrict__ r, size_t n,
                              unsigned long long* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = r[i];
    // first u in [0, dim] with scan[u] > x
    size_t lo = 0, hi = dim + 1;
    while (lo < hi)
    {
        const size_t mid = (lo + hi) >> 1;
        if (scan[mid] > x) hi = mid;
        else lo = mid + 1;
    }
    out[i] = (lo >= 1 && lo <= dim) ? (unsigned long long)(lo - 1) : 0ull;
}
void launch_sample(const double* scan, size_t dim, const double* r, size_t n, unsigned long long* out, cudaStream_t s)
{
    if (n == 0) return;
    sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(scan, dim, r, n, out);
}

__global__ void gather_split_kernel(const double2* __restrict__ buf, const __grid_constant__ LayoutArgs L,
                                    unsigned long long first, unsigned long long count, double* __restrict__ re,
                                    double* __restrict__ im)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long mask = (L.M >= 64) ? ~0ull : ((1ull << L.M) - 1ull);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const unsigned long long p = to_phys(first + i, L);
        const double2 v = buf[p & mask];
        re[i] = v.x;
        im[i] = L.conj ? -v.y : v.y;
    }
}
void launch_gather_split(const double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                         double* re, double* im, cudaStream_t s)
{
    const unsigned grid = (unsigned)min((unsigned long long)148 * 16, (count + 255) / 256);
    gather_split_kernel<<<grid, 256, 0, s>>>(buf, L, first, count, re, im);
}

__global__ void scatter_split_kernel(double2* __restrict__ buf, const __grid_constant__ LayoutArgs L,
                                     unsigned long long first, unsigned long long count, const double* __restrict__ re,
                                     const double* __restrict__ im)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long mask = (L.M >= 64) ? ~0ull : ((1ull << L.M) - 1ull);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const unsigned long long p = to_phys(first + i, L);
        buf[p & mask] = make_double2(re[i], L.conj ? -im[i] : im[i]);
    }
}
void launch_scatter_split(double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                          const double* re, const double* im, cudaStream_t s)
{
    const unsigned grid = (unsigned)min((unsigned long long)148 * 16, (count + 255) / 256);
    scatter_split_kernel<<<grid, 256, 0, s>>>(buf, L, first, count, re, im);
}
} // namespace dmb
