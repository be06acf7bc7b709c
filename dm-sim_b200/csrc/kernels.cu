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

__device__ __forceinline__ unsigned insert0(unsigned x, int pos)
{
    const unsigned low = x & ((1u << pos) - 1u);
    return ((x >> pos) << (pos + 1)) | low;
}

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
// op bodies: every thread of the CTA walks the pairs / quads of the tile
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void op_dense2(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j0 = op->j0, j1 = op->j1;
    const int lo = min(j0, j1), hi = max(j0, j1);
    const unsigned b0 = 1u << j0, b1 = 1u << j1;
    double2 m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = __ldg(&op->m[i]);
    const unsigned nquads = 1u << (k - 2);
    for (unsigned g = t; g < nquads; g += NT)
    {
        const unsigned x = insert0(insert0(g, lo), hi);
        const unsigned i0 = swz(x), i1 = swz(x | b1), i2 = swz(x | b0), i3 = swz(x | b0 | b1);
        const double2 v0 = tile[i0], v1 = tile[i1], v2 = tile[i2], v3 = tile[i3];
        double2 o0 = cfma(m[3], v3, cfma(m[2], v2, cfma(m[1], v1, cmul(m[0], v0))));
        double2 o1 = cfma(m[7], v3, cfma(m[6], v2, cfma(m[5], v1, cmul(m[4], v0))));
        double2 o2 = cfma(m[11], v3, cfma(m[10], v2, cfma(m[9], v1, cmul(m[8], v0))));
        double2 o3 = cfma(m[15], v3, cfma(m[14], v2, cfma(m[13], v1, cmul(m[12], v0))));
        tile[i0] = o0; tile[i1] = o1; tile[i2] = o2; tile[i3] = o3;
    }
}

template <int NT>
__device__ __forceinline__ void op_mono2(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j0 = op->j0, j1 = op->j1, aux = op->aux;
    const int lo = min(j0, j1), hi = max(j0, j1);
    const unsigned b0 = 1u << j0, b1 = 1u << j1;
    unsigned off[4], soff[4];
    double2 ph[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        off[r] = ((r & 2) ? b0 : 0u) | ((r & 1) ? b1 : 0u);
        const int s = (aux >> (2 * r)) & 3;
        soff[r] = ((s & 2) ? b0 : 0u) | ((s & 1) ? b1 : 0u);
        ph[r] = __ldg(&op->m[r]);
    }
    const int skip = (aux >> 8) & 15;
    const bool unit = (aux >> 12) & 1;
    const unsigned nquads = 1u << (k - 2);
    for (unsigned g = t; g < nquads; g += NT)
    {
        const unsigned x = insert0(insert0(g, lo), hi);
        double2 v[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (!((skip >> r) & 1)) v[r] = tile[swz(x | soff[r])];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (!((skip >> r) & 1)) tile[swz(x | off[r])] = unit ? v[r] : cmul(ph[r], v[r]);
    }
}

template <int NT>
__device__ __forceinline__ void op_diag2(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j0 = op->j0, j1 = op->j1, aux = op->aux;
    const int lo = min(j0, j1), hi = max(j0, j1);
    const unsigned b0 = 1u << j0, b1 = 1u << j1;
    const int skip = (aux >> 8) & 15;
    const unsigned nquads = 1u << (k - 2);
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        if ((skip >> r) & 1) continue;
        const double2 d = __ldg(&op->m[r]);
        const unsigned o = ((r & 2) ? b0 : 0u) | ((r & 1) ? b1 : 0u);
        for (unsigned g = t; g < nquads; g += NT)
        {
            const unsigned i = swz(insert0(insert0(g, lo), hi) | o);
            tile[i] = cmul(d, tile[i]);
        }
    }
}

template <int NT>
__device__ __forceinline__ void op_dense1(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j = op->j0;
    const unsigned b = 1u << j;
    const double2 m0 = __ldg(&op->m[0]), m1 = __ldg(&op->m[1]), m2 = __ldg(&op->m[2]), m3 = __ldg(&op->m[3]);
    const unsigned npairs = 1u << (k - 1);
    for (unsigned g = t; g < npairs; g += NT)
    {
        const unsigned x = insert0(g, j);
        const unsigned i0 = swz(x), i1 = swz(x | b);
        const double2 v0 = tile[i0], v1 = tile[i1];
        tile[i0] = cfma(m1, v1, cmul(m0, v0));
        tile[i1] = cfma(m3, v1, cmul(m2, v0));
    }
}

template <int NT>
__device__ __forceinline__ void op_diag1(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j = op->j0, aux = op->aux;
    const unsigned b = 1u << j;
    const int skip = (aux >> 8) & 3;
    const unsigned npairs = 1u << (k - 1);
#pragma unroll
    for (int r = 0; r < 2; r++)
    {
        if ((skip >> r) & 1) continue;
        const double2 d = __ldg(&op->m[r]);
        for (unsigned g = t; g < npairs; g += NT)
        {
            const unsigned i = swz(insert0(g, j) | (r ? b : 0u));
            tile[i] = cmul(d, tile[i]);
        }
    }
}

template <int NT>
__device__ __forceinline__ void op_mono1(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j = op->j0;
    const unsigned b = 1u << j;
    const double2 m0 = __ldg(&op->m[0]), m1 = __ldg(&op->m[1]);
    const bool unit = (op->aux >> 12) & 1;
    const unsigned npairs = 1u << (k - 1);
    for (unsigned g = t; g < npairs; g += NT)
    {
        const unsigned x = insert0(g, j);
        const unsigned i0 = swz(x), i1 = swz(x | b);
        const double2 v0 = tile[i0], v1 = tile[i1];
        tile[i0] = unit ? v1 : cmul(m0, v1);
        tile[i1] = unit ? v0 : cmul(m1, v0);
    }
}

// reference SRN_GATE (:1253-1266): re0'=re1'=(re0+re1)/2, im0'=(im0-im1)/2, im1'=(-im0+im1)/2
template <int NT>
__device__ __forceinline__ void op_srn1(double2* tile, const DevOp* __restrict__ op, int k, int t)
{
    const int j = op->j0;
    const unsigned b = 1u << j;
    const unsigned npairs = 1u << (k - 1);
    for (unsigned g = t; g < npairs; g += NT)
    {
        const unsigned x = insert0(g, j);
        const unsigned i0 = swz(x), i1 = swz(x | b);
        const double2 v0 = tile[i0], v1 = tile[i1];
        const double re = 0.5 * (v0.x + v1.x);
        tile[i0] = make_double2(re, 0.5 * (v0.y - v1.y));
        tile[i1] = make_double2(re, 0.5 * (-v0.y + v1.y));
    }
}

// ------------------------------------------------------------------------------------------------
// the sweep kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads, 2) sweep_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);
    constexpr int NT = kTileThreads;
    const int t = threadIdx.x;
    const int k = a.k;
    const unsigned tile_elems = 1u << k;
    const int klo = k < 8 ? k : 8;

    // per-thread part of the address maps (low 8 loop bits come from the thread index)
    unsigned long long g_in_lo = 0, g_out_lo = 0;
    unsigned s_out_lo = 0;
    for (int i = 0; i < klo; i++)
    {
        const unsigned long long bit = (t >> i) & 1;
        g_in_lo |= bit << a.gin[i];
        g_out_lo |= bit << a.gout[i];
        s_out_lo |= (unsigned)bit << a.sout[i];
    }

    for (unsigned long long tile_id = blockIdx.x; tile_id < a.n_tiles; tile_id += gridDim.x)
    {
        unsigned long long base_in = 0, base_out = 0;
        for (int i = 0; i < a.n_comp; i++)
        {
            const unsigned long long bit = (tile_id >> i) & 1ull;
            base_in |= bit << a.cin[i];
            base_out |= bit << a.cout[i];
        }
        // ---- load: 128-bit async copies, >= 2^low_bits * 16 B contiguous per run ----
        for (unsigned f = t; f < tile_elems; f += NT)
        {
            unsigned long long g = base_in | g_in_lo;
            const unsigned hi = f >> 8;
            for (int i = 8; i < k; i++) g |= (unsigned long long)((hi >> (i - 8)) & 1u) << a.gin[i];
            cp_async16(&tile[swz(f)], a.in + g);
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- apply the sweep's ops on the staged tile ----
        for (int o = 0; o < a.n_ops; o++)
        {
            const DevOp* op = a.ops + o;
            switch (__ldg(&op->cls))
            {
            case CLS_DENSE2: op_dense2<NT>(tile, op, k, t); break;
            case CLS_MONO2: op_mono2<NT>(tile, op, k, t); break;
            case CLS_DIAG2: op_diag2<NT>(tile, op, k, t); break;
            case CLS_DENSE1: op_dense1<NT>(tile, op, k, t); break;
            case CLS_DIAG1: op_diag1<NT>(tile, op, k, t); break;
            case CLS_MONO1: op_mono1<NT>(tile, op, k, t); break;
            case CLS_SRN1: op_srn1<NT>(tile, op, k, t); break;
            default: break;
            }
            __syncthreads();
        }

        // ---- store ----
        for (unsigned f = t; f < tile_elems; f += NT)
        {
            unsigned long long g = base_out | g_out_lo;
            unsigned e = s_out_lo;
            const unsigned hi = f >> 8;
            for (int i = 8; i < k; i++)
            {
                const unsigned bit = (hi >> (i - 8)) & 1u;
                g |= (unsigned long long)bit << a.gout[i];
                e |= bit << a.sout[i];
            }
            st_stream(a.out + g, tile[swz(e)]);
        }
        __syncthreads();
    }
}

static int g_num_sms = 0;
static int g_grid_for_k[kMaxTileBits + 1];

int sweep_max_grid(int k)
{
    if (g_num_sms == 0)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (16 << kMaxTileBits));
        for (int i = 0; i <= kMaxTileBits; i++) g_grid_for_k[i] = 0;
    }
    if (k < 0) k = 0;
    if (k > kMaxTileBits) k = kMaxTileBits;
    if (g_grid_for_k[k] == 0)
    {
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel, kTileThreads, (size_t)16 << k);
        if (occ < 1) occ = 1;
        g_grid_for_k[k] = g_num_sms * occ;
    }
    return g_grid_for_k[k];
}

void launch_sweep(const SweepArgs& a, int grid, cudaStream_t s)
{
    sweep_kernel<<<grid, kTileThreads, (size_t)16 << a.k, s>>>(a);
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
