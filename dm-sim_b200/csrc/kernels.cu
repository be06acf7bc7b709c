// dm-sim_b200/csrc/kernels.cu -- measurement and layout kernels of the density-matrix engine (sm_100a).
// (the hot path, sweep_kernel, lives in sweep_kernel.cu)
//
// diag/trace/purity/scan/sample : measurement path, replaces the host loops of measure() (:521-549).
// gather/scatter_split          : layout conversion to the reference's split dm_real_res / dm_imag_res.
#include "kernels.cuh"
#include "plan.hpp"

namespace dmb
{
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
    const int grid = (int)min((size_t)device_num_sms() * 16, (n + 255) / 256);
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
    const unsigned grid = (unsigned)min((unsigned long long)device_num_sms(), (dim + 255) / 256);
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
    const unsigned grid = (unsigned)min((size_t)device_num_sms() * 8, (n + 255) / 256);
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

// owner: 0 = the shard is the whole state; 1 = write only the elements this rank owns (several ranks fill ONE staging
// buffer, possibly in a peer's memory); 2 = write zeros for the elements another rank owns (the ranks' buffers are summed)
__global__ void gather_split_kernel(const double2* __restrict__ buf, const __grid_constant__ LayoutArgs L,
                                    unsigned long long first, unsigned long long count, double* __restrict__ re,
                                    double* __restrict__ im, int owner)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long mask = (L.M >= 64) ? ~0ull : ((1ull << L.M) - 1ull);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const unsigned long long p = to_phys(first + i, L);
        const bool mine = owner == 0 || (p >> L.M) == (unsigned long long)L.rank;
        if (!mine && owner == 1) continue;
        const double2 v = mine ? buf[p & mask] : make_double2(0.0, 0.0);
        re[i] = v.x;
        im[i] = L.conj ? -v.y : v.y;
    }
}
void launch_gather_split(const double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                         double* re, double* im, int owner, cudaStream_t s)
{
    const unsigned grid = (unsigned)min((unsigned long long)device_num_sms() * 16, (count + 255) / 256);
    gather_split_kernel<<<grid, 256, 0, s>>>(buf, L, first, count, re, im, owner);
}

// arbitrary logical flat indices (col*dim + row) -> values; elements another rank owns come back as zero
__global__ void gather_elements_kernel(const double2* __restrict__ buf, const __grid_constant__ LayoutArgs L,
                                       const unsigned long long* __restrict__ idx, unsigned long long count,
                                       double* __restrict__ re, double* __restrict__ im)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long mask = (L.M >= 64) ? ~0ull : ((1ull << L.M) - 1ull);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    {
        const unsigned long long p = to_phys(idx[i], L);
        const bool mine = (p >> L.M) == (unsigned long long)L.rank;
        const double2 v = mine ? buf[p & mask] : make_double2(0.0, 0.0);
        re[i] = v.x;
        im[i] = L.conj ? -v.y : v.y;
    }
}
void launch_gather_elements(const double2* buf, const LayoutArgs& L, const unsigned long long* idx, unsigned long long count,
                            double* re, double* im, cudaStream_t s)
{
    const unsigned grid = (unsigned)min((unsigned long long)device_num_sms() * 16, (count + 255) / 256);
    gather_elements_kernel<<<grid, 256, 0, s>>>(buf, L, idx, count, re, im);
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
        if ((p >> L.M) != (unsigned long long)L.rank) continue; // another rank owns this element
        buf[p & mask] = make_double2(re[i], L.conj ? -im[i] : im[i]);
    }
}
void launch_scatter_split(double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                          const double* re, const double* im, cudaStream_t s)
{
    const unsigned grid = (unsigned)min((unsigned long long)device_num_sms() * 16, (count + 255) / 256);
    scatter_split_kernel<<<grid, 256, 0, s>>>(buf, L, first, count, re, im);
}
} // namespace dmb
