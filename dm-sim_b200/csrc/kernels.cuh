// dm-sim_b200/csrc/kernels.cuh -- device-side data structures and launch wrappers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmb
{
// One op of a sweep as the device sees it.  272 bytes, 16-byte aligned.
struct __align__(16) DevOp
{
    int cls;      // OpClass
    int j0, j1;   // tile-local bit positions (j0 = matrix MSB for 2-bit ops)
    int aux;      // MONO2: src[r] in bits 2r..2r+1, skip-row mask in bits 8..11; DIAG*: skip mask in bits 8..11
    double2 m[16];
};

constexpr int kMaxTileBits = 12;
constexpr int kTileThreads = 256;

// Kernel parameter block of one sweep (passed by value, lives in the constant bank).
struct SweepArgs
{
    const double2* in;
    double2* out;
    const DevOp* ops;
    int n_ops;
    int k;                        // tile bits
    int n_comp;                   // M - k
    unsigned long long n_tiles;   // 2^(M-k)
    unsigned char gin[kMaxTileBits];   // loop bit i -> physical bit when loading (ascending); smem bit = i
    unsigned char gout[kMaxTileBits];  // loop bit i -> physical bit when storing (ascending)
    unsigned char sout[kMaxTileBits];  // loop bit i -> tile-local (smem) bit when storing
    unsigned char cin[40];        // physical bits enumerated by the tile id when loading (ascending)
    unsigned char cout[40];       // ... when storing
};

struct LayoutArgs
{
    int n;                  // qubits
    int M;                  // local bits
    int rank;
    int conj;               // stored array is the complex conjugate of the state (see Plan::conj_start)
    unsigned char phys[40]; // physical bit of logical bit l
};

void launch_sweep(const SweepArgs& a, int grid, cudaStream_t s);
int sweep_max_grid(int k); // resident CTAs for tile size 2^k (SMs * occupancy)

void launch_init_state(double2* buf, size_t n_elems, bool owns_origin, cudaStream_t s);
void launch_diag(const double2* buf, const LayoutArgs& L, double* out_real, double* out_abs, cudaStream_t s);
void launch_trace(const double2* buf, const LayoutArgs& L, double* out, cudaStream_t s);   // *out must be zeroed
void launch_purity(const double2* buf, size_t n_elems, double* out, cudaStream_t s);       // *out must be zeroed
void launch_scan(const double* p, double* scan /* dim+1 */, size_t dim, cudaStream_t s);
void launch_sample(const double* scan, size_t dim, const double* r, size_t n, unsigned long long* out, cudaStream_t s);
// logical [first, first+count) of the flat col*dim+row index -> split real / imag staging
void launch_gather_split(const double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                         double* re, double* im, cudaStream_t s);
void launch_scatter_split(double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                          const double* re, const double* im, cudaStream_t s);
} // namespace dmb
