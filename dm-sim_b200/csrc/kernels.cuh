// dm-sim_b200/csrc/kernels.cuh -- device-side data structures and launch wrappers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devop.hpp"

namespace dmb
{
struct LayoutArgs
{
    int n;                  // qubits
    int M;                  // local bits
    int rank;
    int conj;               // stored array is the complex conjugate of the state (see Plan::conj_start)
    unsigned char phys[40]; // physical bit of logical bit l
};

size_t sweep_smem_bytes(const SweepArgs& a);
cudaError_t launch_sweep(const SweepArgs& a, int grid, cudaStream_t s);
void set_sweep_dual(bool on);           // experiments: full-size tiles on the two-iterations-resident kernel
// every sweep of a small-state run in ONE cooperative launch (grid barrier between sweeps)
cudaError_t launch_multi_sweep(const SweepArgs* d_list, int n, unsigned op_mask_union, size_t max_body_smem, unsigned long long max_tiles,
                               cudaStream_t s, int* grid_out);
int device_num_sms();                   // SM count of the current device
int sweep_max_grid(const SweepArgs& a); // resident CTAs for this sweep's shared-memory footprint (SMs * occupancy)
void sweep_setup();                     // one-time function attributes (must not run inside a stream capture)
int sweep_smem_limit();                 // dynamic shared-memory limit the sweep kernels are set up with on the current device
// launch on a run-time specialised kernel of the sweep (jit.cu)
int sweep_max_grid_fn(const void* fn, const SweepArgs& a);
cudaError_t launch_sweep_fn(const void* fn, const SweepArgs& a, int grid, cudaStream_t s);

void launch_init_state(double2* buf, size_t n_elems, bool owns_origin, cudaStream_t s);
void launch_diag(const double2* buf, const LayoutArgs& L, double* out_real, double* out_abs, cudaStream_t s);
void launch_trace(const double2* buf, const LayoutArgs& L, double* out, cudaStream_t s);   // *out must be zeroed
void launch_purity(const double2* buf, size_t n_elems, double* out, cudaStream_t s);       // *out must be zeroed
void launch_scan(const double* p, double* scan /* dim+1 */, size_t dim, cudaStream_t s);
void launch_sample(const double* scan, size_t dim, const double* r, size_t n, unsigned long long* out, cudaStream_t s);
// logical [first, first+count) of the flat col*dim+row index -> split real / imag staging
// owner: 0 whole state in this shard, 1 write owned elements only, 2 zero the elements of other ranks
void launch_gather_split(const double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                         double* re, double* im, int owner, cudaStream_t s);
void launch_gather_elements(const double2* buf, const LayoutArgs& L, const unsigned long long* idx, unsigned long long count,
                            double* re, double* im, cudaStream_t s);
void launch_scatter_split(double2* buf, const LayoutArgs& L, unsigned long long first, unsigned long long count,
                          const double* re, const double* im, cudaStream_t s);
} // namespace dmb
