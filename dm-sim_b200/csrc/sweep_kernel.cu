// dm-sim_b200/csrc/sweep_kernel.cu -- the hot path (sm_100a).
//
// (device code: sweep_device.cuh; this file holds the ahead-of-time instantiations and the launchers)
#include <cooperative_groups.h>

#include "kernels.cuh"
#include "sweep_device.cuh"

namespace dmb
{

template <unsigned MASK, int NI>
__global__ void __launch_bounds__(kTileThreads, NI == 2 ? 2 : 3) sweep_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    sweep_body<MASK, NI>(a, smem_raw);
}

// Small states (every sweep works on tiles smaller than the full size: the whole state is a few MiB and stays in L2): ALL
// sweeps of a run in ONE cooperative launch, a grid barrier between sweeps instead of a kernel launch -- what the
// reference does with grid.sync() after every gate (:1001), here after every fused sweep.  The parameter block of the
// current sweep is staged behind the largest tile + program footprint of the run (`args_off`).
template <unsigned MASK>
__global__ void __launch_bounds__(kTileThreads, 3) multi_sweep_kernel(const SweepArgs* __restrict__ list, int n_sweeps, int args_off)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    SweepArgs* cur = reinterpret_cast<SweepArgs*>(smem_raw + args_off);
    for (int i = 0; i < n_sweeps; i++)
    {
        const int4* src = reinterpret_cast<const int4*>(list + i);
        int4* dst = reinterpret_cast<int4*>(cur);
        for (int j = threadIdx.x; j < (int)(sizeof(SweepArgs) / 16); j += kTileThreads) dst[j] = __ldcg(src + j);
        __syncthreads();
        sweep_body<MASK, 1>(*cur, smem_raw);
        grid.sync(); // every tile of sweep i is in memory before any CTA starts sweep i + 1 (and before `cur` is overwritten)
    }
}

constexpr int kMaxDevices = 64;
static int g_num_sms[kMaxDevices] = {0}; // per device: cudaFuncSetAttribute and the SM count are per-device state

// instantiations, smallest first: a sweep runs on the first one that covers its op codes
#define BIT(c) (1u << (c))
constexpr unsigned kVariantMasks[] = {
    0u,                                                                              // pure data movement (remap pack)
    BIT(RC_DENSE2_LU),                                                               // random C2 blocks (in-place L U form)
    BIT(RC_DENSE2) | BIT(RC_DENSE2_LU),                                              // ... with badly conditioned factors too
    BIT(RC_DIAGR) | BIT(RC_DIAGP) | BIT(RC_CP2) | BIT(RC_QFT2) | BIT(RC_DENSE1_RR) | BIT(RC_HAD),                                               // H + diagonal
    BIT(RC_DIAGR) | BIT(RC_DIAGP) | BIT(RC_CP2) | BIT(RC_QFT2) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_STAR),                                // QFT-like
    BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_PERM2),                                               // H / CX
    BIT(RC_DENSE2_LU) | BIT(RC_DENSE1) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_DENSE1_RI),      // dense 1- and 2-qubit blocks
    BIT(RC_DENSE2) | BIT(RC_DENSE2_LU) | BIT(RC_DENSE1) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_DENSE1_RI),
    BIT(RC_DIAGR) | BIT(RC_DIAGP) | BIT(RC_CP2) | BIT(RC_QFT2) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_DENSE1_RI) | BIT(RC_PERM2) | BIT(RC_MONO1),               // Clifford+T style
    BIT(RC_DIAGR) | BIT(RC_DIAGP) | BIT(RC_CP2) | BIT(RC_QFT2) | BIT(RC_DENSE1) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_DENSE1_RI) | BIT(RC_PERM2) | BIT(RC_MONO1) | BIT(RC_STAR), // no dense 4x4
    BIT(RC_DIAGR) | BIT(RC_DIAGP) | BIT(RC_CP2) | BIT(RC_QFT2) | BIT(RC_DENSE2_LU) | BIT(RC_DENSE1) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_DENSE1_RI) | BIT(RC_PERM2) |
        BIT(RC_MONO1) | BIT(RC_STAR),                                                // everything but SRN and the direct 4x4
    BIT(RC_DIAGR) | BIT(RC_DIAGP) | BIT(RC_CP2) | BIT(RC_QFT2) | BIT(RC_DENSE2) | BIT(RC_DENSE2_LU) | BIT(RC_DENSE1) | BIT(RC_DENSE1_RR) | BIT(RC_HAD) | BIT(RC_DENSE1_RI) |
        BIT(RC_PERM2) | BIT(RC_MONO1) | BIT(RC_SRN1) | BIT(RC_STAR),                 // everything
};
constexpr int kNumVariants = sizeof(kVariantMasks) / sizeof(kVariantMasks[0]);
typedef void (*SweepFn)(const SweepArgs);
typedef void (*MultiSweepFn)(const SweepArgs*, int, int);
static MultiSweepFn g_multi[kNumVariants];
template <int I>
struct VariantTable
{
    static void fill(SweepFn* t, SweepFn* t2)
    {
        g_multi[I] = multi_sweep_kernel<kVariantMasks[I]>;
        t[I] = sweep_kernel<kVariantMasks[I], 1>;
        t2[I] = sweep_kernel<kVariantMasks[I], 1>; // (the NI = 2 kernels are not built: measured slower, profiles/README.md)
        VariantTable<I - 1>::fill(t, t2);
    }
};
template <>
struct VariantTable<-1>
{
    static void fill(SweepFn*, SweepFn*) {}
};
static SweepFn g_variants[kNumVariants], g_variants2[kNumVariants];
static bool g_dual = false; // full-size tiles on the dual kernel (two resident iterations, two CTAs per SM)
void set_sweep_dual(bool on) { g_dual = on; }
static SweepFn pick_kernel(const SweepArgs& a);

static int pick_variant(unsigned mask)
{
    for (int i = 0; i < kNumVariants; i++)
        if (!(mask & ~kVariantMasks[i])) return i;
    return kNumVariants - 1;
}

static int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}

static SweepFn pick_kernel(const SweepArgs& a)
{
    const int i = pick_variant(a.op_mask);
    return (g_dual && a.k == kMaxTileBits) ? g_variants2[i] : g_variants[i];
}

void sweep_setup()
{
    const int dev = current_device();
    if (g_num_sms[dev]) return;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    VariantTable<kNumVariants - 1>::fill(g_variants, g_variants2);
    const int max_smem = (16 << kMaxTileBits) + 16 + kMaxOpsPerSweep * (int)(sizeof(DevOp) + sizeof(DevRound) + sizeof(DevGroup)) +
                         kMaxStarsPerSweep * kStarSmemBytes;
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int smem_limit = optin > 0 && optin < max_smem ? optin : max_smem;
    for (int i = 0; i < kNumVariants; i++)
    {
        cudaFuncSetAttribute(g_variants[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit);
        cudaFuncSetAttribute(g_variants2[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit);
        cudaFuncSetAttribute(g_multi[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit);
    }
    g_num_sms[dev] = sms > 0 ? sms : 1;
}

int sweep_smem_limit()
{
    const int max_smem = (16 << kMaxTileBits) + 16 + kMaxOpsPerSweep * (int)(sizeof(DevOp) + sizeof(DevRound) + sizeof(DevGroup)) +
                         kMaxStarsPerSweep * kStarSmemBytes;
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, current_device());
    return optin > 0 && optin < max_smem ? optin : max_smem;
}

int device_num_sms()
{
    sweep_setup();
    return g_num_sms[current_device()];
}

size_t sweep_smem_bytes(const SweepArgs& a)
{
    return ((size_t)16 << a.k) + 16 + (size_t)a.ops_bytes + (size_t)a.n_rounds * sizeof(DevRound) +
           (size_t)a.n_groups * sizeof(DevGroup) + (size_t)a.n_stars * kStarSmemBytes;
}

// all sweeps of `list` (device copy of n SweepArgs, none of them a TMA / full-size sweep) in one cooperative launch
cudaError_t launch_multi_sweep(const SweepArgs* d_list, int n, unsigned op_mask_union, size_t max_body_smem, unsigned long long max_tiles,
                               cudaStream_t s, int* grid_out)
{
    sweep_setup();
    static_assert(sizeof(SweepArgs) % 16 == 0, "SweepArgs is copied in 16-byte pieces");
    const int v = pick_variant(op_mask_union);
    int args_off = (int)((max_body_smem + 127) / 128 * 128);
    const size_t smem = (size_t)args_off + sizeof(SweepArgs);
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, g_multi[v], kTileThreads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    int grid = g_num_sms[current_device()] * occ; // a cooperative grid must be resident as a whole
    if ((unsigned long long)grid > max_tiles) grid = (int)max_tiles;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = grid;
    void* params[] = {(void*)&d_list, (void*)&n, (void*)&args_off};
    return cudaLaunchCooperativeKernel((const void*)g_multi[v], dim3(grid), dim3(kTileThreads), params, smem, s);
}

int sweep_max_grid(const SweepArgs& a)
{
    sweep_setup();
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_kernel(a), kTileThreads, sweep_smem_bytes(a));
    if (occ < 1) occ = 1;
    return g_num_sms[current_device()] * occ;
}

cudaError_t launch_sweep(const SweepArgs& a, int grid, cudaStream_t s)
{
    sweep_setup();
    void* params[] = {const_cast<SweepArgs*>(&a)};
    return cudaLaunchKernel((const void*)pick_kernel(a), dim3(grid), dim3(kTileThreads), params, sweep_smem_bytes(a), s);
}

// the same launch on a run-time specialised kernel of this sweep (jit.cu): same parameter block, same shared-memory layout
int sweep_max_grid_fn(const void* fn, const SweepArgs& a)
{
    sweep_setup();
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kTileThreads, sweep_smem_bytes(a));
    if (occ < 1) occ = 1;
    return g_num_sms[current_device()] * occ;
}
cudaError_t launch_sweep_fn(const void* fn, const SweepArgs& a, int grid, cudaStream_t s)
{
    void* params[] = {const_cast<SweepArgs*>(&a)};
    return cudaLaunchKernel(fn, dim3(grid), dim3(kTileThreads), params, sweep_smem_bytes(a), s);
}
} // namespace dmb
