// dm-sim_b200/csrc/capi.cu -- the C-ABI (include/dmsim_b200.h): state ownership, CUDA-graph executor,
// NCCL qubit-remap exchange, result readback.  Replaces Simulation::{ctor,upload,sim,measure}
// (reference src/dmsim_nvgpu_omp.cuh:196-549).  There is no CPU fallback: every compute entry point
// fails with DMB_ECUDA when no usable GPU / driver is present.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dmsim_b200.h"
#include "encode.hpp"
#include "jit.hpp"
#include "kernels.cuh"
#include "plan.hpp"

using namespace dmb;

// ------------------------------------------------------------------------------------------------
// errors / options
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
#define CU(call)                                                                                             \
    do                                                                                                       \
    {                                                                                                        \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(DMB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +  \
                                       std::to_string(__LINE__) + ")");                                      \
    } while (0)

static PlanOptions g_opt;
static int g_use_graph = 1;
static int g_plan_cache = 1;    // reuse the plan (and the device tables) of a circuit that is set again (DMB_PLAN_CACHE=0 / option "plan_cache")
static int g_persistent = 0;    // small states: all sweeps of a run in one cooperative launch (DMB_PERSISTENT=1 / option "persistent"); measured SLOWER than the captured graph on B200 (vqe_uccsd_n8: 10.1 vs 7.8 ms, profiles/README.md): off by default
// run-time specialised sweep kernels (jit.cu): 0 off, 1 tiered (the interpreter kernel runs a sweep until its compiled kernel
// is ready), 2 wait for the compiler (DMB_JIT / option "jit"); only for shards of >= 2^jit_min_bits elements
// (DMB_JIT_MIN_BITS / option "jit_min_bits")
static int g_jit = 1;
static int g_jit_min_bits = 16;
// tiered mode: a plan's kernels are only compiled from its `jit_hot`-th run on (DMB_JIT_HOT / option "jit_hot").  A one-shot
// run, or a circuit that continues from the layout its previous run left (planned anew every time), would never get to
// use them: compiling for those only burns host cores
static int g_jit_hot = 2;
static int g_jit_hot_small = 8; // ... of shards below 2^24 elements (their sweeps take microseconds: only a real loop over one circuit pays for
                                // a compilation; DMB_JIT_HOT_SMALL / option "jit_hot_small")
// experiment for the next round (never measured: no GPU time was left): L2 promotion of the tile maps' requests (0 none, 1 / 2 /
// 3 = 64 / 128 / 256 bytes).  Sweeps whose tile has only the 3 lowest bits contiguous read 128-byte runs and reach 0.69 of the HBM
// peak where a sweep with 1-KiB runs reaches 0.86; 256-byte promotion would fetch the neighbouring tile's run with each request
static int g_tma_l2_promotion = 0;
static int g_tma_prefetch = 0;  // L2 prefetch of a CTA's next tile (DMB_TMA_PREFETCH=0 / option "tma_prefetch")
static int g_grid_per_sm = 0;   // experiments: resident CTAs per SM of the sweep kernel (0 = what the occupancy query says)
static int g_sparse_start = 1; // skip the tiles that are still all-zero after dmb_reset_dm (DMB_SPARSE=0 / option "sparse")
static bool g_opt_init = false;
static void init_options()
{
    if (g_opt_init) return;
    g_opt_init = true;
    if (const char* e = getenv("DMB_TILE_BITS")) g_opt.tile_bits = atoi(e);
    if (const char* e = getenv("DMB_LOW_BITS")) g_opt.low_bits = atoi(e);
    if (const char* e = getenv("DMB_MIN_TILES_LOG2")) g_opt.min_tiles_log2 = atoi(e);
    if (const char* e = getenv("DMB_GRAPH")) g_use_graph = atoi(e);
    if (const char* e = getenv("DMB_CPHASE")) g_opt.cphase = atoi(e) != 0;
    if (const char* e = getenv("DMB_SPARSE")) g_sparse_start = atoi(e);
    if (const char* e = getenv("DMB_TMA")) g_opt.tma = atoi(e) != 0;
    if (const char* e = getenv("DMB_TMA_BOX_BITS")) set_sweep_tma_box_bits(atoi(e));
    if (const char* e = getenv("DMB_DENSE2_LU")) set_sweep_dense2_lu(atoi(e) != 0);
    if (const char* e = getenv("DMB_GRID_PER_SM")) g_grid_per_sm = atoi(e);
    if (const char* e = getenv("DMB_DUAL")) set_sweep_dual(atoi(e) != 0);
    if (const char* e = getenv("DMB_PERSISTENT")) g_persistent = atoi(e);
    g_opt.small_state_bits = g_persistent ? 20 : 0;
    if (const char* e = getenv("DMB_JIT")) g_jit = atoi(e);
    if (const char* e = getenv("DMB_JIT_MIN_BITS")) g_jit_min_bits = atoi(e);
    if (const char* e = getenv("DMB_JIT_HOT")) g_jit_hot = atoi(e);
    if (const char* e = getenv("DMB_JIT_HOT_SMALL")) g_jit_hot_small = atoi(e);
    if (const char* e = getenv("DMB_PLAN_CACHE")) g_plan_cache = atoi(e);
    if (const char* e = getenv("DMB_DIRECT_STORE")) set_sweep_direct_store(atoi(e) != 0);
    if (const char* e = getenv("DMB_HEAVY_LAST")) set_sweep_heavy_last(atoi(e) != 0);
    if (const char* e = getenv("DMB_SPREAD_PEERS")) g_opt.spread_peers = atoi(e) != 0;
    if (const char* e = getenv("DMB_LIGHT_FIRST")) set_sweep_light_first(atoi(e) != 0);
    if (const char* e = getenv("DMB_TMA_PREFETCH")) g_tma_prefetch = atoi(e);
    if (const char* e = getenv("DMB_TMA_L2_PROMOTION")) g_tma_l2_promotion = atoi(e);
}

// ------------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched through the runtime: no link against libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried)
    {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
// the shard `buf` (2^M complex FP64) as the dense 5-D FP64 tensor TmaGeom describes; box = the tile bits of each dimension
static bool make_tile_map(const TmaGeom& g, int M, const void* buf, TmaDesc& out, std::string& why)
{
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap size");
    EncodeTiledFn enc = tma_encoder();
    if (!enc) { why = "cuTensorMapEncodeTiled is not available"; return false; }
    cuuint64_t dim[5], stride[4];
    cuuint32_t box[5], estride[5] = {1, 1, 1, 1, 1};
    for (int d = 0; d < 5; d++)
    {
        dim[d] = (cuuint64_t)1 << g.span[d];
        box[d] = (cuuint32_t)1 << g.box_log2[d];
        if (d > 0) stride[d - 1] = (cuuint64_t)16 << g.start[d];
    }
    dim[0] *= 2; // FP64 units: two per complex element
    box[0] *= 2;
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<void*>(buf), dim, stride, box, estride,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           g_tma_l2_promotion == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                           : g_tma_l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                           : g_tma_l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        why = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (M=" + std::to_string(M) + ")";
        return false;
    }
    memcpy(&out, &m, sizeof(out));
    return true;
}

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU users never need it
// ------------------------------------------------------------------------------------------------
namespace
{
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl
{
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
const int kNcclDouble = 8; // ncclFloat64 in ncclDataType_t

bool load_nccl(std::string& why)
{
    if (g_nccl.lib) return true;
    const char* names[] = {getenv("DMB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names)
    {
        if (!nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib)
    {
        why = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return false;
    }
#define SYM(field, name)                                                              \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                               \
    if (!g_nccl.field) { why = std::string("missing NCCL symbol ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
}
} // namespace

// ------------------------------------------------------------------------------------------------
// the simulation object
// ------------------------------------------------------------------------------------------------
struct dmb_sim
{
    int n = 0, g = 0, world = 1, rank = 0, device = 0;
    int N = 0, M = 0;
    size_t shard_elems = 0;
    double2* buf[2] = {nullptr, nullptr};
    int cur = 0;
    std::vector<int> layout; // physical bit of logical bit
    bool conj_flag = false;  // stored array = conj(state), see Plan::conj_start
    bool non_hermitian = false; // an SRN run or dmb_set_dm may have left a non-Hermitian state (see plan.cpp)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_comm; // pairs around exchanges

    // circuit
    std::vector<dmb_gate> gates;
    std::vector<double> mats;
    bool have_circuit = false;
    Plan plan;
    std::vector<int> plan_layout; // layout the plan was made for
    bool plan_conj = false, plan_nonherm = false;
    unsigned char* d_ops = nullptr;   // op streams of all sweeps, back to back
    size_t d_ops_cap = 0;
    DevStar* d_stars = nullptr;
    size_t d_stars_cap = 0;
    std::vector<size_t> op_offset;    // per step: byte offset of its op stream
    std::vector<size_t> star_offset;  // per step: first DevStar
    std::vector<int> n_dev_stars;
    std::vector<size_t> round_offset, group_offset; // per step: first DevRound / DevGroup
    std::vector<int> n_dev_ops, n_dev_rounds, n_dev_groups;
    std::vector<unsigned> op_masks;   // per step: register-op codes present (kernel instantiation)
    std::vector<DevDirect> directs;   // per step: direct store of the last round
    DevRound* d_rounds = nullptr;
    DevGroup* d_groups = nullptr;
    size_t d_rounds_cap = 0, d_groups_cap = 0;
    cudaGraphExec_t graph_exec = nullptr;
    int graph_cur = -1;
    // small states: the parameter blocks of all sweeps of the run, for the one-launch cooperative executor
    SweepArgs* d_multi = nullptr;
    size_t d_multi_cap = 0;
    bool multi_valid = false;
    int multi_cur = -1, multi_end_cur = 0, multi_n = 0;
    unsigned long long multi_support = ~0ull, multi_max_tiles = 1;
    unsigned multi_mask = 0;
    size_t multi_smem = 0;
    // Sparse start: every element whose shard index has a 1 in a physical bit OUTSIDE `support` is exactly zero (after
    // dmb_reset_dm only element 0 is non-zero: support = 0).  A sweep maps each tile onto itself, so tiles with such a
    // bit set stay zero and are not launched at all; the sweep adds its tile bits to the support.  Single GPU only.
    unsigned long long support = ~0ull;
    unsigned long long graph_support = ~0ull; // support at the start of the captured run
    uint64_t h2d_bytes = 0;

    // scratch for results
    double* d_scratch = nullptr;
    size_t d_scratch_bytes = 0;

    ncclComm_t comm = nullptr;
    // peer-memory exchange: IPC-mapped shard buffers of every rank (peer[b][r] = rank r's buf[b]); see dmb_comm_import
    bool p2p = false;
    double2* peer[2][8] = {{nullptr}};
    double* d_barrier = nullptr; // 2 doubles: all-reduce source / sink used as the cross-GPU barrier

    // host copies of the device tables of the current plan (a group shares ONE plan between its shards)
    std::vector<unsigned char> host_ops;
    std::vector<DevStar> host_stars;
    std::vector<DevRound> host_rounds;
    std::vector<DevGroup> host_groups;
    unsigned long long fp64_per_lane = 0; // FP64 instructions per kRegElems shard elements over all sweeps of the plan

    // run-time specialised kernels of the current plan's sweeps: (step << 4 | I/O variant) -> cache entry (nullptr: the
    // generator does not cover the sweep); jit_mode: 0 resolve and launch, 1 dry pass that only queues the compilations,
    // 2 dry pass that resolves (waits / loads) without launching, 3 launching inside a stream capture (nothing may be loaded)
    std::unordered_map<unsigned long long, JitKernel*> jit_memo;
    int jit_mode = 0;
    unsigned plan_runs = 0;                      // dmb_run calls on the device tables of the current plan (tiered mode: hotness)
    unsigned jit_missing = 0, jit_used = 0;      // sweeps of the last enqueue that ran interpreted because their kernel was not ready / specialised
    unsigned graph_jit_missing = 0;              // ... of the captured graph
    unsigned long long graph_jit_epoch = 0;      // jit_ready_count() when the graph was captured

    // plan cache: the host pipeline (expand / fuse / schedule / encode) is skipped when the same circuit is set again on
    // the same layout (repeated dmb_set_circuit of one circuit, runs that alternate between a few circuits); when the
    // device still holds that plan's tables the upload and the captured graph survive as well
    struct CachedPlan
    {
        unsigned long long key[2] = {0, 0};
        Plan plan;
        std::vector<int> plan_layout;
        bool plan_conj = false, plan_nonherm = false;
        std::vector<size_t> op_offset, star_offset, round_offset, group_offset;
        std::vector<int> n_dev_stars, n_dev_ops, n_dev_rounds, n_dev_groups;
        std::vector<unsigned> op_masks;
        std::vector<DevDirect> directs;
        std::vector<unsigned char> host_ops;
        std::vector<DevStar> host_stars;
        std::vector<DevRound> host_rounds;
        std::vector<DevGroup> host_groups;
        unsigned long long fp64_per_lane = 0;
    };
    std::vector<CachedPlan> plan_cache; // most recently used first, at most kPlanCacheEntries
    unsigned long long device_key[2] = {0, 0}; // key of the plan whose tables are on the device (0, 0: none)
    unsigned long long plan_key[2] = {0, 0};   // key of the current plan

    // Single-process multi-GPU (reference Simulation(n_qubits, n_gpus), :196-271: one host process drives all devices):
    // a GROUP handle (rank == DMB_ALL_RANKS) owns one shard object per device and no buffers of its own.  The shards
    // see each other's buffers through cudaDeviceEnablePeerAccess (no IPC), the cross-GPU barrier of the remap is a
    // pair of CUDA events per shard (no NCCL).
    std::vector<dmb_sim*> shards;
    bool in_group = false;
    cudaEvent_t ev_pre = nullptr, ev_post = nullptr;
};

static LayoutArgs layout_args(const dmb_sim* s)
{
    LayoutArgs L;
    memset(&L, 0, sizeof(L));
    L.n = s->n; L.M = s->M; L.rank = s->rank; L.conj = s->conj_flag ? 1 : 0;
    for (int l = 0; l < s->N; l++) L.phys[l] = (unsigned char)s->layout[l];
    return L;
}

static int ensure_scratch(dmb_sim* s, size_t bytes)
{
    if (s->d_scratch_bytes >= bytes) return DMB_OK;
    if (s->d_scratch) cudaFree(s->d_scratch);
    s->d_scratch = nullptr; s->d_scratch_bytes = 0;
    CU(cudaMalloc(&s->d_scratch, bytes));
    s->d_scratch_bytes = bytes;
    return DMB_OK;
}

static int ensure_second_buffer(dmb_sim* s)
{
    if (s->buf[1]) return DMB_OK;
    CU(cudaMalloc(&s->buf[1], s->shard_elems * sizeof(double2)));
    return DMB_OK;
}

static void drop_graph(dmb_sim* s)
{
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    s->graph_exec = nullptr;
    s->graph_cur = -1;
    s->multi_valid = false;
}


// ------------------------------------------------------------------------------------------------
// helpers shared by the single-shard and the group forms
// ------------------------------------------------------------------------------------------------
static bool is_group(const dmb_sim* s) { return !s->shards.empty(); }

constexpr size_t kPlanCacheEntries = 4;
static unsigned long long g_option_epoch = 1; // bumped by dmb_set_option: plans made under other options are not reused
static void fnv(unsigned long long (&h)[2], const void* data, size_t bytes)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; i++)
    {
        h[0] = (h[0] ^ p[i]) * 1099511628211ull;
        h[1] = (h[1] + p[i] + 0x9e3779b97f4a7c15ull) * 0xff51afd7ed558ccdull;
        h[1] ^= h[1] >> 29;
    }
}
static void plan_key_of(const dmb_sim* s, unsigned long long (&h)[2])
{
    h[0] = 1469598103934665603ull; h[1] = 0x2545f4914f6cdd1dull;
    const unsigned long long hdr[6] = {(unsigned long long)s->n, (unsigned long long)s->world, s->gates.size(), s->mats.size(),
                                       (unsigned long long)(s->conj_flag ? 1 : 0) | (s->non_hermitian ? 2 : 0), g_option_epoch};
    fnv(h, hdr, sizeof(hdr));
    if (!s->gates.empty()) fnv(h, s->gates.data(), s->gates.size() * sizeof(dmb_gate));
    if (!s->mats.empty()) fnv(h, s->mats.data(), s->mats.size() * sizeof(double));
    if (!s->layout.empty()) fnv(h, s->layout.data(), s->layout.size() * sizeof(int));
    if (!h[0] && !h[1]) h[0] = 1;
}

// cross-rank sum of `count` doubles in place on s->stream (one-process-per-GPU form, communicator attached)
static int all_reduce_sum(dmb_sim* s, double* d_buf, size_t count)
{
    const int nr = g_nccl.AllReduce(d_buf, d_buf, count, kNcclDouble, 0 /* ncclSum */, s->comm, s->stream);
    if (nr != 0) return fail(DMB_ECOMM, std::string("NCCL all-reduce failed: ") + g_nccl.GetErrorString(nr));
    return DMB_OK;
}
// results of a sharded state are GLOBAL (every rank gets the full answer) when the ranks can talk to each other
static bool collective(const dmb_sim* s) { return s->world > 1 && !s->in_group && s->comm != nullptr; }

static int create_shard(int n_qubits, int world_size, int g, int rank, int device, dmb_sim** out)
{
    CU(cudaSetDevice(device));
    dmb_sim* s = new dmb_sim;
    s->n = n_qubits; s->g = g; s->world = world_size; s->rank = rank; s->device = device;
    s->N = 2 * n_qubits; s->M = s->N - g;
    s->shard_elems = (size_t)1 << s->M;
    cudaError_t ea = cudaMalloc(&s->buf[0], s->shard_elems * sizeof(double2));
    if (ea != cudaSuccess)
    {
        delete s;
        return fail(DMB_ENOMEM, std::string("cudaMalloc of the state shard failed: ") + cudaGetErrorString(ea));
    }
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev_begin));
    CU(cudaEventCreate(&s->ev_end));
    s->layout.resize(s->N);
    *out = s;
    return DMB_OK;
}

static int destroy_shard(dmb_sim* s)
{
    cudaSetDevice(s->device);
    drop_graph(s);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    if (s->p2p && !s->in_group)
        for (int b = 0; b < 2; b++)
            for (int r = 0; r < s->world; r++)
                if (r != s->rank && s->peer[b][r]) cudaIpcCloseMemHandle(s->peer[b][r]);
    if (s->d_barrier) cudaFree(s->d_barrier);
    for (auto ev : s->ev_comm) cudaEventDestroy(ev);
    if (s->ev_begin) cudaEventDestroy(s->ev_begin);
    if (s->ev_end) cudaEventDestroy(s->ev_end);
    if (s->ev_pre) cudaEventDestroy(s->ev_pre);
    if (s->ev_post) cudaEventDestroy(s->ev_post);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->buf[0]) cudaFree(s->buf[0]);
    if (s->buf[1]) cudaFree(s->buf[1]);
    if (s->d_ops) cudaFree(s->d_ops);
    if (s->d_stars) cudaFree(s->d_stars);
    if (s->d_rounds) cudaFree(s->d_rounds);
    if (s->d_groups) cudaFree(s->d_groups);
    if (s->d_scratch) cudaFree(s->d_scratch);
    if (s->d_multi) cudaFree(s->d_multi);
    delete s;
    return DMB_OK;
}

static int reset_shard(dmb_sim* s)
{
    CU(cudaSetDevice(s->device));
    for (int l = 0; l < s->N; l++) s->layout[l] = l;
    s->conj_flag = false;
    s->non_hermitian = false;
    // (the current buffer index is kept: with peer-memory remaps every rank must agree on it, and a rank that has not
    // reached this reset yet may still be read or written through the other buffer)
    s->support = (s->world == 1 && g_sparse_start) ? 0ull : ~0ull;
    launch_init_state(s->buf[s->cur], s->shard_elems, s->rank == 0, s->stream);
    CU(cudaGetLastError());
    // no synchronisation: everything that touches the state is ordered on s->stream, so the 16 B/element clear overlaps
    // the host-side planning of the next dmb_set_circuit
    return DMB_OK;
}

template <typename T>
static int grow_device(T*& ptr, size_t& cap, size_t need)
{
    if (need <= cap) return DMB_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    CU(cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
    return DMB_OK;
}

extern "C" {

const char* dmb_last_error(void) { return g_err.c_str(); }
const char* dmb_version(void) { return "dmsim-b200 0.2 (sm_100a)"; }

int dmb_set_option(const char* name, int64_t value)
{
    init_options();
    if (!name) return fail(DMB_EINVAL, "null option name");
    g_option_epoch++;
    if (!strcmp(name, "tile_bits")) g_opt.tile_bits = (int)value;
    else if (!strcmp(name, "low_bits")) g_opt.low_bits = (int)value;
    else if (!strcmp(name, "min_tiles_log2")) g_opt.min_tiles_log2 = (int)value;
    else if (!strcmp(name, "graph")) g_use_graph = (int)value;
    else if (!strcmp(name, "cphase")) g_opt.cphase = value != 0;
    else if (!strcmp(name, "sparse")) g_sparse_start = (int)value;
    else if (!strcmp(name, "move_h")) g_opt.move_h = value != 0;
    else if (!strcmp(name, "hot_low")) g_opt.hot_low = value != 0;
    else if (!strcmp(name, "tma")) g_opt.tma = value != 0;
    else if (!strcmp(name, "tma_box_bits")) set_sweep_tma_box_bits((int)value);
    else if (!strcmp(name, "dense2_lu")) set_sweep_dense2_lu(value != 0);
    else if (!strcmp(name, "tma_prefetch")) g_tma_prefetch = (int)value;
    else if (!strcmp(name, "tma_l2_promotion")) g_tma_l2_promotion = (int)value;
    else if (!strcmp(name, "direct_store")) set_sweep_direct_store(value != 0);
    else if (!strcmp(name, "spread_peers")) g_opt.spread_peers = value != 0;
    else if (!strcmp(name, "heavy_last")) set_sweep_heavy_last(value != 0);
    else if (!strcmp(name, "light_first")) set_sweep_light_first(value != 0);
    else if (!strcmp(name, "persistent"))
    {
        g_persistent = (int)value;
        g_opt.small_state_bits = value ? 20 : 0; // (the cooperative executor runs sweeps with the plain tile I/O only)
    }
    else if (!strcmp(name, "jit")) g_jit = (int)value;
    else if (!strcmp(name, "jit_min_bits")) g_jit_min_bits = (int)value;
    else if (!strcmp(name, "jit_hot")) g_jit_hot = (int)value;
    else if (!strcmp(name, "jit_hot_small")) g_jit_hot_small = (int)value;
    else if (!strcmp(name, "plan_cache")) g_plan_cache = (int)value;
    else return fail(DMB_EINVAL, std::string("unknown option ") + name);
    return DMB_OK;
}

int dmb_create(int n_qubits, int world_size, int rank, int device, dmb_handle* out)
{
    init_options();
    if (!out) return fail(DMB_EINVAL, "null out handle");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 20) return fail(DMB_EINVAL, "n_qubits must be in [1, 20]");
    int g = 0;
    while ((1 << g) < world_size) g++;
    // reference ctor (:218-229): n_gpus must be a power of two and divide 2^n
    if (world_size < 1 || (1 << g) != world_size) return fail(DMB_EINVAL, "world_size must be a power of two");
    if (g > n_qubits) return fail(DMB_EINVAL, "world_size must divide 2^n_qubits");
    const bool group = rank == DMB_ALL_RANKS && world_size > 1;
    if (rank == DMB_ALL_RANKS && world_size == 1) rank = 0;
    if (!group && (rank < 0 || rank >= world_size)) return fail(DMB_EINVAL, "rank out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(DMB_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                                   " (this engine has no CPU fallback)");
    if (!group)
    {
        if (device < 0) CU(cudaGetDevice(&device));
        if (device >= ndev) return fail(DMB_EINVAL, "device index out of range");
        dmb_sim* s = nullptr;
        int rc = create_shard(n_qubits, world_size, g, rank, device, &s);
        if (rc) return rc;
        *out = s;
        return dmb_reset_dm(s);
    }
    // ---- group: all world_size shards in this process, devices [first, first + world_size) ----
    if (world_size > 8) return fail(DMB_EINVAL, "a single-process group supports up to 8 GPUs (one NVSwitch node)");
    const int first = device < 0 ? 0 : device;
    if (first + world_size > ndev)
        return fail(DMB_EINVAL, "n_gpus = " + std::to_string(world_size) + " but only " + std::to_string(ndev) +
                                    " CUDA devices are visible");
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    dmb_sim* grp = new dmb_sim;
    grp->n = n_qubits; grp->g = g; grp->world = world_size; grp->rank = DMB_ALL_RANKS; grp->device = first;
    grp->N = 2 * n_qubits; grp->M = grp->N - g;
    grp->shard_elems = (size_t)1 << grp->M;
    auto bail = [&](int rc) {
        for (dmb_sim* c : grp->shards) destroy_shard(c);
        delete grp;
        cudaSetDevice(prev_dev);
        return rc;
    };
    // every pair must be able to map each other's memory (reference :266-269 requires the same full P2P mesh)
    for (int a = 0; a < world_size; a++)
    {
        if (cudaSetDevice(first + a) != cudaSuccess) return bail(fail(DMB_ECUDA, "cudaSetDevice failed"));
        for (int b = 0; b < world_size; b++)
        {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, first + a, first + b);
            if (!can)
                return bail(fail(DMB_ECOMM, "device " + std::to_string(first + a) + " cannot access device " +
                                                std::to_string(first + b) + " (peer access is required for n_gpus > 1)"));
            cudaError_t pe = cudaDeviceEnablePeerAccess(first + b, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                return bail(fail(DMB_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)));
            cudaGetLastError();
        }
    }
    for (int r = 0; r < world_size; r++)
    {
        dmb_sim* c = nullptr;
        int rc = create_shard(n_qubits, world_size, g, r, first + r, &c);
        if (rc) return bail(rc);
        c->in_group = true;
        grp->shards.push_back(c);
        rc = ensure_second_buffer(c);
        if (rc) return bail(rc);
        if (cudaEventCreateWithFlags(&c->ev_pre, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_post, cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(DMB_ECUDA, "cudaEventCreate failed"));
    }
    for (dmb_sim* c : grp->shards)
    {
        for (int r = 0; r < world_size; r++)
            for (int b = 0; b < 2; b++) c->peer[b][r] = grp->shards[r]->buf[b];
        c->p2p = true;
    }
    grp->layout.resize(grp->N);
    *out = grp;
    int rc = dmb_reset_dm(grp);
    cudaSetDevice(prev_dev);
    return rc;
}

int dmb_destroy(dmb_handle s)
{
    if (!s) return DMB_OK;
    if (is_group(s))
    {
        for (dmb_sim* c : s->shards) destroy_shard(c);
        delete s;
        return DMB_OK;
    }
    return destroy_shard(s);
}

int dmb_reset_dm(dmb_handle s)
{
    if (!s) return fail(DMB_EINVAL, "null handle");
    if (is_group(s))
    {
        for (dmb_sim* c : s->shards)
        {
            int rc = reset_shard(c);
            if (rc) return rc;
        }
        return DMB_OK;
    }
    return reset_shard(s);
}

// Load an arbitrary state.  Every shard scatters the elements it owns out of the staged host chunk.
int dmb_set_dm(dmb_handle s, const double* real, const double* imag)
{
    if (!s || !real || !imag) return fail(DMB_EINVAL, "null argument");
    std::vector<dmb_sim*> mem = is_group(s) ? s->shards : std::vector<dmb_sim*>{s};
    const unsigned long long total = 1ull << s->N;
    const unsigned long long chunk = std::min<unsigned long long>(total, 1ull << 24);
    for (dmb_sim* c : mem)
    {
        CU(cudaSetDevice(c->device));
        for (int l = 0; l < c->N; l++) c->layout[l] = l;
        c->conj_flag = false;
        c->non_hermitian = true; // arbitrary input: keep the reference's exact frame semantics from here on
        c->support = ~0ull;
        int rc = ensure_scratch(c, chunk * 2 * sizeof(double));
        if (rc) return rc;
        const LayoutArgs L = layout_args(c);
        double* d_re = c->d_scratch;
        double* d_im = c->d_scratch + chunk;
        for (unsigned long long first = 0; first < total; first += chunk)
        {
            CU(cudaMemcpyAsync(d_re, real + first, chunk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_im, imag + first, chunk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            launch_scatter_split(c->buf[c->cur], L, first, chunk, d_re, d_im, c->stream);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(c->stream));
        }
    }
    return DMB_OK;
}

// ---- circuit ----------------------------------------------------------------------------------
// host part: plan + encode into s->host_* (no CUDA calls)
static int plan_and_encode_uncached(dmb_sim* s);
static int plan_and_encode(dmb_sim* s)
{
    unsigned long long key[2];
    plan_key_of(s, key);
    s->plan_key[0] = key[0]; s->plan_key[1] = key[1];
    if (!g_plan_cache)
    {
        s->plan_cache.clear();
        s->plan_key[0] = s->plan_key[1] = 0; // (never equal to device_key: the tables are uploaded every time)
        return plan_and_encode_uncached(s);
    }
    for (size_t i = 0; i < s->plan_cache.size(); i++)
    {
        dmb_sim::CachedPlan& c = s->plan_cache[i];
        if (c.key[0] != key[0] || c.key[1] != key[1]) continue;
        s->plan = c.plan; s->plan_layout = c.plan_layout; s->plan_conj = c.plan_conj; s->plan_nonherm = c.plan_nonherm;
        s->op_offset = c.op_offset; s->star_offset = c.star_offset; s->round_offset = c.round_offset; s->group_offset = c.group_offset;
        s->n_dev_stars = c.n_dev_stars; s->n_dev_ops = c.n_dev_ops; s->n_dev_rounds = c.n_dev_rounds; s->n_dev_groups = c.n_dev_groups;
        s->op_masks = c.op_masks; s->directs = c.directs; s->fp64_per_lane = c.fp64_per_lane;
        s->host_ops = c.host_ops; s->host_stars = c.host_stars; s->host_rounds = c.host_rounds; s->host_groups = c.host_groups;
        if (i) std::rotate(s->plan_cache.begin(), s->plan_cache.begin() + (long)i, s->plan_cache.begin() + (long)i + 1);
        return DMB_OK;
    }
    int rc = plan_and_encode_uncached(s);
    if (rc) return rc;
    dmb_sim::CachedPlan c;
    c.key[0] = key[0]; c.key[1] = key[1];
    c.plan = s->plan; c.plan_layout = s->plan_layout; c.plan_conj = s->plan_conj; c.plan_nonherm = s->plan_nonherm;
    c.op_offset = s->op_offset; c.star_offset = s->star_offset; c.round_offset = s->round_offset; c.group_offset = s->group_offset;
    c.n_dev_stars = s->n_dev_stars; c.n_dev_ops = s->n_dev_ops; c.n_dev_rounds = s->n_dev_rounds; c.n_dev_groups = s->n_dev_groups;
    c.op_masks = s->op_masks; c.directs = s->directs; c.fp64_per_lane = s->fp64_per_lane;
    c.host_ops = s->host_ops; c.host_stars = s->host_stars; c.host_rounds = s->host_rounds; c.host_groups = s->host_groups;
    s->plan_cache.insert(s->plan_cache.begin(), std::move(c));
    if (s->plan_cache.size() > kPlanCacheEntries) s->plan_cache.pop_back();
    return DMB_OK;
}

static int plan_and_encode_uncached(dmb_sim* s)
{
    try
    {
        s->plan = make_plan(s->n, s->world, s->gates.data(), s->gates.size(), s->mats.data(), s->mats.size() / 32,
                            s->layout, g_opt, s->conj_flag, s->non_hermitian);
    }
    catch (const std::invalid_argument& e)
    {
        return fail(DMB_EINVAL, e.what());
    }
    catch (const std::exception& e)
    {
        return fail(DMB_ESTATE, e.what());
    }
    s->plan_layout = s->layout;
    s->plan_conj = s->conj_flag;
    s->plan_nonherm = s->non_hermitian;
    // device op / group tables: one contiguous upload each (the reference does 3 CUDA calls per gate per GPU, :112-163)
    s->host_ops.clear(); s->host_stars.clear(); s->host_rounds.clear(); s->host_groups.clear();
    const size_t nsteps = s->plan.steps.size();
    s->op_offset.assign(nsteps, 0);
    s->round_offset.assign(nsteps, 0);
    s->group_offset.assign(nsteps, 0);
    s->n_dev_ops.assign(nsteps, 0);
    s->n_dev_rounds.assign(nsteps, 0);
    s->n_dev_groups.assign(nsteps, 0);
    s->op_masks.assign(nsteps, 0u);
    s->directs.assign(nsteps, DevDirect{});
    s->star_offset.assign(nsteps, 0);
    s->n_dev_stars.assign(nsteps, 0);
    EncodedSweep enc;
    s->fp64_per_lane = 0;
    for (size_t i = 0; i < nsteps; i++)
    {
        s->op_offset[i] = s->host_ops.size();
        s->round_offset[i] = s->host_rounds.size();
        s->group_offset[i] = s->host_groups.size();
        s->star_offset[i] = s->host_stars.size();
        if (s->plan.steps[i].kind != 0) continue;
        try
        {
            encode_sweep(s->plan.steps[i].sweep, enc);
        }
        catch (const std::exception& e)
        {
            return fail(DMB_ESTATE, e.what());
        }
        s->fp64_per_lane += enc.fp64_per_lane;
        s->n_dev_ops[i] = (int)enc.stream.size(); // bytes
        s->n_dev_stars[i] = (int)enc.stars.size();
        s->op_masks[i] = enc.op_mask;
        s->directs[i] = enc.direct;
        s->host_stars.insert(s->host_stars.end(), enc.stars.begin(), enc.stars.end());
        s->n_dev_rounds[i] = (int)enc.rounds.size();
        s->n_dev_groups[i] = (int)enc.groups.size();
        s->host_ops.insert(s->host_ops.end(), enc.stream.begin(), enc.stream.end());
        s->host_rounds.insert(s->host_rounds.end(), enc.rounds.begin(), enc.rounds.end());
        s->host_groups.insert(s->host_groups.end(), enc.groups.begin(), enc.groups.end());
    }
    return DMB_OK;
}

// shard `s` takes over the plan (and host tables) of `src` (another shard of the same group, same layout)
static void adopt_plan(dmb_sim* s, const dmb_sim* src)
{
    s->plan = src->plan;
    s->plan_layout = src->plan_layout; s->plan_conj = src->plan_conj; s->plan_nonherm = src->plan_nonherm;
    s->op_offset = src->op_offset; s->round_offset = src->round_offset; s->group_offset = src->group_offset;
    s->star_offset = src->star_offset; s->n_dev_ops = src->n_dev_ops; s->n_dev_rounds = src->n_dev_rounds;
    s->n_dev_groups = src->n_dev_groups; s->n_dev_stars = src->n_dev_stars; s->op_masks = src->op_masks;
    s->directs = src->directs;
    s->host_ops = src->host_ops; s->host_stars = src->host_stars; s->host_rounds = src->host_rounds;
    s->host_groups = src->host_groups;
    s->fp64_per_lane = src->fp64_per_lane;
    s->plan_key[0] = src->plan_key[0]; s->plan_key[1] = src->plan_key[1];
}

// device part: the single H2D of the step (stream-ordered; the copies are flushed before returning because the host
// vectors may be rebuilt by the next dmb_set_circuit)
static int upload_tables(dmb_sim* s)
{
    if (s->device_key[0] == s->plan_key[0] && s->device_key[1] == s->plan_key[1] && (s->plan_key[0] || s->plan_key[1]))
    {
        s->h2d_bytes = 0; // the device still holds this plan's tables (and the graph / parameter list captured from them)
        return DMB_OK;
    }
    drop_graph(s);
    s->jit_memo.clear();
    s->plan_runs = 0;
    s->device_key[0] = s->device_key[1] = 0;
    CU(cudaSetDevice(s->device));
    int rc;
    if ((rc = grow_device(s->d_ops, s->d_ops_cap, s->host_ops.size()))) return rc;
    if ((rc = grow_device(s->d_stars, s->d_stars_cap, s->host_stars.size()))) return rc;
    if ((rc = grow_device(s->d_rounds, s->d_rounds_cap, s->host_rounds.size()))) return rc;
    if ((rc = grow_device(s->d_groups, s->d_groups_cap, s->host_groups.size()))) return rc;
    s->h2d_bytes = s->host_ops.size() + s->host_stars.size() * sizeof(DevStar) + s->host_rounds.size() * sizeof(DevRound) +
                   s->host_groups.size() * sizeof(DevGroup);
    if (!s->host_ops.empty())
    {
        CU(cudaMemcpyAsync(s->d_ops, s->host_ops.data(), s->host_ops.size(), cudaMemcpyHostToDevice, s->stream));
        if (!s->host_stars.empty())
            CU(cudaMemcpyAsync(s->d_stars, s->host_stars.data(), s->host_stars.size() * sizeof(DevStar), cudaMemcpyHostToDevice,
                               s->stream));
        CU(cudaMemcpyAsync(s->d_rounds, s->host_rounds.data(), s->host_rounds.size() * sizeof(DevRound), cudaMemcpyHostToDevice,
                           s->stream));
        CU(cudaMemcpyAsync(s->d_groups, s->host_groups.data(), s->host_groups.size() * sizeof(DevGroup), cudaMemcpyHostToDevice,
                           s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    s->device_key[0] = s->plan_key[0]; s->device_key[1] = s->plan_key[1];
    return DMB_OK;
}

static int build_plan(dmb_sim* s)
{
    if (is_group(s))
    {
        dmb_sim* lead = s->shards[0];
        int rc = plan_and_encode(lead); // ONE plan: it depends on (world, layout), not on the rank
        if (rc) return rc;
        for (dmb_sim* c : s->shards)
        {
            if (c != lead) adopt_plan(c, lead);
            if ((rc = upload_tables(c))) return rc;
        }
        return DMB_OK;
    }
    int rc = plan_and_encode(s);
    return rc ? rc : upload_tables(s);
}

int dmb_set_circuit(dmb_handle s, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats)
{
    if (!s) return fail(DMB_EINVAL, "null handle");
    if (n_gates && !gates) return fail(DMB_EINVAL, "null gate list");
    dmb_sim* owner = is_group(s) ? s->shards[0] : s;
    owner->gates.assign(gates, gates + n_gates);
    owner->mats.assign(mats ? mats : nullptr, mats ? mats + 32 * n_mats : nullptr);
    s->have_circuit = false;
    int rc = build_plan(s);
    if (rc) return rc;
    s->have_circuit = true;
    return DMB_OK;
}

int dmb_clear_circuit(dmb_handle s)
{
    if (!s) return fail(DMB_EINVAL, "null handle");
    dmb_sim* owner = is_group(s) ? s->shards[0] : s;
    owner->gates.clear();
    owner->mats.clear();
    s->have_circuit = false;
    for (dmb_sim* c : s->shards) drop_graph(c);
    drop_graph(s);
    return DMB_OK;
}

// ---- execution --------------------------------------------------------------------------------
static int fill_sweep_args(const dmb_sim* s, size_t step, const double2* in, double2* out, SweepArgs& a,
                           unsigned long long support = ~0ull)
{
    memset(&a, 0, sizeof(a));
    const Sweep& sw = s->plan.steps[step].sweep;
    fill_sweep_tables(sw, s->M, a);
    if (~support & ((s->M >= 64 ? 0ull : (1ull << s->M)) - 1ull))
    {
        // sparse start: enumerate only the tiles whose bits outside the tile lie inside the support (in-place sweeps:
        // cin == cout); the others hold zeros before and after
        int nc = 0;
        for (int i = 0; i < a.n_comp; i++)
            if ((support >> a.cin[i]) & 1ull)
            {
                a.cin[nc] = a.cin[i];
                a.cout[nc] = a.cout[i];
                nc++;
            }
        a.n_comp = nc;
        a.n_tiles = 1ull << nc;
    }
    a.in = in;
    a.out = out;
    a.ops = s->d_ops + s->op_offset[step];
    a.rounds = s->d_rounds + s->round_offset[step];
    a.groups = s->d_groups + s->group_offset[step];
    a.ops_bytes = s->n_dev_ops[step];
    a.stars = s->d_stars + s->star_offset[step];
    a.n_stars = s->n_dev_stars[step];
    a.rank_bits = (unsigned long long)s->rank << s->M;
    a.peer_shift = -1;
    a.n_rounds = s->n_dev_rounds[step];
    a.n_groups = s->n_dev_groups[step];
    a.op_mask = s->op_masks[step];
    a.tma_prefetch = g_tma_prefetch;
    if (a.n_comp > 21) return fail(DMB_EINVAL, "shard too large for the tile-base tables");
    fill_base_tables(a);
    // direct store of the last round: in-place TMA tiles over the whole shard (not the reduced enumeration of a sparse start)
    a.direct = s->directs[step];
    if (!a.tma_load || !a.tma_store || in != out || a.n_comp != s->M - a.k) a.direct.enabled = 0;
    if (a.direct.enabled) a.tma_store = 0;
    if (a.tma_load)
    {
        std::string why;
        if (!make_tile_map(a.tma, s->M, in, a.tmap_in, why)) return fail(DMB_ECUDA, why);
        if (a.tma_store && !make_tile_map(a.tma, s->M, out, a.tmap_out, why)) return fail(DMB_ECUDA, why);
    }
    return DMB_OK;
}

static int jit_hot_of(const dmb_sim* s) { return s->M >= 24 ? g_jit_hot : g_jit_hot_small; }

// the run-time specialised kernel of sweep `step` with the I/O variant of `a`, or nullptr (interpreter kernel)
static const void* jit_resolve(dmb_sim* s, size_t step, const SweepArgs& a)
{
    if (!g_jit || s->M < g_jit_min_bits) return nullptr;
    if (g_jit == 1 && (int)s->plan_runs < jit_hot_of(s))
    {
        s->jit_missing++; // (cold plan: interpreted for now)
        return nullptr;
    }
    const unsigned long long key = ((unsigned long long)step << 4) | (a.tma_load ? 1u : 0u) | (a.tma_store ? 2u : 0u) | (a.direct.enabled ? 4u : 0u) |
                                   (a.peer_shift >= 0 ? 8u : 0u);
    JitKernel* k = nullptr;
    auto it = s->jit_memo.find(key);
    if (it != s->jit_memo.end()) k = it->second;
    else
    {
        const unsigned char* stream = s->host_ops.data() + s->op_offset[step];
        const DevRound* rounds = s->host_rounds.data() + s->round_offset[step];
        const DevGroup* groups = s->host_groups.data() + s->group_offset[step];
        unsigned long long sk[2];
        bool known = false;
        const bool keyed = jit_struct_key(a, stream, rounds, groups, sk);
        if (keyed) k = jit_lookup_struct(sk, &known);
        if (!known)
        {
            std::string defines, program;
            if (jit_available(nullptr) && jit_generate(a, stream, rounds, groups, defines, program)) k = jit_request(defines, program);
            if (keyed) jit_remember_struct(sk, k);
        }
        s->jit_memo[key] = k;
    }
    if (!k || s->jit_mode == 1) return nullptr;
    int st = k->state.load(std::memory_order_acquire);
    if (st == 0 && g_jit >= 2 && s->jit_mode != 3) { jit_wait(k); st = k->state.load(std::memory_order_acquire); }
    if (st == 0) { s->jit_missing++; return nullptr; }
    if (st != 1) return nullptr;
    if (s->jit_mode == 3 && !(k->kern && s->device >= 0 && s->device < 64 && ((k->attr_mask >> s->device) & 1ull)))
    {
        s->jit_missing++; // became ready after the dry pass: the graph is re-captured at the next run
        return nullptr;
    }
    const void* fn = nullptr;
    if (jit_kernel(k, s->device, sweep_smem_limit(), &fn) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return fn;
}

// one sweep on the stream: its specialised kernel when there is one, the interpreter kernel otherwise
static int launch_sweep_any(dmb_sim* s, size_t step, const SweepArgs& a, int grid_per_sm)
{
    const void* fn = jit_resolve(s, step, a);
    if (s->jit_mode == 1 || s->jit_mode == 2) return DMB_OK; // dry pass
    if (fn)
    {
        const int grid = (int)std::min<unsigned long long>(a.n_tiles, (unsigned long long)(grid_per_sm > 0 ? grid_per_sm * device_num_sms() : sweep_max_grid_fn(fn, a)));
        CU(launch_sweep_fn(fn, a, grid, s->stream));
        s->jit_used++;
    }
    else
    {
        const int grid = (int)std::min<unsigned long long>(a.n_tiles, (unsigned long long)(grid_per_sm > 0 ? grid_per_sm * device_num_sms() : sweep_max_grid(a)));
        CU(launch_sweep(a, grid, s->stream));
    }
    CU(cudaGetLastError());
    return DMB_OK;
}

static int comm_events(dmb_sim* s, size_t comm_idx)
{
    CU(cudaSetDevice(s->device)); // an event belongs to the device that is current when it is created
    while (s->ev_comm.size() < 2 * (comm_idx + 1))
    {
        cudaEvent_t a_;
        CU(cudaEventCreate(&a_));
        s->ev_comm.push_back(a_);
    }
    return DMB_OK;
}

static int nccl_barrier(dmb_sim* s)
{
    const int nr = g_nccl.AllReduce(s->d_barrier, s->d_barrier + 1, 1, kNcclDouble, 0 /* ncclSum */, s->comm, s->stream);
    if (nr != 0) return fail(DMB_ECOMM, std::string("NCCL barrier failed: ") + g_nccl.GetErrorString(nr));
    return DMB_OK;
}

// the permuting sweep of a remap whose stores go straight into the destination ranks' shards (peer memory)
static int launch_fused_remap(dmb_sim* s, size_t i, int cur)
{
    SweepArgs a;
    int rc = fill_sweep_args(s, i, s->buf[cur], s->buf[cur ^ 1], a);
    if (rc) return rc;
    a.tma_store = 0;
    a.peer_shift = s->M - s->g;
    a.peer_rank = s->rank;
    for (int r = 0; r < s->world; r++) a.peer_out[r] = (unsigned long long)s->peer[cur ^ 1][r];
    return launch_sweep_any(s, i, a, 0);
}

static bool is_fused_remap(const dmb_sim* s, size_t i)
{
    const Step& st = s->plan.steps[i];
    return st.kind == 0 && st.sweep.out_of_place && s->p2p && i + 1 < s->plan.steps.size() && s->plan.steps[i + 1].kind == 1;
}

// one local sweep of shard s (step i) on its stream
static int enqueue_sweep(dmb_sim* s, size_t i, int& cur, uint64_t& launches, unsigned long long& support)
{
    const Sweep& sw = s->plan.steps[i].sweep;
    double2* in = s->buf[cur];
    double2* out = in;
    if (sw.out_of_place)
    {
        int rc = ensure_second_buffer(s);
        if (rc) return rc;
        out = s->buf[cur ^ 1];
    }
    SweepArgs a;
    if (sw.out_of_place) support = ~0ull; // (only multi-GPU remaps permute; kept general)
    int rca = fill_sweep_args(s, i, in, out, a, support);
    if (rca) return rca;
    for (int j = 0; j < sw.k; j++) support |= 1ull << sw.in_pos[j];
    int rcl = launch_sweep_any(s, i, a, g_grid_per_sm);
    if (rcl) return rcl;
    launches++;
    if (sw.out_of_place) cur ^= 1;
    return DMB_OK;
}

// enqueue every step of the plan on s->stream (one shard per process); cur is updated as buffers flip
static int enqueue_steps(dmb_sim* s, int& cur, uint64_t& launches, bool allow_exchange, unsigned long long& support)
{
    size_t comm_idx = 0;
    if (s->world != 1) support = ~0ull;
    for (size_t i = 0; i < s->plan.steps.size(); i++)
    {
        const Step& st = s->plan.steps[i];
        // qubit remap with peer memory: the permuting sweep stores straight into the destination ranks' shards (one
        // kernel = pack + all-to-all over NVLink) between two one-element all-reduces: the first makes sure every peer
        // is done with the buffer that is about to be overwritten (its earlier sweeps, readouts or resets may still be
        // running), the second that every peer's stores have landed
        const bool dry = s->jit_mode == 1 || s->jit_mode == 2; // (resolving the specialised kernels only: nothing is enqueued)
        if (dry && (st.kind == 1 || (allow_exchange && is_fused_remap(s, i))))
        {
            if (st.kind == 0)
            {
                int rc = launch_fused_remap(s, i, cur);
                if (rc) return rc;
                i++;
            }
            cur ^= 1;
            continue;
        }
        if (allow_exchange && is_fused_remap(s, i))
        {
            int rc = comm_events(s, comm_idx);
            if (rc) return rc;
            CU(cudaEventRecord(s->ev_comm[2 * comm_idx], s->stream));
            if ((rc = nccl_barrier(s))) return rc;
            if ((rc = launch_fused_remap(s, i, cur))) return rc;
            if ((rc = nccl_barrier(s))) return rc;
            CU(cudaEventRecord(s->ev_comm[2 * comm_idx + 1], s->stream));
            comm_idx++;
            launches += 3;
            cur ^= 1;
            i++; // the exchange step is done
            continue;
        }
        if (st.kind == 0)
        {
            int rc = enqueue_sweep(s, i, cur, launches, support);
            if (rc) return rc;
        }
        else
        {
            if (!allow_exchange || !s->comm) return fail(DMB_ECOMM, "plan needs a qubit-remap exchange but no communicator is attached (dmb_comm_init)");
            int rc = ensure_second_buffer(s);
            if (rc) return rc;
            const int P = s->world;
            const size_t chunk = s->shard_elems / P; // complex elements per peer
            const double2* src = s->buf[cur];
            double2* dst = s->buf[cur ^ 1];
            if ((rc = comm_events(s, comm_idx))) return rc;
            CU(cudaEventRecord(s->ev_comm[2 * comm_idx], s->stream));
            CU(cudaMemcpyAsync(dst + (size_t)s->rank * chunk, src + (size_t)s->rank * chunk, chunk * sizeof(double2),
                               cudaMemcpyDeviceToDevice, s->stream));
            int nr = g_nccl.GroupStart();
            for (int p = 0; p < P && nr == 0; p++)
            {
                if (p == s->rank) continue;
                nr = g_nccl.Send(src + (size_t)p * chunk, chunk * 2, kNcclDouble, p, s->comm, s->stream);
                if (nr == 0) nr = g_nccl.Recv(dst + (size_t)p * chunk, chunk * 2, kNcclDouble, p, s->comm, s->stream);
            }
            const int ne = g_nccl.GroupEnd();
            if (nr == 0) nr = ne;
            if (nr != 0) return fail(DMB_ECOMM, std::string("NCCL exchange failed: ") + g_nccl.GetErrorString(nr));
            CU(cudaEventRecord(s->ev_comm[2 * comm_idx + 1], s->stream));
            comm_idx++;
            launches += 1;
            cur ^= 1;
        }
    }
    return DMB_OK;
}

// resolves the run-time specialised kernels of the plan without enqueuing anything (mode 1: queue the compilations; 2: wait /
// load as the "jit" option says)
static int jit_dry_pass(dmb_sim* s, int mode)
{
    if (!g_jit || s->M < g_jit_min_bits) return DMB_OK;
    s->jit_mode = mode;
    int c2 = s->cur;
    uint64_t l2 = 0;
    unsigned long long sup2 = s->support;
    const int rc = enqueue_steps(s, c2, l2, true, sup2);
    s->jit_mode = 0;
    return rc;
}

// every shard's stream waits until all the OTHER shards have reached the event `which` (0 = ev_pre, 1 = ev_post)
static int group_barrier(dmb_sim* grp, int which)
{
    for (dmb_sim* c : grp->shards)
    {
        CU(cudaSetDevice(c->device));
        CU(cudaEventRecord(which ? c->ev_post : c->ev_pre, c->stream));
    }
    for (dmb_sim* c : grp->shards)
    {
        CU(cudaSetDevice(c->device));
        for (dmb_sim* o : grp->shards)
            if (o != c) CU(cudaStreamWaitEvent(c->stream, which ? o->ev_post : o->ev_pre, 0));
    }
    return DMB_OK;
}

// single-process group: the host walks the plan step by step and feeds every shard's stream (the launches are
// asynchronous, so the devices run concurrently like the reference's one-OpenMP-thread-per-GPU loop, :397-451)
static int group_run(dmb_sim* grp, dmb_stats* stats)
{
    dmb_sim* lead = grp->shards[0];
    const size_t nsteps = lead->plan.steps.size();
    for (dmb_sim* c : grp->shards)
    {
        c->jit_missing = c->jit_used = 0;
        c->plan_runs++;
    }
    if (lead->jit_memo.empty())
    {
        // queue every compilation of the plan at once (the shards share the kernels); the waiting mode waits here, before
        // the timed region
        CU(cudaSetDevice(lead->device));
        sweep_setup();
        int rcj = jit_dry_pass(lead, 1);
        if (!rcj && g_jit >= 2) rcj = jit_dry_pass(lead, 2);
        if (rcj) return rcj;
    }
    for (dmb_sim* c : grp->shards)
    {
        CU(cudaSetDevice(c->device));
        sweep_setup();
        CU(cudaEventRecord(c->ev_begin, c->stream));
    }
    uint64_t launches = 0;
    size_t comm_idx = 0;
    for (size_t i = 0; i < nsteps; i++)
    {
        const Step& st = lead->plan.steps[i];
        const bool fused = is_fused_remap(lead, i);
        if (fused || st.kind == 1)
        {
            int rc = comm_events(lead, comm_idx);
            if (rc) return rc;
            CU(cudaSetDevice(lead->device));
            CU(cudaEventRecord(lead->ev_comm[2 * comm_idx], lead->stream));
            if ((rc = group_barrier(grp, 0))) return rc; // every shard is done with the buffer the peers will write
            for (dmb_sim* c : grp->shards)
            {
                CU(cudaSetDevice(c->device));
                if (fused)
                {
                    if ((rc = launch_fused_remap(c, i, c->cur))) return rc;
                    launches++;
                }
                else
                {
                    // plain exchange (no permuting sweep in front): chunk p of rank r becomes chunk r of rank p
                    const size_t chunk = c->shard_elems / c->world;
                    for (int p = 0; p < c->world; p++)
                        CU(cudaMemcpyPeerAsync(c->peer[c->cur ^ 1][p] + (size_t)c->rank * chunk, grp->shards[p]->device,
                                               c->buf[c->cur] + (size_t)p * chunk, c->device, chunk * sizeof(double2), c->stream));
                    launches += c->world;
                }
                c->cur ^= 1;
            }
            if ((rc = group_barrier(grp, 1))) return rc; // every peer's stores have landed
            CU(cudaSetDevice(lead->device));
            CU(cudaEventRecord(lead->ev_comm[2 * comm_idx + 1], lead->stream));
            comm_idx++;
            if (fused) i++;
            continue;
        }
        for (dmb_sim* c : grp->shards)
        {
            CU(cudaSetDevice(c->device));
            unsigned long long support = ~0ull;
            int rc = enqueue_sweep(c, i, c->cur, launches, support);
            if (rc) return rc;
        }
    }
    for (dmb_sim* c : grp->shards)
    {
        CU(cudaSetDevice(c->device));
        CU(cudaEventRecord(c->ev_end, c->stream));
    }
    double sim_ms = 0;
    for (dmb_sim* c : grp->shards)
    {
        CU(cudaSetDevice(c->device));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        c->layout = lead->plan.end_layout;
        c->conj_flag = lead->plan.conj_end;
        if (lead->plan.has_srn) c->non_hermitian = true;
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end));
        sim_ms = std::max(sim_ms, (double)ms);
    }
    if (stats)
    {
        memset(stats, 0, sizeof(*stats));
        stats->sim_ms = sim_ms;
        double comm = 0;
        CU(cudaSetDevice(lead->device));
        for (size_t i = 0; i < comm_idx; i++)
        {
            float c = 0;
            CU(cudaEventElapsedTime(&c, lead->ev_comm[2 * i], lead->ev_comm[2 * i + 1]));
            comm += c;
        }
        stats->comm_ms = comm;
        stats->comp_ms = sim_ms - comm;
        stats->n_gates = lead->plan.n_gates;
        stats->n_primitives = lead->plan.n_primitives;
        stats->n_blocks = lead->plan.n_blocks;
        stats->n_sweeps = lead->plan.n_sweeps;
        stats->n_exchanges = lead->plan.n_exchanges;
        stats->n_launches = launches;
        stats->sweep_bytes = 32ull * lead->shard_elems;
        stats->exchange_bytes = lead->plan.n_exchanges * (uint64_t)(lead->world - 1) * (lead->shard_elems / lead->world) * 16ull;
        stats->h2d_bytes = lead->h2d_bytes * grp->shards.size();
        stats->fp64_ops = lead->shard_elems >= (size_t)kRegElems ? lead->fp64_per_lane * (lead->shard_elems / kRegElems) : lead->fp64_per_lane;
    }
    return DMB_OK;
}

int dmb_run(dmb_handle s, dmb_stats* stats)
{
    if (!s) return fail(DMB_EINVAL, "null handle");
    if (!s->have_circuit) return fail(DMB_ESTATE, "dmb_run before dmb_set_circuit");
    if (is_group(s))
    {
        dmb_sim* lead = s->shards[0];
        if (lead->plan_layout != lead->layout || lead->plan_conj != lead->conj_flag || lead->plan_nonherm != lead->non_hermitian)
        {
            int rc = build_plan(s);
            if (rc) return rc;
        }
        int prev = 0;
        cudaGetDevice(&prev);
        int rc = group_run(s, stats);
        cudaSetDevice(prev);
        return rc;
    }
    CU(cudaSetDevice(s->device));
    if (s->plan_layout != s->layout || s->plan_conj != s->conj_flag || s->plan_nonherm != s->non_hermitian)
    {
        int rc = build_plan(s); // state layout changed since planning (reset / previous run): re-plan
        if (rc) return rc;
    }
    s->plan_runs++;
    uint64_t launches = 0;
    int cur = s->cur;
    // small states (no sweep of the plan uses TMA tile I/O: PlanOptions::small_state_bits) -> ONE cooperative launch for the whole run
    bool persistent = g_persistent && s->world == 1 && s->plan.n_exchanges == 0 && s->plan.n_sweeps > 1;
    for (const Step& st : s->plan.steps)
        if (st.kind != 0 || st.sweep.swz_mode != kSwzXor3) persistent = false;
    if (persistent)
    {
        if (!s->multi_valid || s->multi_cur != s->cur || s->multi_support != s->support)
        {
            std::vector<SweepArgs> list;
            int c2 = s->cur;
            unsigned long long sup = s->support, max_tiles = 1;
            unsigned mask = 0;
            size_t smem = 0;
            for (size_t i = 0; i < s->plan.steps.size(); i++)
            {
                const Sweep& sw = s->plan.steps[i].sweep;
                double2* in = s->buf[c2];
                double2* out = in;
                if (sw.out_of_place)
                {
                    int rc = ensure_second_buffer(s);
                    if (rc) return rc;
                    out = s->buf[c2 ^ 1];
                    sup = ~0ull;
                }
                SweepArgs a;
                int rc = fill_sweep_args(s, i, in, out, a, sup);
                if (rc) return rc;
                for (int j = 0; j < sw.k; j++) sup |= 1ull << sw.in_pos[j];
                if (sw.out_of_place) c2 ^= 1;
                mask |= a.op_mask;
                smem = std::max(smem, sweep_smem_bytes(a));
                max_tiles = std::max(max_tiles, a.n_tiles);
                list.push_back(a);
            }
            if (list.size() > s->d_multi_cap)
            {
                if (s->d_multi) cudaFree(s->d_multi);
                s->d_multi = nullptr;
                s->d_multi_cap = 0;
                CU(cudaMalloc(&s->d_multi, list.size() * sizeof(SweepArgs)));
                s->d_multi_cap = list.size();
            }
            CU(cudaMemcpyAsync(s->d_multi, list.data(), list.size() * sizeof(SweepArgs), cudaMemcpyHostToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream)); // (list is a local vector)
            s->h2d_bytes += list.size() * sizeof(SweepArgs);
            s->multi_valid = true;
            s->multi_cur = s->cur; s->multi_support = s->support; s->multi_end_cur = c2; s->multi_n = (int)list.size();
            s->multi_mask = mask; s->multi_smem = smem; s->multi_max_tiles = max_tiles;
        }
        CU(cudaEventRecord(s->ev_begin, s->stream));
        int grid = 0;
        cudaError_t le = launch_multi_sweep(s->d_multi, s->multi_n, s->multi_mask, s->multi_smem, s->multi_max_tiles, s->stream, &grid);
        if (le == cudaSuccess)
        {
            CU(cudaEventRecord(s->ev_end, s->stream));
            launches = 1;
            cur = s->multi_end_cur;
            for (const Step& st : s->plan.steps)
                for (int j = 0; j < st.sweep.k; j++) s->support |= 1ull << st.sweep.in_pos[j];
        }
        else
        {
            cudaGetLastError(); // (no cooperative launch here: the per-sweep executors below take over)
            persistent = false;
        }
    }
    const bool graphable = !persistent && g_use_graph && s->plan.n_exchanges == 0 && s->plan.n_sweeps > 1;
    if (persistent) {}
    else if (graphable)
    {
        // (tiered execution: sweeps captured on the interpreter kernel move to their specialised kernel once it is compiled)
        if (s->graph_exec && g_jit == 1 && (int)s->plan_runs == jit_hot_of(s) && s->jit_memo.empty())
        {
            // the plan has turned hot: queue its compilations and capture again -- kernels of a known structure are ready at once,
            // the others move in when they are done (the `stale` test below)
            int rcj = jit_dry_pass(s, 1);
            if (rcj) return rcj;
            drop_graph(s);
        }
        const bool stale = s->graph_exec && s->graph_jit_missing > 0 && jit_ready_count() != s->graph_jit_epoch;
        if (!s->graph_exec || s->graph_cur != s->cur || s->graph_support != s->support || stale)
        {
            drop_graph(s);
            for (const Step& st : s->plan.steps)
                if (st.kind == 0 && st.sweep.out_of_place)
                {
                    int rc = ensure_second_buffer(s);
                    if (rc) return rc;
                }
            sweep_setup(); // attribute setup must not happen inside capture
            // ... nor the loading of run-time specialised kernels: resolve them first (all compilations queued, then waited for / loaded)
            int rcj = jit_dry_pass(s, 1);
            if (!rcj) rcj = jit_dry_pass(s, 2);
            if (rcj) return rcj;
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            int c2 = s->cur;
            uint64_t l2 = 0;
            unsigned long long sup2 = s->support;
            s->jit_mode = 3;
            s->jit_missing = s->jit_used = 0;
            int rc = enqueue_steps(s, c2, l2, false, sup2);
            s->jit_mode = 0;
            s->graph_jit_missing = s->jit_missing;
            s->graph_jit_epoch = jit_ready_count();
            cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            CU(ce);
            CU(cudaGraphInstantiate(&s->graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            s->graph_cur = s->cur;
            s->graph_support = s->support;
        }
        CU(cudaEventRecord(s->ev_begin, s->stream));
        CU(cudaGraphLaunch(s->graph_exec, s->stream));
        CU(cudaEventRecord(s->ev_end, s->stream));
        launches = s->plan.n_sweeps;
        for (const Step& st : s->plan.steps)
        {
            if (st.kind != 0) continue;
            if (st.sweep.out_of_place) cur ^= 1;
            for (int j = 0; j < st.sweep.k; j++) s->support |= 1ull << st.sweep.in_pos[j];
        }
    }
    else
    {
        sweep_setup();
        if (s->jit_memo.empty())
        {
            // queue every compilation of the plan at once; the waiting mode waits HERE, before the timed region
            int rcj = jit_dry_pass(s, 1);
            if (!rcj && g_jit >= 2) rcj = jit_dry_pass(s, 2);
            if (rcj) return rcj;
        }
        s->jit_missing = s->jit_used = 0;
        CU(cudaEventRecord(s->ev_begin, s->stream));
        int rc = enqueue_steps(s, cur, launches, true, s->support);
        if (rc) return rc;
        CU(cudaEventRecord(s->ev_end, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    s->cur = cur;
    s->layout = s->plan.end_layout;
    s->conj_flag = s->plan.conj_end;
    if (s->plan.has_srn) s->non_hermitian = true;
    if (stats)
    {
        memset(stats, 0, sizeof(*stats));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev_begin, s->ev_end));
        stats->sim_ms = ms;
        double comm = 0;
        for (size_t i = 0; i < s->plan.n_exchanges && !graphable; i++)
        {
            float c = 0;
            CU(cudaEventElapsedTime(&c, s->ev_comm[2 * i], s->ev_comm[2 * i + 1]));
            comm += c;
        }
        stats->comm_ms = comm;
        stats->comp_ms = stats->sim_ms - comm;
        stats->n_gates = s->plan.n_gates;
        stats->n_primitives = s->plan.n_primitives;
        stats->n_blocks = s->plan.n_blocks;
        stats->n_sweeps = s->plan.n_sweeps;
        stats->n_exchanges = s->plan.n_exchanges;
        stats->n_launches = launches;
        stats->sweep_bytes = 32ull * s->shard_elems;
        stats->exchange_bytes = s->plan.n_exchanges * (uint64_t)(s->world - 1) * (s->shard_elems / s->world) * 16ull;
        stats->h2d_bytes = s->h2d_bytes;
        stats->fp64_ops = s->shard_elems >= (size_t)kRegElems ? s->fp64_per_lane * (s->shard_elems / kRegElems) : s->fp64_per_lane;
    }
    return DMB_OK;
}

// ---- results ----------------------------------------------------------------------------------
// A sharded state answers GLOBALLY: a group handle gathers over its shards; a rank of a one-process-per-GPU job with
// a communicator attached all-reduces (every rank must make the call, every rank gets the full answer -- the
// reference's distributed measure() does the same with MPI_Gather / MPI_Bcast, src/dmsim_nvgpu_mpi.cuh:472-520).
int dmb_get_dm(dmb_handle s, double* real, double* imag)
{
    if (!s || !real || !imag) return fail(DMB_EINVAL, "null argument");
    if (s->world != 1 && !is_group(s) && !collective(s))
        return fail(DMB_ESTATE, "dmb_get_dm on one rank of a sharded state needs a communicator (dmb_comm_init)");
    const unsigned long long total = 1ull << s->N;
    const unsigned long long chunk = std::min<unsigned long long>(total, 1ull << 24);
    dmb_sim* lead = is_group(s) ? s->shards[0] : s;
    CU(cudaSetDevice(lead->device));
    int rc = ensure_scratch(lead, chunk * 2 * sizeof(double));
    if (rc) return rc;
    double* d_re = lead->d_scratch;
    double* d_im = lead->d_scratch + chunk;
    for (unsigned long long first = 0; first < total; first += chunk)
    {
        if (is_group(s))
        {
            // every shard writes the elements it owns straight into the lead device's staging buffers (peer memory)
            for (dmb_sim* c : s->shards)
            {
                CU(cudaSetDevice(c->device));
                launch_gather_split(c->buf[c->cur], layout_args(c), first, chunk, d_re, d_im, 1 /* owned only */, c->stream);
                CU(cudaGetLastError());
            }
            for (dmb_sim* c : s->shards)
            {
                CU(cudaSetDevice(c->device));
                CU(cudaStreamSynchronize(c->stream));
            }
            CU(cudaSetDevice(lead->device));
        }
        else
        {
            launch_gather_split(s->buf[s->cur], layout_args(s), first, chunk, d_re, d_im, s->world > 1 ? 2 /* zero the rest */ : 0,
                                s->stream);
            CU(cudaGetLastError());
            if (collective(s) && (rc = all_reduce_sum(s, d_re, 2 * chunk))) return rc;
        }
        CU(cudaMemcpyAsync(real + first, d_re, chunk * sizeof(double), cudaMemcpyDeviceToHost, lead->stream));
        CU(cudaMemcpyAsync(imag + first, d_im, chunk * sizeof(double), cudaMemcpyDeviceToHost, lead->stream));
        CU(cudaStreamSynchronize(lead->stream));
    }
    return DMB_OK;
}

// one shard's part of n arbitrary elements (zero where another rank owns the element) into host arrays
static int shard_elements(dmb_sim* s, const uint64_t* flat_index, size_t n, double* real, double* imag, bool reduce)
{
    CU(cudaSetDevice(s->device));
    int rc = ensure_scratch(s, n * (sizeof(unsigned long long) + 2 * sizeof(double)));
    if (rc) return rc;
    double* d_re = s->d_scratch;
    double* d_im = d_re + n;
    unsigned long long* d_idx = reinterpret_cast<unsigned long long*>(d_im + n);
    CU(cudaMemcpyAsync(d_idx, flat_index, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, s->stream));
    launch_gather_elements(s->buf[s->cur], layout_args(s), d_idx, n, d_re, d_im, s->stream);
    CU(cudaGetLastError());
    if (reduce && (rc = all_reduce_sum(s, d_re, 2 * n))) return rc;
    CU(cudaMemcpyAsync(real, d_re, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(imag, d_im, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DMB_OK;
}

int dmb_get_elements(dmb_handle s, const uint64_t* flat_index, size_t n, double* real, double* imag)
{
    if (!s || (n && (!flat_index || !real || !imag))) return fail(DMB_EINVAL, "null argument");
    if (!n) return DMB_OK;
    const uint64_t total = 1ull << s->N;
    for (size_t i = 0; i < n; i++)
        if (flat_index[i] >= total) return fail(DMB_EINVAL, "element index out of range");
    if (!is_group(s)) return shard_elements(s, flat_index, n, real, imag, collective(s));
    std::vector<double> re(n), im(n);
    std::fill(real, real + n, 0.0);
    std::fill(imag, imag + n, 0.0);
    for (dmb_sim* c : s->shards)
    {
        int rc = shard_elements(c, flat_index, n, re.data(), im.data(), false);
        if (rc) return rc;
        for (size_t i = 0; i < n; i++) { real[i] += re[i]; imag[i] += im[i]; }
    }
    return DMB_OK;
}

// diagonal of one shard (zero where another rank owns the entry): real parts or their absolute values, left on the
// device in s->d_scratch[0 .. dim) (all-reduced when `reduce`)
static int shard_diag_device(dmb_sim* s, bool abs_values, bool reduce, size_t extra_bytes = 0)
{
    CU(cudaSetDevice(s->device));
    const size_t dim = (size_t)1 << s->n;
    int rc = ensure_scratch(s, dim * sizeof(double) + extra_bytes);
    if (rc) return rc;
    launch_diag(s->buf[s->cur], layout_args(s), abs_values ? nullptr : s->d_scratch, abs_values ? s->d_scratch : nullptr, s->stream);
    CU(cudaGetLastError());
    if (reduce && (rc = all_reduce_sum(s, s->d_scratch, dim))) return rc;
    return DMB_OK;
}

// the full diagonal (or |diagonal|) of a group on the host
static int group_diag_host(dmb_sim* grp, bool abs_values, std::vector<double>& acc)
{
    const size_t dim = (size_t)1 << grp->n;
    acc.assign(dim, 0.0);
    std::vector<double> part(dim);
    for (dmb_sim* c : grp->shards)
    {
        int rc = shard_diag_device(c, abs_values, false);
        if (rc) return rc;
        CU(cudaMemcpyAsync(part.data(), c->d_scratch, dim * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < dim; i++) acc[i] += part[i];
    }
    return DMB_OK;
}

int dmb_get_diag(dmb_handle s, double* diag)
{
    if (!s || !diag) return fail(DMB_EINVAL, "null argument");
    const size_t dim = (size_t)1 << s->n;
    if (is_group(s))
    {
        std::vector<double> acc;
        int rc = group_diag_host(s, false, acc);
        if (rc) return rc;
        memcpy(diag, acc.data(), dim * sizeof(double));
        return DMB_OK;
    }
    int rc = shard_diag_device(s, false, collective(s));
    if (rc) return rc;
    CU(cudaMemcpyAsync(diag, s->d_scratch, dim * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DMB_OK;
}

static int reduce_scalar(dmb_sim* s, bool purity, double* out)
{
    if (!s || !out) return fail(DMB_EINVAL, "null argument");
    if (is_group(s))
    {
        double acc = 0.0;
        for (dmb_sim* c : s->shards)
        {
            double v = 0.0;
            int rc = reduce_scalar(c, purity, &v);
            if (rc) return rc;
            acc += v;
        }
        *out = acc;
        return DMB_OK;
    }
    CU(cudaSetDevice(s->device));
    int rc = ensure_scratch(s, sizeof(double));
    if (rc) return rc;
    CU(cudaMemsetAsync(s->d_scratch, 0, sizeof(double), s->stream));
    if (purity) launch_purity(s->buf[s->cur], s->shard_elems, s->d_scratch, s->stream);
    else launch_trace(s->buf[s->cur], layout_args(s), s->d_scratch, s->stream);
    CU(cudaGetLastError());
    if (collective(s) && (rc = all_reduce_sum(s, s->d_scratch, 1))) return rc;
    CU(cudaMemcpyAsync(out, s->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return DMB_OK;
}
int dmb_trace(dmb_handle s, double* trace) { return reduce_scalar(s, false, trace); }
int dmb_purity(dmb_handle s, double* purity) { return reduce_scalar(s, true, purity); }

int dmb_sample(dmb_handle s, const double* r, size_t n, uint64_t* out, double* total)
{
    if (!s || (n && (!r || !out))) return fail(DMB_EINVAL, "null argument");
    if (s->world != 1 && !is_group(s) && !collective(s))
        return fail(DMB_ESTATE, "dmb_sample on one rank of a sharded state needs a communicator (dmb_comm_init)");
    const size_t dim = (size_t)1 << s->n;
    // scratch of the sampling device: p[dim] | scan[dim+1] | r[n] | out[n]
    const size_t bytes = (2 * dim + 1 + n) * sizeof(double) + n * sizeof(unsigned long long);
    dmb_sim* lead = is_group(s) ? s->shards[0] : s;
    int rc;
    if (is_group(s))
    {
        std::vector<double> p;
        if ((rc = group_diag_host(s, true, p))) return rc; // |Re rho_ii| of every shard, summed (one owner per entry)
        CU(cudaSetDevice(lead->device));
        if ((rc = ensure_scratch(lead, bytes))) return rc;
        CU(cudaMemcpyAsync(lead->d_scratch, p.data(), dim * sizeof(double), cudaMemcpyHostToDevice, lead->stream));
        CU(cudaStreamSynchronize(lead->stream)); // p is a local vector
    }
    else if ((rc = shard_diag_device(s, true, collective(s), bytes - dim * sizeof(double))))
        return rc;
    double* d_p = lead->d_scratch;
    double* d_scan = d_p + dim;
    double* d_r = d_scan + dim + 1;
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(d_r + n);
    launch_scan(d_p, d_scan, dim, lead->stream);
    if (n)
    {
        CU(cudaMemcpyAsync(d_r, r, n * sizeof(double), cudaMemcpyHostToDevice, lead->stream));
        launch_sample(d_scan, dim, d_r, n, d_out, lead->stream);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, d_out, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, lead->stream));
    }
    if (total) CU(cudaMemcpyAsync(total, d_scan + dim, sizeof(double), cudaMemcpyDeviceToHost, lead->stream));
    CU(cudaStreamSynchronize(lead->stream));
    CU(cudaGetLastError());
    return DMB_OK;
}

int dmb_measure(dmb_handle s, unsigned seed, size_t repetition, uint64_t* out, double* total)
{
    // reference measure() :534-539: srand(RAND_SEED); r = rand()/RAND_MAX per shot
    std::vector<double> r(repetition);
    srand(seed);
    for (size_t i = 0; i < repetition; i++) r[i] = (double)rand() / (double)RAND_MAX;
    return dmb_sample(s, r.data(), repetition, out, total);
}

int dmb_get_shard(dmb_handle s, double* interleaved, int32_t* phys_of_logical)
{
    if (!s) return fail(DMB_EINVAL, "null handle");
    if (is_group(s))
    {
        // group: the shards back to back in rank order (2 * 4^n doubles), i.e. the whole state in PHYSICAL order
        for (dmb_sim* c : s->shards)
        {
            int rc = dmb_get_shard(c, interleaved ? interleaved + 2 * c->shard_elems * (size_t)c->rank : nullptr, phys_of_logical);
            if (rc) return rc;
        }
        return DMB_OK;
    }
    CU(cudaSetDevice(s->device));
    if (interleaved)
    {
        CU(cudaMemcpyAsync(interleaved, s->buf[s->cur], s->shard_elems * sizeof(double2), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    if (phys_of_logical)
        for (int l = 0; l < s->N; l++) phys_of_logical[l] = s->layout[l];
    return DMB_OK;
}

// ---- multi-GPU --------------------------------------------------------------------------------
int dmb_comm_unique_id(uint8_t id[128])
{
    std::string why;
    if (!load_nccl(why)) return fail(DMB_ECOMM, why);
    ncclUniqueId u;
    int rc = g_nccl.GetUniqueId(&u);
    if (rc) return fail(DMB_ECOMM, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    memcpy(id, u.internal, 128);
    return DMB_OK;
}

int dmb_comm_init(dmb_handle s, const uint8_t id[128])
{
    if (!s || !id) return fail(DMB_EINVAL, "null argument");
    std::string why;
    if (!load_nccl(why)) return fail(DMB_ECOMM, why);
    CU(cudaSetDevice(s->device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    int rc = g_nccl.CommInitRank(&s->comm, s->world, u, s->rank);
    if (rc) return fail(DMB_ECOMM, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
    return DMB_OK;
}

int dmb_comm_export(dmb_handle s, uint8_t out[128])
{
    if (!s || !out) return fail(DMB_EINVAL, "null argument");
    CU(cudaSetDevice(s->device));
    int rc = ensure_second_buffer(s);
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int b = 0; b < 2; b++)
    {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, s->buf[b]));
        memcpy(out + 64 * b, &h, 64);
    }
    return DMB_OK;
}

int dmb_comm_import(dmb_handle s, const uint8_t* all_handles)
{
    if (!s || !all_handles) return fail(DMB_EINVAL, "null argument");
    if (s->world > 8) return fail(DMB_EINVAL, "peer-memory exchange supports up to 8 ranks (one NVSwitch node)");
    if (!s->comm) return fail(DMB_ESTATE, "dmb_comm_import needs dmb_comm_init first (the barrier runs on the communicator)");
    CU(cudaSetDevice(s->device));
    for (int r = 0; r < s->world; r++)
        for (int b = 0; b < 2; b++)
        {
            if (r == s->rank) { s->peer[b][r] = s->buf[b]; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + 128 * (size_t)r + 64 * b, 64);
            void* ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
            {
                cudaGetLastError();
                return fail(DMB_ECOMM, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
            }
            s->peer[b][r] = reinterpret_cast<double2*>(ptr);
        }
    if (!s->d_barrier)
    {
        CU(cudaMalloc(&s->d_barrier, 2 * sizeof(double)));
        CU(cudaMemset(s->d_barrier, 0, 2 * sizeof(double)));
    }
    s->p2p = true;
    return DMB_OK;
}

int dmb_comm_p2p(dmb_handle s, int enable)
{
    if (!s) return fail(DMB_EINVAL, "null handle");
    if (enable && !s->d_barrier) return fail(DMB_ESTATE, "dmb_comm_p2p(1) needs a successful dmb_comm_import first");
    s->p2p = enable != 0;
    return DMB_OK;
}

// ---- planner introspection (host only) ----------------------------------------------------------
int64_t dmb_plan_json(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats,
                      size_t n_mats, const int32_t* start_layout, int conj_state, char* out, size_t cap)
{
    init_options();
    try
    {
        std::vector<int> start;
        if (start_layout) start.assign(start_layout, start_layout + 2 * n_qubits);
        Plan p = make_plan(n_qubits, world_size, gates, n_gates, mats, n_mats, start, g_opt, (conj_state & 1) != 0,
                           (conj_state & 2) != 0);
        std::vector<std::string> extra(p.steps.size());
        for (size_t i = 0; i < p.steps.size(); i++)
        {
            if (p.steps[i].kind != 0) continue;
            EncodedSweep enc;
            encode_sweep(p.steps[i].sweep, enc);
            SweepArgs a;
            memset(&a, 0, sizeof(a));
            fill_sweep_tables(p.steps[i].sweep, 2 * n_qubits - p.g, a);
            extra[i] = encoded_to_json(enc, a);
        }
        std::string js = plan_to_json(p, &extra);
        if (out && cap > 0)
        {
            const size_t ncopy = std::min(cap - 1, js.size());
            memcpy(out, js.data(), ncopy);
            out[ncopy] = 0;
        }
        return (int64_t)js.size() + 1;
    }
    catch (const std::invalid_argument& e)
    {
        return fail(DMB_EINVAL, e.what());
    }
    catch (const std::exception& e)
    {
        return fail(DMB_ESTATE, e.what());
    }
}

// host-only: plan + encode the circuit and generate the program text of sweep `sweep_index` (0: no such sweep / not covered)
static int jit_text_of(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
                       int sweep_index, int peer, std::string& defines, std::string& program)
{
    Plan p = make_plan(n_qubits, world_size, gates, n_gates, mats, n_mats, std::vector<int>(), g_opt, false, false);
    int seen = -1;
    for (size_t i = 0; i < p.steps.size(); i++)
    {
        if (p.steps[i].kind != 0 || ++seen != sweep_index) continue;
        const Sweep& sw = p.steps[i].sweep;
        EncodedSweep enc;
        encode_sweep(sw, enc);
        SweepArgs a;
        memset(&a, 0, sizeof(a));
        fill_sweep_tables(sw, 2 * n_qubits - p.g, a);
        a.ops_bytes = (int)enc.stream.size();
        a.n_rounds = (int)enc.rounds.size();
        a.n_groups = (int)enc.groups.size();
        a.n_stars = (int)enc.stars.size();
        a.op_mask = enc.op_mask;
        a.peer_shift = peer ? 0 : -1;
        a.direct = enc.direct;
        if (!a.tma_load || !a.tma_store || sw.out_of_place) a.direct.enabled = 0;
        if (a.direct.enabled || peer) a.tma_store = 0;
        return jit_generate(a, enc.stream.data(), enc.rounds.data(), enc.groups.data(), defines, program) ? 1 : 0;
    }
    return 0;
}

// The CUDA text the run-time compiler would be given for sweep `sweep_index` of the plan (its structural #defines, a
// separator line "//---- program", then the generated body of "dmb_jit_program.inc"): host only, for the CPU test-suite.
// peer != 0: the variant whose stores go to the peers' shards.  Returns the bytes needed (including NUL), 0 when the step is
// not a sweep or the generator does not cover it.
int64_t dmb_jit_source(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
                       int sweep_index, int peer, char* out, size_t cap)
{
    init_options();
    try
    {
        std::string defines, program;
        if (!jit_text_of(n_qubits, world_size, gates, n_gates, mats, n_mats, sweep_index, peer, defines, program)) return 0;
        const std::string js = defines + "//---- program\n" + program;
        if (out && cap > 0)
        {
            const size_t ncopy = std::min(cap - 1, js.size());
            memcpy(out, js.data(), ncopy);
            out[ncopy] = 0;
        }
        return (int64_t)js.size() + 1;
    }
    catch (const std::invalid_argument& e)
    {
        return fail(DMB_EINVAL, e.what());
    }
    catch (const std::exception& e)
    {
        return fail(DMB_ESTATE, e.what());
    }
}

// Hands that text to the library's own run-time compiler (worker threads, memory + disk cache; compiling needs no GPU).
// wait != 0: until the kernel is built.  Returns 1 built, 0 queued / still compiling, -1 the compilation failed
// (dmb_last_error has the log), -2 no such sweep / not covered, or a DMB_E* code.
int dmb_jit_compile(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
                    int sweep_index, int peer, int wait)
{
    init_options();
    try
    {
        std::string defines, program, why;
        if (!jit_available(&why)) return fail(DMB_ESTATE, "run-time compiler unavailable: " + why);
        if (!jit_text_of(n_qubits, world_size, gates, n_gates, mats, n_mats, sweep_index, peer, defines, program)) return -2;
        JitKernel* k = jit_request(defines, program);
        if (wait) jit_wait(k);
        const int st = k->state.load();
        if (st < 0) { fail(DMB_ESTATE, "run-time compilation failed: " + k->log); return -1; }
        return st;
    }
    catch (const std::invalid_argument& e)
    {
        return fail(DMB_EINVAL, e.what());
    }
    catch (const std::exception& e)
    {
        return fail(DMB_ESTATE, e.what());
    }
}

int dmb_device_synchronize(void)
{
    CU(cudaDeviceSynchronize());
    return DMB_OK;
}

// Counters of the run-time compiler (process-wide) and of handle h's last dmb_run (h may be NULL for the process-wide ones):
// "jit_compiled", "jit_disk_hits", "jit_failed", "jit_compile_ms", "jit_ready", "jit_available";
// "jit_sweeps" (sweeps of the last run on specialised kernels), "jit_pending" (... still interpreted: kernel not ready yet).
int dmb_query(dmb_handle h, const char* name, double* value)
{
    if (!name || !value) return fail(DMB_EINVAL, "null argument");
    unsigned long long compiled = 0, disk = 0, failed = 0;
    double ms = 0;
    jit_counters(&compiled, &disk, &failed, &ms);
    if (!strcmp(name, "jit_compiled")) *value = (double)compiled;
    else if (!strcmp(name, "jit_disk_hits")) *value = (double)disk;
    else if (!strcmp(name, "jit_failed")) *value = (double)failed;
    else if (!strcmp(name, "jit_compile_ms")) *value = ms;
    else if (!strcmp(name, "jit_ready")) *value = (double)jit_ready_count();
    else if (!strcmp(name, "jit_available")) *value = jit_available(nullptr) ? 1.0 : 0.0;
    else if (h && (!strcmp(name, "jit_sweeps") || !strcmp(name, "jit_pending")))
    {
        const bool used = !strcmp(name, "jit_sweeps");
        double v = 0;
        if (is_group(h)) for (const dmb_sim* c : h->shards) v += used ? c->jit_used : c->jit_missing;
        else v = used ? h->jit_used : h->jit_missing;
        *value = v;
    }
    else return fail(DMB_EINVAL, std::string("unknown counter ") + name);
    return DMB_OK;
}

} // extern "C"
