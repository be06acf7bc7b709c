/* include/dmsim_b200.h -- the C-ABI of the B200-native density-matrix gate engine.
 *
 * This is the drop-in boundary for the GPU backend of pnnl/DM-Sim: plain C, plain pointers and sizes,
 * no torch / pybind types.  Every entry point names the reference interface it replaces
 * (reference paths are relative to the DM-Sim tree, `src/dmsim_nvgpu_omp.cuh` unless noted).
 * The C++ classes DMSim::Gate / DMSim::Simulation (include/dmsim_b200.hpp), the pybind11 module
 * libdmsim_py_nvgpu_omp and the Python host mirror (dm-sim_b200/) are all thin layers over it.
 *
 * Conventions
 *  - All functions return DMB_OK (0) or a negative DMB_E* code; dmb_last_error() gives the text.
 *  - The state is sharded on the top g = log2(world_size) bits of the 2n-bit PHYSICAL index: rank r
 *    holds the flat indices whose top g bits equal r.  world_size == 1 is the single-GPU engine.
 *    Two ways to drive P = 2^g GPUs of one node:
 *      (a) ONE process, ONE handle (the reference's Simulation(n_qubits, n_gpus), :196-271):
 *          dmb_create(n, P, DMB_ALL_RANKS, first_device, &h) owns all P shards on devices
 *          [first_device, first_device + P); the shards reach each other through peer access
 *          (cudaDeviceEnablePeerAccess, no IPC, no NCCL), every call below works on the whole state.
 *      (b) one process per GPU (torchrun / MPI style): dmb_create(n, P, rank, device, &h) owns ONE
 *          shard; dmb_comm_* below attach the NCCL communicator (and optionally the peers' buffers).
 *    The only cross-rank step of a run is the qubit-remap exchange.
 *  - State layout in HBM: interleaved complex FP64 (re,im), flat logical index = col*dim + row with
 *    qubit q = bit q of row (reference :989-998), i.e. a 2n-bit vector on which a gate U on qubit q
 *    acts as U on bit q and conj(U) on bit q+n.  What is stored is rho^T, as in the reference.
 *    A tracked bit permutation maps logical bits to physical bits (identity on one GPU).
 */
#ifndef DMSIM_B200_H
#define DMSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_OK 0
#define DMB_EINVAL (-1)   /* bad argument (qubit out of range, bad op, ...) */
#define DMB_ECUDA (-2)    /* CUDA runtime / driver error, or no usable GPU */
#define DMB_ESTATE (-3)   /* call sequence error (run before set_circuit, ...) */
#define DMB_ENOMEM (-4)
#define DMB_ECOMM (-5)    /* multi-GPU exchange needed but no communicator attached */

#define DMB_ALL_RANKS (-1) /* rank argument of dmb_create: this handle drives all world_size GPUs from one process */

/* enum OP, reference :42-48 (order matters: xacc/DmSimApi.hpp:6-45 mirrors it).
 * DMB_OP_C1 / DMB_OP_C2 expose the reference's C1_GATE (:1004-1025) and C2_GATE (:1028-1122), which
 * have no enum value there; they are appended far after RYY so nothing is renumbered. */
enum dmb_op
{
    DMB_OP_U3, DMB_OP_U2, DMB_OP_U1, DMB_OP_CX, DMB_OP_ID, DMB_OP_X, DMB_OP_Y, DMB_OP_Z, DMB_OP_H, DMB_OP_S,
    DMB_OP_SDG, DMB_OP_T, DMB_OP_TDG, DMB_OP_RX, DMB_OP_RY, DMB_OP_RZ, DMB_OP_CZ, DMB_OP_CY, DMB_OP_SWAP, DMB_OP_CH,
    DMB_OP_CCX, DMB_OP_CSWAP, DMB_OP_CRX, DMB_OP_CRY, DMB_OP_CRZ, DMB_OP_CU1, DMB_OP_CU3, DMB_OP_RXX, DMB_OP_RZZ, DMB_OP_RCCX,
    DMB_OP_RC3X, DMB_OP_C3X, DMB_OP_C3SQRTX, DMB_OP_C4X, DMB_OP_R, DMB_OP_SRN, DMB_OP_W, DMB_OP_RYY,
    DMB_OP_COUNT,
    DMB_OP_C1 = 100, /* qb[0]; matrix #mat: 2x2 complex row-major */
    DMB_OP_C2 = 101  /* qb[0]=qubit1, qb[1]=qubit2; matrix #mat: 4x4, index = 2*bit(qubit1)+bit(qubit2) */
};

/* POD mirror of class Gate (reference :99-191) without the device function pointer. 56 bytes. */
typedef struct dmb_gate
{
    int32_t op;            /* enum dmb_op */
    int32_t qb[5];         /* qb0..qb4 */
    double theta, phi, lambda;
    int64_t mat;           /* C1/C2 only: index into the matrix table (32 doubles per slot, (re,im) pairs) */
} dmb_gate;

/* Per-run figures; replaces the reference's printf summary line (:484-490). */
typedef struct dmb_stats
{
    double sim_ms;         /* device time of the whole run (CUDA events), = reference "sim:" */
    double comm_ms;        /* device time inside qubit-remap exchanges, = reference "comm:" */
    double comp_ms;        /* sim_ms - comm_ms */
    uint64_t n_gates;      /* Gate objects in the circuit (reference "ngates") */
    uint64_t n_primitives; /* after expansion of composites (Appendix A.3 of SURVEY.md) */
    uint64_t n_blocks;     /* fused 1-/2-qubit blocks per side */
    uint64_t n_sweeps;     /* HBM passes (tile-kernel launches) executed */
    uint64_t n_exchanges;  /* all-to-all qubit remaps executed */
    uint64_t n_launches;   /* kernels launched by this run */
    uint64_t sweep_bytes;  /* algorithmic bytes of one sweep of the local shard = 32 * 4^n / P */
    uint64_t exchange_bytes; /* bytes sent per rank over NVLink by this run */
    uint64_t h2d_bytes;    /* device op tables copied host -> device by the dmb_set_circuit this run executes */
    uint64_t fp64_ops;     /* algorithmic FP64 instructions (DFMA / DMUL / DADD, one pipe slot each) per GPU of this run:
                              what the register-level ops of all sweeps spend on the local shard */
} dmb_stats;

typedef struct dmb_sim* dmb_handle;

/* ---- life cycle: replaces Simulation::Simulation / ~Simulation (:196-301) ---- */
/* device < 0: current CUDA device.  Allocates the shard (16 * 4^n / world_size bytes; a second buffer
 * of the same size is allocated lazily the first time an out-of-place step needs one).
 * rank == DMB_ALL_RANKS with world_size > 1: a single-process group on devices [max(device, 0), +world_size)
 * (needs peer access between every pair, as the reference does, :266-269; both buffers are allocated up front). */
int dmb_create(int n_qubits, int world_size, int rank, int device, dmb_handle* out);
int dmb_destroy(dmb_handle h);
/* reset_dm (:308-329): rho[0][0] = 1, identity bit layout. */
int dmb_reset_dm(dmb_handle h);
/* Load an arbitrary state (split real/imag host arrays of 4^n doubles, [col][row] like dm_real_res): every shard
 * keeps the elements it owns.  Not in the reference (its state is only reachable by running gates). */
int dmb_set_dm(dmb_handle h, const double* real, const double* imag);

/* ---- circuit: replaces append/upload/clear_circuit (:331-377, :496-520) ----
 * Takes the whole gate list (host memory), validates it (qubit range as append()'s asserts; distinct
 * qubits for multi-qubit ops), expands composites (reference :1493-1813), fuses, schedules the tile
 * sweeps and copies the device op tables (the single H2D of the step).  mats may be NULL when no
 * C1/C2 op is present. */
int dmb_set_circuit(dmb_handle h, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats);
int dmb_clear_circuit(dmb_handle h);

/* ---- execution: replaces Simulation::sim (:390-494) and simulation_kernel (:918-969) ----
 * Blocking.  State continues from the previous run (GPU-backend semantics). stats may be NULL. */
int dmb_run(dmb_handle h, dmb_stats* stats);

/* ---- results: replaces the D2H of dm_real_res / dm_imag_res (:458-466) and measure (:521-549) ----
 * Results of a sharded state are GLOBAL: a group handle gathers over its shards; one rank of a one-process-per-GPU
 * job answers through the communicator (dmb_comm_init): the call is then COLLECTIVE -- every rank makes it and every
 * rank receives the full answer (the reference's MPI flavour does the same, src/dmsim_nvgpu_mpi.cuh:472-520).
 * Without a communicator a rank reports its own part (zeros for what other ranks own; the sum over ranks is the
 * answer) where that makes sense (dmb_get_elements, dmb_get_diag, dmb_trace, dmb_purity) and fails otherwise. */
/* Full matrix into split host arrays (4^n doubles each, [col][row], holds rho^T). */
int dmb_get_dm(dmb_handle h, double* real, double* imag);
/* n arbitrary elements dm_real_res[f] / dm_imag_res[f], f = col*dim + row (spot checks of states too large to copy
 * back). */
int dmb_get_elements(dmb_handle h, const uint64_t* flat_index, size_t n, double* real, double* imag);
/* Real part of the diagonal (2^n doubles), dm_real_res[i*dim+i]. */
int dmb_get_diag(dmb_handle h, double* diag);
/* sum_i Re rho_ii and sum_ij |rho_ij|^2. */
int dmb_trace(dmb_handle h, double* trace);
int dmb_purity(dmb_handle h, double* purity);
/* measure(): p_i = |Re rho_ii|, prefix sum, one basis index per uniform number r[i] in [0,1]
 * (index j with scan[j] <= r < scan[j+1], else 0 -- exactly the reference's rule :536-543).
 * total (may be NULL) receives scan[dim]. */
int dmb_sample(dmb_handle h, const double* r, size_t n, uint64_t* out, double* total);
/* measure() with the reference's generator: srand(seed); r = rand()/RAND_MAX per shot (:534-539). */
int dmb_measure(dmb_handle h, unsigned seed, size_t repetition, uint64_t* out, double* total);

/* ---- multi-GPU: replaces the peer-copy all-to-all (:423-441) and packing/unpacking (:858-916) ----
 * One process per GPU.  Rank 0 calls dmb_comm_unique_id(), the host plumbing (torch.distributed, MPI,
 * a file...) broadcasts the 128 bytes, every rank calls dmb_comm_init().  NCCL is dlopen'ed
 * (libnccl.so.2), so single-GPU users need no NCCL at all. */
int dmb_comm_unique_id(uint8_t id[128]);
int dmb_comm_init(dmb_handle h, const uint8_t id[128]);
/* Optional, ranks of ONE node (NVLink / NVSwitch): peer-memory exchange.  Every rank exports the IPC handles of its two
 * shard buffers (2 x 64 bytes), the host plumbing all-gathers them (world_size x 128 bytes, rank order) and every rank
 * imports the lot after dmb_comm_init.  From then on a qubit remap is ONE kernel: the permuting sweep stores straight
 * into the destination ranks' shards over NVLink (pack + all-to-all fused), followed by a one-element all-reduce as
 * the cross-GPU barrier.  Without it the remap is pack sweep + ncclSend/ncclRecv. */
int dmb_comm_export(dmb_handle h, uint8_t out[128]);
int dmb_comm_import(dmb_handle h, const uint8_t* all_handles);
/* Switches the peer-memory form of the remap on (only after a successful dmb_comm_import) or off.  EVERY rank must use
 * the same form: a host that could not import the handles on some rank calls dmb_comm_p2p(h, 0) on all of them. */
int dmb_comm_p2p(dmb_handle h, int enable);
/* Raw shard + layout, for tests and host-side gathers: copies the local shard (interleaved complex,
 * 2 * 4^n / world_size doubles, PHYSICAL order; a group handle: all shards back to back, 2 * 4^n doubles) and the
 * logical->physical bit map (2n ints). */
int dmb_get_shard(dmb_handle h, double* interleaved, int32_t* phys_of_logical);

/* ---- planner introspection (host only, needs no GPU): used by the CPU test-suite ----
 * Plans a circuit exactly as dmb_set_circuit would for (n_qubits, world_size) and writes a JSON
 * description of the sweeps / exchanges (tile bit positions, fused op matrices, final bit layout).
 * conj_state: bit 0 = the stored array is the conjugate of the state, bit 1 = the state may be non-Hermitian
 * (both 0 after dmb_reset_dm).  Returns the number of bytes needed (including NUL); writes at most cap bytes. */
int64_t dmb_plan_json(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats,
                      size_t n_mats, const int32_t* start_layout /* NULL = identity */, int conj_state, char* out,
                      size_t cap);

/* ---- run-time specialised sweep kernels (csrc/jit.cu) ----
 * dmb_jit_source: the CUDA text the run-time compiler is given for sweep `sweep_index` of the plan of this circuit (host
 * only, needs no GPU; for the CPU test-suite).  Returns the bytes needed (including NUL), 0 if there is no such sweep or
 * the generator does not cover it; writes at most cap bytes.
 * dmb_query: counters -- process-wide "jit_available", "jit_compiled", "jit_disk_hits", "jit_failed", "jit_compile_ms",
 * "jit_ready"; of handle h's last dmb_run "jit_sweeps" (sweeps that ran on specialised kernels) and "jit_pending"
 * (sweeps that were still interpreted because their kernel was not compiled yet). */
int64_t dmb_jit_source(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats,
                       size_t n_mats, int sweep_index, int peer, char* out, size_t cap);
/* dmb_jit_compile: hands that text to the library's run-time compiler (worker threads, memory + disk cache; compiling
 * needs no GPU).  Returns 1 built, 0 queued or still compiling (wait == 0), -1 compilation failed, -2 no such sweep. */
int dmb_jit_compile(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
                    int sweep_index, int peer, int wait);
int dmb_query(dmb_handle h /* may be NULL */, const char* name, double* value);

/* Tuning knobs (process-wide; also read once from env DMB_TILE_BITS, DMB_LOW_BITS, DMB_GRAPH):
 * "tile_bits" (<= 12), "low_bits" (contiguous run = 2^low_bits elements), "graph" (0/1). */
int dmb_set_option(const char* name, int64_t value);

/* cudaDeviceSynchronize() on the current device (for host-only users of the C++ header: gpu_timer). */
int dmb_device_synchronize(void);

const char* dmb_last_error(void);
const char* dmb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DMSIM_B200_H */
