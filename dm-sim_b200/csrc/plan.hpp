// dm-sim_b200/csrc/plan.hpp -- host-side circuit compiler of the B200 density-matrix engine.
//
// Pipeline (all host, no CUDA):  dmb_gate list
//   --expand-->  primitives on n qubits          (reference *_GATE bodies, src/dmsim_nvgpu_omp.cuh:1004-1813)
//   --fuse---->  1-/2-qubit blocks (2x2 / 4x4)   (new: the reference applies one primitive per HBM sweep)
//   --mirror-->  ops on the 2n-bit flat index: U on bit q ("L part"), conj(U) on bit q+n ("R part")
//   --schedule-> tile sweeps: each sweep stages 2^k elements (k tile bits, always including the lowest
//                physical bits so HBM runs stay contiguous) in shared memory and applies every op whose
//                bits are inside the tile; with P = 2^g GPUs the top g physical bits are the rank and an
//                exchange step swaps them with the top g local bits (qubit remap).
// Replaces circuit()/Gate::exe_op/*_OP dispatch (:816-822, :165-168, :1821-2051) and the forward /
// adjoint / backward phase structure of simulation_kernel (:918-969).
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/dmsim_b200.h"

namespace dmb
{
typedef std::complex<double> cplx;

// device-visible op classes (kernels.cu switches on these)
enum OpClass : int
{
    CLS_DENSE1 = 0, // 2x2 dense:        m[0..3]
    CLS_DIAG1 = 1,  // diag(m0, m1)
    CLS_MONO1 = 2,  // anti-diagonal:    out0 = m0*v1, out1 = m1*v0
    CLS_SRN1 = 3,   // reference SRN_GATE (:1253-1266), real-linear, not a matrix
    CLS_DENSE2 = 4, // 4x4 dense:        m[0..15]
    CLS_DIAG2 = 5,  // diag(m0..m3)
    CLS_MONO2 = 6,  // monomial:         out[r] = m[r] * v[src[r]]
    CLS_CPHASE = 7  // controlled phase diag(1, 1, 1, phi): only ONE of its bits has to be inside the tile -- the other
                    // may be any physical bit (even a rank bit), it merely selects whether the phase applies
};

// one primitive / fused block on the half circuit (logical qubits 0..n-1)
struct Block
{
    int nq = 1;      // 1 or 2
    int q[2] = {0, 0}; // nq==2: matrix index = 2*bit(q[0]) + bit(q[1])
    cplx m[16];      // row-major 2x2 (m[0..3]) or 4x4
    bool srn = false;
    int weight = 1;  // reference primitives merged into this block
};

// op inside one sweep, in tile-local bit positions
struct TileOp
{
    int cls = 0;
    int j0 = 0, j1 = 0; // tile-local bits; j0 carries the matrix MSB for 2-bit ops
    int p1 = -1;        // CLS_CPHASE with its second bit outside the tile: j1 = -1 and p1 = that PHYSICAL bit
    int nb = 1;
    cplx m[16];         // FULL matrix (2x2 or 4x4) as applied (already conjugated for R parts)
    int weight = 1;
};

struct Sweep
{
    int k = 0;                  // tile bits
    std::vector<int> in_pos;    // physical bit of tile-local bit j when loading (ascending)
    std::vector<int> out_pos;   // physical bit tile-local bit j is stored to (== in_pos unless permuting)
    bool out_of_place = false;  // reads buffer A, writes buffer B (then the buffers swap roles)
    int swz_mode = 0;           // shared-memory swizzle of the tile (devop.hpp: kSwzXor3 / kSwzTma); kSwzTma <=> the tile
                                // is moved by TMA (full-size tiles when PlanOptions::tma is on)
    std::vector<TileOp> ops;
    int weight = 0;
    int spread_top = 0;         // pack sweep of a qubit remap on 2^spread_top ranks: the top spread_top LOCAL bits of an output index
                                // select the rank the element goes to.  Those of them that are not tile bits become the LOWEST
                                // bits of the tile id, so that the CTAs running at the same time store to different peers (with
                                // the ascending enumeration they would all target the same one or two ranks for a long stretch:
                                // measured 413 GB/s per direction on 4 GPUs against 639-664 on 2 / 8, where the rank-selecting
                                // bits happened to be tile bits)
};

struct Step
{
    int kind = 0; // 0 = sweep, 1 = exchange (swap physical bits [M-g, M) with the rank bits [M, N))
    Sweep sweep;
};

struct PlanOptions
{
    int tile_bits = 12; // max k (2^12 complex FP64 = 64 KiB of shared memory)
    int low_bits = 3;   // physical bits 0..low_bits-1 are always tile bits (2^3 * 16 B = 128 B runs)
    int min_tiles_log2 = 10; // prefer >= 2^10 tiles when the state is small (keeps 148 SMs busy)
    int max_ops = 112;       // ops per sweep (the kernel keeps the sweep's op table in shared memory)
    bool hot_low = true;     // swap bits with pending work into the always-in-tile low positions (plan.cpp)
    bool move_h = false;     // second identity of rewrite_hadamard_cx (plan.cpp)
    int scan_window = 1536;  // pending ops a tile-selection scan looks at
    int max_cphase = 160;    // controlled phases per sweep (<= kMaxStarsPerSweep: each may need its own star slot)
    bool cphase = true;      // schedule controlled phases with only one bit in the tile (CLS_CPHASE)
    bool tma = true;         // full-size tiles (k == 12) are loaded / stored by TMA (128-byte hardware swizzle)
    int small_state_bits = 0;  // shards of <= 2^small_state_bits elements keep the plain tile I/O (no TMA): set to 20 (16 MiB, L2
                               // resident) by the option "persistent", whose one-launch cooperative executor needs it
    bool spread_peers = true;  // remap pack sweeps enumerate their tiles rank-selecting bits first (Sweep::spread_top)
    int tma_box_bits = 10;   // largest TMA box: 2^10 elements = 16 KiB (the rest of the tile bits: separate copies)
};

struct Plan
{
    int n = 0;      // qubits
    int g = 0;      // log2(world_size)
    std::vector<int> start_layout; // physical bit of logical bit l (size 2n)
    std::vector<int> end_layout;
    std::vector<Step> steps;
    uint64_t n_gates = 0, n_primitives = 0, n_blocks = 0, n_sweeps = 0, n_exchanges = 0;
    bool has_srn = false;
    // The stored array X relates to the state S the reference would hold by S = conj^f(X) (f = conj flag).
    // Every reference sim() ends in its transposed frame, S' = conj(X'^T) (identical to X' only for Hermitian
    // states; SRN, reference :1253-1266, breaks Hermiticity): instead of a transpose sweep the plan swaps the
    // row / column halves of end_layout and flips the flag.
    bool conj_start = false, conj_end = false;
};

// Throws std::invalid_argument on bad gates.
void expand_gates(int n_qubits, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
                  std::vector<Block>& prims);
void fuse_blocks(int n_qubits, const std::vector<Block>& prims, std::vector<Block>& blocks, bool split_cphase = true);
Plan make_plan(int n_qubits, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats,
               size_t n_mats, const std::vector<int>& start_layout, const PlanOptions& opt, bool conj_state = false,
               bool non_hermitian = false);
// extra (optional): one JSON object text per step, spliced into sweep steps as "dev": {...}
std::string plan_to_json(const Plan& p, const std::vector<std::string>* extra = nullptr);

// classify a 2x2 / 4x4 matrix; for MONO fills src[] (column of the single non-zero of each row)
int classify(int nb, const cplx* m, int* src);

const char* op_name(int op);
} // namespace dmb
