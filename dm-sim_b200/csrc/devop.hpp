// dm-sim_b200/csrc/devop.hpp -- plain-old-data shared by the host encoder (encode.cpp, system compiler) and the
// device code (kernels.cu).  No CUDA types here.
#pragma once
#ifdef __CUDACC_RTC__ // (run-time compilation of the sweep program: no host headers)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <cstdint>
#endif

namespace dmb
{
// CTA geometry (compile-time): threads per CTA and tile elements a lane keeps in registers per round.
//   DMB_THREAD_BITS = 7, DMB_REG_BITS = 4 (default): 4 warps x 3 CTAs per SM, 16 resident elements (64 registers) per
//     lane: the fewest shared-memory round trips and dispatches per op;
//   DMB_THREAD_BITS = 8, DMB_REG_BITS = 3: 8 warps x 3 CTAs = 24 warps per SM (80 registers per thread), 8 resident
//     elements per lane.  Measured on B200 (profiles/README.md, r2j): twice the warps but 60 % more instructions per
//     sweep (more rounds, dispatches and star multiplications per element): qft_n15 31.7 ms vs 24.8 ms, random_c1c2_n15
//     505 ms vs 378 ms.  Kept buildable (python dm-sim_b200/build.py with DMB_GEOM=8,3) and covered by the CPU emulators.
#ifndef DMB_THREAD_BITS
#define DMB_THREAD_BITS 7
#endif
#ifndef DMB_REG_BITS
#define DMB_REG_BITS 4
#endif
constexpr int kMaxTileBits = 12;  // 2^12 complex FP64 = 64 KiB of shared memory per CTA; 3 CTAs per SM overlap each
                                  // other's load / compute / store phases
constexpr int kThreadBits = DMB_THREAD_BITS;  // log2(kTileThreads)
constexpr int kTileThreads = 1 << kThreadBits;
constexpr int kWarpBits = kThreadBits - 5;    // log2(warps per CTA)
constexpr int kRegBits = DMB_REG_BITS;        // a lane keeps 2^kRegBits tile elements resident per round
constexpr int kRegElems = 1 << kRegBits;
constexpr int kMaxIter = 1 << (kMaxTileBits - kThreadBits); // load / store iterations per thread
constexpr int kStarW = 16;        // (warp << iteration bits) | iteration of a round: warps per CTA x iterations <= 16
static_assert(kRegBits == 3 || kRegBits == 4, "register rounds hold 8 or 16 elements per lane");
static_assert((1 << kWarpBits) * (1 << (kMaxTileBits - kThreadBits - kRegBits)) <= kStarW, "star table size");
constexpr int kMaxOpsPerSweep = 112;

// XOR swizzle of the shared-memory tile (element = 16 B): the low 3 element bits (the 16-byte bank group) are
// XORed with bits 3..5, 6..8 and 9..11, so that tile bit p moves the bank group by the unit vector e_(p mod 3):
// any three tile bits with different residues mod 3 enumerate 8 different bank groups.  The encoder gives lane bits
// 0..2 (the 8 lanes of one LDS.128 / STS.128 phase) such positions.  Linear (thread-order) access is conflict free
// for any swizzle of this form.  GF(2)-linear: swz(a ^ b) == swz(a) ^ swz(b) -- the encoder pre-swizzles every
// index contribution.
//
// Mode 1 (kSwzTma) is the hardware's 128-byte swizzle of a TMA tile (CU_TENSOR_MAP_SWIZZLE_128B on a 1024-byte aligned
// buffer: the 16-byte chunk index of a 128-byte row is XORed with the low 3 bits of the row index): only tile bits 3..5
// move the bank group.  Sweeps whose tile is loaded / stored by TMA use it; the encoder then picks lane bits 0..2 from
// three different residues among tile bits 0..5.
constexpr int kSwzXor3 = 0, kSwzTma = 1;
inline unsigned swz_host(unsigned e, int mode = kSwzXor3)
{
    return mode == kSwzTma ? e ^ ((e >> 3) & 7u) : e ^ ((e >> 3) & 7u) ^ ((e >> 6) & 7u) ^ ((e >> 9) & 7u);
}

// Register-level op codes (what the device switches on).  Two-bit ops are canonicalised by the encoder so that
// the matrix MSB sits on the HIGHER register bit; `pos` then selects one of the pairs (1,0) (2,0) (2,1).
enum RegOpCode : int32_t
{
    RC_DENSE1 = 0, // m[0..3]
    RC_DIAG1 = 1,  // host-side only: folded into RC_DIAGR by the encoder
    RC_MONO1 = 2,  // out0 = m0*v1, out1 = m1*v0 (aux bit 12: unit phases)
    RC_SRN1 = 3,   // reference SRN_GATE (:1253-1266)
    RC_DENSE2 = 4, // m[0..15]
    RC_DIAG2 = 5,  // host-side only: folded into RC_DIAGR by the encoder
    RC_PERM2 = 6,  // monomial with a row permutation from {CX(msb ctrl), CX(lsb ctrl), SWAP}: aux bits 0..1 = which,
                   // m[0..3] = row phases, aux bit 12: unit phases
    RC_DIAGR = 7,  // product of consecutive diagonal ops of a round: v[c] *= m[c] for the 16 registers, skip mask in
                   // aux bits 0..15 (entries equal to 1)
    RC_DENSE1_RR = 8, // 2x2 with four REAL entries [[d0,d1],[d2,d3]] (H, RY, ...), |d0| not small, in the pivoted in-place
                      // form m[0..3] = {d0, d1, d2/d0, det/d0}: half the FP64 work of RC_DENSE1 and no register copies
    RC_DENSE1_RI = 9, // [[d0, i d1], [i d2, d3]] with real d (RX, W, ...), same form with det = d0 d3 + d1 d2
    RC_HAD = 11,      // unscaled butterflies [[1, 1], [1, -1]] (a' = a + b, b' = a' - 2 b: 2 FP64 instructions per pair and
                      // component, in place, no payload) on every register bit of the mask in aux bits 0..3; the
                      // 1/sqrt(2) factors of the H gates of a SWEEP are folded by the encoder into the payload of a dense
                      // op of the same sweep (scalars commute with everything)
    RC_DIAGP = 12,    // diagonal whose non-unit entries all have register bit `pos` set (phases controlled by that
                      // bit, e.g. the controlled phases between the register bits of a QFT round): m[j] multiplies the
                      // element whose other three register bits spell j; skip mask (unit entries) in aux bits 0..7
    RC_CP2 = 13,      // one controlled phase between two register bits: the 4 elements with both bits set are multiplied
                      // by m[0]; `pos` as for the other two-bit ops
    RC_QFT2 = 14,     // butterfly on the LOWER register bit of the pair, controlled phase m[0] between the two, butterfly on
                      // the HIGHER bit: the radix-4 step of a QFT round as ONE dispatch (`pos` as for the two-bit ops)
    RC_DENSE2_LU = 15, // dense 4x4 as M = L U (unit lower, upper; no pivoting -- the encoder only emits it when the factors
                       // stay small): U is applied top-down and L bottom-up IN PLACE, the same 16 complex multiply-adds as
                       // the direct form but no copies of the inputs (the direct form spent 11 % of its instructions on register
                       // moves).  m[0..9] = U row by row (u00 u01 u02 u03 u11 u12 u13 u22 u23 u33), m[10..15] = l10 l20 l21 l30 l31 l32
    RC_STAR = 10      // controlled-phase star: for every register bit p in aux bits 0..3, the elements with that bit set
                      // are multiplied by  L_p[lane] * WO_p[warp, iteration]  (DevStar slot star[p]): the product of the
                      // phases of all controlled-phase ops between register bit p and the partner bits that are set in
                      // this lane's / warp's / tile's index
};

// One (star op, register bit) pair.  l[] and w[] are host-computed partial products over the partner bits that are
// lane bits resp. iteration / warp bits of the round; the partners outside the tile (any physical bit, rank bits
// included) are multiplied in once per tile by the kernel's prologue (WO = w * prod phi[j] over set bits).
constexpr int kMaxStarOut = 28;
struct alignas(16) DevStar
{
    double w[2 * kStarW]; // kStarW complex: index = (warp << iteration bits) | iteration
    double la[16]; // lane part, factored so that it fits shared memory: L[lane] = la[lane & 7] * lb[lane >> 3]
    double lb[8];
    int32_t n_out;
    int32_t pad[3];
    int32_t bit[kMaxStarOut];    // physical bit of the full index (>= M: rank bits)
    double phi[2 * kMaxStarOut]; // (re, im)
};
static_assert(sizeof(DevStar) == 768 + 16 * kStarW, "DevStar layout");
constexpr int kMaxStarsPerSweep = 160; // shared memory of each: WO[kStarW] (rebuilt per tile) | L[32] = la x lb (once)
constexpr int kStarSmemBytes = 16 * (kStarW + 32);

// Device op stream: 16-byte header + payload (the used part of DevOp::m), 16-byte granularity; a zero header ends it.
struct alignas(16) DevOpHdr
{
    int32_t vid;    // dev_vid(code, pos, aux)
    int32_t aux;
    int32_t size16; // header + payload in 16-byte units
    int32_t star[1]; // RC_STAR: star[0] = first DevStar slot of this op (one per set aux bit, ascending)
};
// The kernel's jump-table index of an op: dense over (code, position) -- and over the register-bit MASK for RC_HAD and
// RC_STAR, so that those two need no header read at all.
constexpr int kVidDense2 = 0, kVidPerm2 = 6, kVidCp2 = 12, kVidDense1 = 18, kVidRR = 22, kVidRI = 26, kVidMono1 = 30,
              kVidSrn1 = 34, kVidDiagP = 38, kVidDiagR = 42, kVidQft2 = 43, kVidLu2 = 49, kVidHad = 55, kVidStar = 70,
              kNumVids = 85;
constexpr int dev_vid(int code, int pos, int aux)
{
    switch (code)
    {
    case RC_DENSE2: return kVidDense2 + pos;
    case RC_PERM2: return kVidPerm2 + pos;
    case RC_CP2: return kVidCp2 + pos;
    case RC_DENSE1: return kVidDense1 + pos;
    case RC_DENSE1_RR: return kVidRR + pos;
    case RC_DENSE1_RI: return kVidRI + pos;
    case RC_MONO1: return kVidMono1 + pos;
    case RC_SRN1: return kVidSrn1 + pos;
    case RC_DIAGP: return kVidDiagP + pos;
    case RC_DIAGR: return kVidDiagR;
    case RC_QFT2: return kVidQft2 + pos;
    case RC_DENSE2_LU: return kVidLu2 + pos;
    case RC_HAD: return kVidHad + (aux & 15) - 1;
    case RC_STAR: return kVidStar + (aux & 15) - 1;
    default: return kNumVids;
    }
}
constexpr int dev_op_payload_bytes(int code)
{
    switch (code)
    {
    case RC_DENSE1: return 64;
    case RC_MONO1: return 32;
    case RC_DENSE2: return 256;
    case RC_DENSE2_LU: return 256;
    case RC_PERM2: return 64;
    case RC_DIAGR: return 16 * kRegElems;
    case RC_DIAGP: return 8 * kRegElems;
    case RC_CP2: return 16;
    case RC_QFT2: return 16;
    case RC_DENSE1_RR: return 32;
    case RC_DENSE1_RI: return 32;
    default: return 0;
    }
}

// One op of a round, applied to the lane's 2^kRegBits resident elements (host-side description; the device reads the
// compact stream of DevOpHdr + payload built from it).
struct alignas(16) DevOp
{
    int32_t code; // RegOpCode
    int32_t aux;
    int32_t pos;  // 1-bit ops: register bit 0..3; 2-bit ops: (msb, lsb) = (1,0) (2,0) (2,1) (3,0) (3,1) (3,2) -> 0..5
    int32_t vid;  // host side: RC_STAR: first DevStar slot (the kernel's jump-table index is dev_vid(code, pos, aux))
    double m[32]; // up to 16 complex entries (re, im)
};
static_assert(sizeof(DevOp) == 272, "DevOp layout");

// A round: the lane loads the 2^kRegBits elements  base ^ roff[c]  (base = lane_tab[lane] ^ iter_tab[iter] ^
// wtab[warp], everything pre-swizzled), applies ops [first, first+count) in registers and stores them back:
// ONE shared-memory round trip for `count` ops.  lane_tab / iter_tab / roff / DevGroup::wtab hold BYTE offsets into the
// tile (element index * 16 <= 65520): the XOR of the four is the address, no scaling on the device.
constexpr int kMaxOpsPerRound = 16; // the encoder splits longer rounds (same tables, one more shared-memory round trip)
struct alignas(16) DevRound
{
    int32_t first, count; // first: offset of the round's first op in the op stream, in 16-byte units
    int32_t n_iter;   // iterations per lane
    int32_t n_active; // active lanes (32 unless the sub-tile has fewer work items)
    uint16_t lane_tab[32];
    uint16_t iter_tab[8];
    uint16_t roff[16];
    uint8_t vids[kMaxOpsPerRound]; // dev_vid() of the round's ops: the kernel dispatches from this packed list (in
                                   // registers) and every op body advances the stream pointer by its static size, so
                                   // that no shared-memory load sits on the dispatch path
    int32_t star0;                 // DevStar slot of the round's first star bit (the following ones are consecutive)
    int32_t pad[3];
};
static_assert(sizeof(DevRound) == 160, "DevRound layout");

// A run of consecutive rounds that leave kWarpBits tile bits untouched: warp w owns the sub-tile where those bits
// equal w and runs the whole group with __syncwarp() only; CTA barriers happen between groups.
struct alignas(16) DevGroup
{
    int32_t first, count; // rounds [first, first+count) of the sweep
    int32_t n_warps;      // 2^kWarpBits, or 1 when the tile is too small to split
    int32_t pad;
    uint16_t wtab[16];    // swizzled tile-index contribution of the warp
};
static_assert(sizeof(DevGroup) == 48, "DevGroup layout");

// A tile as TMA boxes.  The shard is a dense 5-D tensor of FP64 pairs: dimension d covers the physical bits
// [start[d], start[d] + span[d]) of the element index, its low box_log2[d] bits are tile bits (the box), the bits above them
// are not (they are part of the box's start coordinate).  Dimension 0 is always the 128-byte run (physical bits 0..2).
// The tile bits that do not fit the box (at most 5 dimensions, <= 256 elements each, <= max box size) are enumerated by
// n_copies = 2^n_enum separate copies; copy j lands at byte offset j * box_bytes of the tile (ascending tile-local order).
struct TmaGeom
{
    int32_t n_copies, box_bytes;
    unsigned char start[5], span[5], box_log2[5];
    unsigned char n_enum;
    unsigned long long enum_off[32]; // element offset of copy j (its enumerated tile bits deposited at their positions)
};
struct alignas(64) TmaDesc // CUtensorMap (opaque, 128 bytes, filled by cuTensorMapEncodeTiled)
{
    unsigned long long opaque[16];
};

// Direct store of the LAST round (full-size TMA-loaded tiles that do not permute bits): the lanes write the round's results
// straight from registers to global memory (streaming 128-bit stores; lane bits 0..2 are the tile's 128-byte run, so every
// quarter warp writes one full line) instead of going back through shared memory and a TMA store.  The tile buffer is
// dead as soon as every warp has LOADED its last-round elements: the next tile's TMA load is issued then (per half when
// the round's iteration bit is one of the TMA copy-enumeration bits), so that it lands while the last round computes
// and nobody waits for a store to drain.
struct DevDirect
{
    int32_t enabled;
    int32_t half_enum;            // the round's iteration bit is enumeration bit `half_enum` of the TMA copies (-1: it is not)
    unsigned char reg_pos[4];     // physical bit of register bit i of the last round
    unsigned char lane_pos[5];    // ... of lane bit i
    unsigned char warp_pos[3];    // ... of warp bit i
    unsigned char iter_pos[4];    // ... of iteration bit i
    unsigned long long reg_off[16]; // BYTE offset of register element c (its register bits deposited at reg_pos[])
    unsigned long long iter_off[8]; // ELEMENT offset of iteration it (its bits deposited at iter_pos[])
};

// Kernel parameter block of one sweep (by value -> constant bank; the per-iteration tables become immediate
// constant operands of the unrolled load / store loops).
struct SweepArgs
{
    const void* in;   // double2*
    void* out;        // double2*
    const void* ops;  // device copies: op stream (DevOpHdr + payload ...)
    const DevRound* rounds;
    const DevGroup* groups;
    const DevStar* stars;
    int ops_bytes, n_rounds, n_groups, n_stars;
    unsigned long long rank_bits; // rank << M: the index bits above the shard
    // fused qubit-remap exchange (peer_shift >= 0): the store phase writes element `off` of the OUTPUT frame straight
    // into the shard of rank p = off >> peer_shift, at (my rank << peer_shift) | (off & low mask), over NVLink
    int peer_shift;
    int peer_rank;
    unsigned long long peer_out[8]; // double2* of every rank's destination buffer (IPC-mapped; own pointer for self)
    unsigned op_mask;           // bit c set <=> some op of the sweep has RegOpCode c (selects the kernel instantiation)
    int k;                      // tile bits
    int n_comp;                 // M - k
    unsigned long long n_tiles; // 2^(M-k)
    unsigned long long hin[kMaxIter];  // BYTE offset contributed by iteration `it` when loading
    unsigned long long hout[kMaxIter]; // ... when storing
    unsigned short hs[kMaxIter];       // swizzled smem index contributed by iteration `it` when storing
    unsigned char gin[12];       // physical bit of loop bit i (< kThreadBits) when loading
    unsigned char gout[12];      // ... when storing
    unsigned char sout[12];      // tile-local bit of loop bit i (< kThreadBits) when storing
    unsigned char cin[40];       // physical bits enumerated by the tile id when loading (ascending)
    unsigned char cout[40];      // ... when storing
    // TMA tile I/O (full-size tiles): swz_mode = kSwzTma, the tile is loaded (tma_load) / stored (tma_store: in-place
    // sweeps that do not permute bits) as boxes of ONE tensor map over the shard buffer
    int swz_mode, tma_load, tma_store;
    int tma_prefetch; // thread 0 also prefetches the CTA's next tile into L2 when it issues a tile's loads
    // tile id -> element offset of the tile (the deposit of the id's bits at cin[] / cout[]), as three 7-bit lookups:
    // base = base_in[0][id & 127] | base_in[1][(id >> 7) & 127] | base_in[2][id >> 14]
    unsigned long long base_in[3][128], base_out[3][128];
    DevDirect direct;
    TmaGeom tma;
    TmaDesc tmap_in, tmap_out;
};
} // namespace dmb
