// dm-sim_b200/csrc/devop.hpp -- plain-old-data shared by the host encoder (encode.cpp, system compiler) and the
// device code (kernels.cu).  No CUDA types here.
#pragma once
#include <cstdint>

namespace dmb
{
constexpr int kMaxTileBits = 12;  // 2^12 complex FP64 = 64 KiB of shared memory per CTA
constexpr int kTileThreads = 256; // 8 warps
constexpr int kWarpBits = 3;      // log2(warps per CTA)
constexpr int kMaxOpsPerSweep = 112;

// XOR swizzle of the shared-memory tile (element = 16 B): low 3 element bits ^= bits 3..5, so that 8 lanes of an
// LDS.128 phase hit 8 different 16-byte bank groups both for unit-stride and for stride-2/4/8 element patterns.
// GF(2)-linear: swz(a ^ b) == swz(a) ^ swz(b) -- the encoder pre-swizzles every index contribution.
inline unsigned swz_host(unsigned e) { return e ^ ((e >> 3) & 7u); }

// One op of a sweep as the device sees it (368 bytes).  Work items (pairs for 1-bit ops, quads for 2-bit ops) of
// a warp's sub-tile are enumerated as item = lane + 32*iter; the tile index of member c of an item is
//     lane_tab[lane] ^ iter_tab[iter] ^ wtab[warp] ^ off[c]          (all already swizzled)
struct alignas(16) DevOp
{
    int32_t cls;      // OpClass
    int32_t aux;      // MONO2: src[r] in bits 2r..2r+1; skip-row mask in bits 8..11; unit-phase flag bit 12
    int32_t n_iter;   // iterations per lane
    int32_t n_active; // active lanes (32 unless the sub-tile has fewer items)
    double m[32];     // up to 16 complex entries (re, im); layout per class as in plan.hpp
    uint16_t lane_tab[32];
    uint16_t iter_tab[8];
    uint16_t off[4];
    uint16_t pad[4];
};
static_assert(sizeof(DevOp) == 368, "DevOp layout");

// A run of consecutive ops that leave kWarpBits tile bits untouched: warp w owns the sub-tile where those bits
// equal w and runs the whole group with __syncwarp() only; CTA barriers happen between groups.
struct alignas(16) DevGroup
{
    int32_t first, count; // ops [first, first+count) of the sweep
    int32_t n_warps;      // 8, or 1 when the tile is too small to split
    int32_t pad;
    uint16_t wtab[8];     // swizzled tile-index contribution of the warp
};
static_assert(sizeof(DevGroup) == 32, "DevGroup layout");

// Kernel parameter block of one sweep (by value -> constant bank; the per-iteration tables become immediate
// constant operands of the unrolled load / store loops).
struct SweepArgs
{
    const void* in;   // double2*
    void* out;        // double2*
    const DevOp* ops; // device copies
    const DevGroup* groups;
    int n_ops, n_groups;
    int k;                      // tile bits
    int n_comp;                 // M - k
    unsigned long long n_tiles; // 2^(M-k)
    unsigned long long hin[16];  // element offset contributed by iteration `it` when loading
    unsigned long long hout[16]; // ... when storing
    unsigned short hs[16];       // swizzled smem index contributed by iteration `it` when storing
    unsigned char gin[8];        // physical bit of loop bit i (< 8) when loading
    unsigned char gout[8];       // ... when storing
    unsigned char sout[8];       // tile-local bit of loop bit i (< 8) when storing
    unsigned char cin[40];       // physical bits enumerated by the tile id when loading (ascending)
    unsigned char cout[40];      // ... when storing
};
} // namespace dmb
