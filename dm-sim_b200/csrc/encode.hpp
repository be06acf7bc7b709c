// dm-sim_b200/csrc/encode.hpp -- host encoder of planned sweeps into device tables (see encode.cpp).
#pragma once
#include <string>
#include <vector>

#include "devop.hpp"
#include "plan.hpp"

namespace dmb
{
struct EncodedSweep
{
    std::vector<DevOp> ops;             // host-side description (JSON / tests); vid of RC_STAR = first star slot
    std::vector<unsigned char> stream;  // what the device reads: DevOpHdr + payload per op, zero header at the end
    std::vector<DevRound> rounds;       // first = offset into `stream` in 16-byte units
    std::vector<DevGroup> groups;
    std::vector<DevStar> stars;
    DevDirect direct;                   // direct store of the last round (enabled = 0: not applicable)
    unsigned op_mask = 0;               // register-op codes present
    unsigned long long fp64_per_lane = 0; // FP64 pipe slots (DFMA / DMUL / DADD) per lane and iteration (kRegElems elements)
};
void encode_sweep(const Sweep& sw, EncodedSweep& out);
// fills k, n_comp, n_tiles and every address table of `a` (pointers / counts are the caller's job)
void fill_sweep_tables(const Sweep& sw, int M, SweepArgs& a);
// largest TMA box in tile bits (2^bits elements of 16 bytes; default 10 = 16 KiB)
void set_sweep_tma_box_bits(int bits);
int sweep_tma_box_bits();
// last round of full-size in-place TMA tiles stores straight from registers to global memory (default on)
void set_sweep_direct_store(bool on);
// a sweep whose last round would be a leftover runs that leftover FIRST (default off: measured neutral)
void set_sweep_light_first(bool on);
// rounds planned from the end of the op list when that ends the sweep on a heavier round (default off: measured slower)
void set_sweep_heavy_last(bool on);
// dense 4x4 ops as in-place L U (RC_DENSE2_LU) where the factors stay small (default on)
void set_sweep_dense2_lu(bool on);
bool sweep_dense2_lu();
// the tile-id -> tile-base lookup tables of `a` from its (possibly reduced: sparse start) cin / cout / n_comp
void fill_base_tables(SweepArgs& a);
// JSON text of the device tables of one sweep (for the CPU test-suite's kernel-indexing emulator)
std::string encoded_to_json(const EncodedSweep& e, const SweepArgs& a);
} // namespace dmb
