// dm-sim_b200/csrc/encode.cpp -- turns a planned Sweep into what sweep_kernel consumes: warp groups, register
// rounds with pre-swizzled index tables, canonicalised register-level ops, and the load/store address tables of
// the kernel parameter block.  Pure host code (system compiler).
#include "encode.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <sstream>

namespace dmb
{
static int g_tma_box_bits = 10;
static bool g_direct_store = true;
void set_sweep_direct_store(bool on) { g_direct_store = on; }
static bool g_dense2_lu = true;
void set_sweep_dense2_lu(bool on) { g_dense2_lu = on; }
bool sweep_dense2_lu() { return g_dense2_lu; }
static bool g_light_first = false; // (measured neutral on B200, r2x: qft_n15 21.85 vs 21.83 ms, bv_n15 22.75 vs 22.44 -- the sweeps that run at 0.86 of the HBM peak differ from the ones at 0.69 by their 1-KiB instead of 128-byte DRAM runs, not by the length of their last round)
void set_sweep_light_first(bool on) { g_light_first = on; }
static bool g_heavy_last = false; // (measured on B200, r2u: qft_n15 23.1 vs 21.6 ms -- the diagonals of a round planned from the back cannot be deferred and merged; random_c1c2_n15 344 vs 347 ms)
void set_sweep_heavy_last(bool on) { g_heavy_last = on; }
void set_sweep_tma_box_bits(int bits) { g_tma_box_bits = bits < 3 ? 3 : (bits > kMaxTileBits ? kMaxTileBits : bits); }
int sweep_tma_box_bits() { return g_tma_box_bits; }
namespace
{
unsigned deposit(unsigned v, const std::vector<int>& pos)
{
    unsigned r = 0;
    for (size_t i = 0; i < pos.size(); i++) r |= ((v >> i) & 1u) << pos[i];
    return r;
}

void put(DevOp& d, int i, cplx v)
{
    d.m[2 * i] = v.real();
    d.m[2 * i + 1] = v.imag();
}

// same 4x4 operator with the roles of its two bits exchanged (index bit swap)
void swap_roles(const cplx* m, cplx* o)
{
    static const int p[4] = {0, 2, 1, 3};
    cplx t[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) t[p[r] * 4 + p[c]] = m[r * 4 + c];
    memcpy(o, t, sizeof(t));
}

// register-level op from a tile op whose bits sit on register bits p0 (matrix MSB) and p1
DevOp make_reg_op(const TileOp& t, int p0, int p1)
{
    DevOp d;
    memset(&d, 0, sizeof(d));
    const cplx one(1.0, 0.0);
    if (t.cls == CLS_SRN1)
    {
        d.code = RC_SRN1;
        d.pos = p0;
        return d;
    }
    if (t.nb == 1)
    {
        d.pos = p0;
        const int cls = classify(1, t.m, nullptr);
        if (cls == CLS_DIAG1)
        {
            d.code = RC_DIAG1;
            int skip = 0;
            for (int r = 0; r < 2; r++)
            {
                put(d, r, t.m[r * 3]);
                if (t.m[r * 3] == one) skip |= 1 << r;
            }
            d.aux = skip << 8;
        }
        else if (cls == CLS_MONO1)
        {
            d.code = RC_MONO1;
            put(d, 0, t.m[1]);
            put(d, 1, t.m[2]);
            d.aux = ((t.m[1] == one && t.m[2] == one) ? 1 : 0) << 12;
        }
        else
        {
            const bool rr = t.m[0].imag() == 0 && t.m[1].imag() == 0 && t.m[2].imag() == 0 && t.m[3].imag() == 0;
            const bool ri = t.m[0].imag() == 0 && t.m[1].real() == 0 && t.m[2].real() == 0 && t.m[3].imag() == 0;
            // pivoted in-place forms (see sweep_kernel.cu): only when the pivot d0 is comfortably away from zero
            const bool pivot_ok = std::abs(t.m[0].real()) >= 0.25;
            if (rr && pivot_ok)
            {
                d.code = RC_DENSE1_RR;
                const double d0 = t.m[0].real(), d1 = t.m[1].real(), d2 = t.m[2].real(), d3 = t.m[3].real();
                d.m[0] = d0; d.m[1] = d1; d.m[2] = d2 / d0; d.m[3] = (d0 * d3 - d1 * d2) / d0;
                // s*[[1,1],[1,-1]]: det/d0 is -2s exactly (the rounded quotient can be an ulp off, which would hide
                // the Hadamard form from fold_hadamard_scales)
                if (d1 == d0 && d2 == d0 && d3 == -d0) { d.m[2] = 1.0; d.m[3] = -2.0 * d0; }
            }
            else if (ri && pivot_ok)
            {
                d.code = RC_DENSE1_RI;
                const double d0 = t.m[0].real(), d1 = t.m[1].imag(), d2 = t.m[2].imag(), d3 = t.m[3].real();
                d.m[0] = d0; d.m[1] = d1; d.m[2] = d2 / d0; d.m[3] = (d0 * d3 + d1 * d2) / d0;
            }
            else
            {
                d.code = RC_DENSE1;
                for (int i = 0; i < 4; i++) put(d, i, t.m[i]);
            }
        }
        return d;
    }
    cplx m[16];
    if (p0 > p1) memcpy(m, t.m, sizeof(m));
    else { swap_roles(t.m, m); std::swap(p0, p1); }
    d.pos = p0 * (p0 - 1) / 2 + p1; // (1,0) (2,0) (2,1) (3,0) (3,1) (3,2) -> 0..5
    int src[4];
    const int cls = classify(2, m, src);
    if (cls == CLS_DIAG2)
    {
        d.code = RC_DIAG2;
        int skip = 0;
        for (int r = 0; r < 4; r++)
        {
            put(d, r, m[r * 5]);
            if (m[r * 5] == one) skip |= 1 << r;
        }
        d.aux = skip << 8;
        return d;
    }
    if (cls == CLS_MONO2)
    {
        static const int perms[3][4] = {{0, 1, 3, 2}, {0, 3, 2, 1}, {0, 2, 1, 3}};
        for (int w = 0; w < 3; w++)
            if (src[0] == perms[w][0] && src[1] == perms[w][1] && src[2] == perms[w][2] && src[3] == perms[w][3])
            {
                d.code = RC_PERM2;
                bool unit = true;
                for (int r = 0; r < 4; r++)
                {
                    put(d, r, m[r * 4 + src[r]]);
                    if (m[r * 4 + src[r]] != one) unit = false;
                }
                d.aux = w | ((unit ? 1 : 0) << 12);
                return d;
            }
    }
    // M = L U without pivoting: applied in place (no copies of the inputs) when the factors stay small
    if (sweep_dense2_lu())
    {
        cplx a[16], l[16];
        memcpy(a, m, sizeof(a));
        for (auto& x : l) x = 0;
        double big = 0;
        bool ok = true;
        for (int k = 0; k < 4 && ok; k++)
        {
            if (std::abs(a[k * 5]) < 1e-3) { ok = false; break; }
            for (int i = k + 1; i < 4; i++)
            {
                l[i * 4 + k] = a[i * 4 + k] / a[k * 5];
                for (int c = k; c < 4; c++) a[i * 4 + c] -= l[i * 4 + k] * a[k * 4 + c];
                a[i * 4 + k] = 0;
            }
        }
        for (int i = 0; ok && i < 16; i++) big = std::max(big, std::max(std::abs(a[i]), std::abs(l[i])));
        if (ok && big <= 64.0)
        {
            d.code = RC_DENSE2_LU;
            int at = 0;
            for (int i = 0; i < 4; i++)
                for (int j = i; j < 4; j++) put(d, at++, a[i * 4 + j]);
            for (int i = 1; i < 4; i++)
                for (int j = 0; j < i; j++) put(d, at++, l[i * 4 + j]);
            return d;
        }
    }
    d.code = RC_DENSE2;
    for (int i = 0; i < 16; i++) put(d, i, m[i]);
    return d;
}
} // namespace

// number of tile bits inside one TMA box (the rest is enumerated by separate copies) and the box's dimensions: runs of
// consecutive physical bits, dimension 0 = the 128-byte run, <= 8 bits per dimension, <= 5 dimensions, <= tma_box_bits bits
static int tma_box_layout(const Sweep& sw, int start[5], int len[5], int& nd)
{
    nd = 0;
    int nbox = 0;
    for (int d = 0; d < 5; d++) start[d] = len[d] = 0;
    for (int i = 0; i < sw.k && nbox < sweep_tma_box_bits(); i++)
    {
        const int p = sw.in_pos[i];
        const bool extend = nd > 0 && p == start[nd - 1] + len[nd - 1] && len[nd - 1] < (nd == 1 ? 3 : 8);
        if (!extend)
        {
            if (nd == 5) break;
            start[nd] = p;
            len[nd++] = 0;
        }
        len[nd - 1]++;
        nbox++;
    }
    return nbox;
}

namespace
{
struct RoundPlan
{
    std::vector<int> ops;   // indices into sw.ops, execution order
    unsigned touched = 0;   // tile bits its ops need in registers
};

inline bool is_cp(const TileOp& t) { return t.cls == CLS_CPHASE; }

// Splits the sweep's op list into register rounds.  Like the tile scheduler one level up: pick <= R register bits by
// gain (how much of the remaining list becomes executable: per-bit program order, diagonal ops hop over skipped
// diagonal ops), run everything that fits, repeat.  A controlled phase (CLS_CPHASE) runs as soon as ONE of its bits is
// a register bit.  Sweeps containing SRN (a full barrier) keep strict list order.
// `backward`: the same greedy choice made from the END of the list (every rule above is symmetric under reversal): the round
// built first -- the fullest one -- then runs LAST and the leftovers first.
// `first_R`: register-bit budget of the FIRST round (<= R): a sweep whose bits do not fill its last round is planned
// "light first" instead, so that the leftover round runs first and a full one last.
static std::vector<RoundPlan> plan_rounds_dir(const Sweep& sw, int R, bool backward, int first_R = -1)
{
    if (first_R < 1 || first_R > R) first_R = R;
    const int n = (int)sw.ops.size();
    std::vector<char> done(n, 0), diag(n, 0);
    bool has_srn = false;
    auto op_at = [&](int i) -> const TileOp& { return sw.ops[backward ? n - 1 - i : i]; };
    for (int i = 0; i < n; i++)
    {
        const TileOp& t = op_at(i);
        if (t.cls == CLS_SRN1) has_srn = true;
        else if (is_cp(t)) diag[i] = 1;
        else
        {
            const int c = classify(t.nb, t.m, nullptr);
            diag[i] = (c == CLS_DIAG1 || c == CLS_DIAG2);
        }
    }
    auto bits_of = [&](int i) { // the op's bits INSIDE the tile
        unsigned m = 1u << op_at(i).j0;
        if (op_at(i).nb == 2 && op_at(i).j1 >= 0) m |= 1u << op_at(i).j1;
        return m;
    };
    std::vector<RoundPlan> rounds;
    int first = 0, left = n;
    if (has_srn)
    {
        if (backward) return rounds; // (SRN sweeps keep strict list order: forward only)
        while (first < n)
        {
            RoundPlan rp;
            while (first < n && __builtin_popcount(rp.touched | bits_of(first)) <= R)
            {
                rp.touched |= bits_of(first);
                rp.ops.push_back(first++);
            }
            rounds.push_back(rp);
        }
        return rounds;
    }
    // which ops run with the register bit set `rb` fixed
    auto scan = [&](unsigned rb, std::vector<int>* picked) {
        unsigned soft = 0, hard = 0; // bits with a skipped diagonal / non-diagonal op
        int score = 0;
        for (int i = first; i < n && (hard & rb) != rb; i++)
        {
            if (done[i]) continue;
            const unsigned m = bits_of(i);
            const bool fits = is_cp(op_at(i)) ? (m & rb) != 0 : (m & ~rb) == 0;
            const bool ok = fits && !(m & hard) && (diag[i] || !(m & soft));
            if (ok)
            {
                score += 1;
                if (picked) picked->push_back(i);
            }
            else if (diag[i]) soft |= m;
            else hard |= m;
        }
        return score;
    };
    while (left > 0)
    {
        while (first < n && done[first]) first++;
        unsigned rb = 0;
        int cur = 0;
        const int Rcur = rounds.empty() ? first_R : R;
        while (__builtin_popcount(rb) < Rcur)
        {
            const int room = Rcur - __builtin_popcount(rb);
            std::vector<unsigned> cands;
            int looked = 0;
            auto add_cand = [&](unsigned miss) {
                if (!miss || __builtin_popcount(miss) > room) return;
                if (std::find(cands.begin(), cands.end(), miss) == cands.end()) cands.push_back(miss);
            };
            for (int i = first; i < n && looked < 96; i++)
            {
                if (done[i]) continue;
                looked++;
                const unsigned m = bits_of(i);
                if (is_cp(op_at(i)))
                {
                    if (m & rb) continue;
                    for (int p = 0; p < 16; p++)
                        if ((m >> p) & 1u) add_cand(1u << p);
                }
                else add_cand(m & ~rb);
            }
            int best_gain = 0, best = -1;
            double best_rate = 0;
            for (size_t c = 0; c < cands.size(); c++)
            {
                const int gain = scan(rb | cands[c], nullptr) - cur;
                const double rate = (double)gain / __builtin_popcount(cands[c]);
                if (gain > 0 && (rate > best_rate || (rate == best_rate && gain > best_gain)))
                {
                    best_rate = rate; best_gain = gain; best = (int)c;
                }
            }
            if (best < 0) break;
            rb |= cands[best];
            cur += best_gain;
        }
        RoundPlan rp;
        scan(rb, &rp.ops);
        for (int i : rp.ops)
        {
            done[i] = 1;
            rp.touched |= is_cp(op_at(i)) ? (bits_of(i) & rb) : bits_of(i);
            left--;
        }
        if (rp.ops.empty()) break; // cannot happen: the first pending op always fits a fresh round
        rounds.push_back(rp);
    }
    if (backward)
    {
        // back to list indices and execution order
        std::reverse(rounds.begin(), rounds.end());
        for (RoundPlan& rp : rounds)
        {
            for (int& i : rp.ops) i = n - 1 - i;
            std::reverse(rp.ops.begin(), rp.ops.end());
        }
    }
    return rounds;
}

// Forward plan, or (option "heavy_last", off by default) the backward one when it needs no more rounds and ends on a heavier
// round.  The last round of a full-size tile shadows the load of the CTA's next tile (DevDirect): on qft_n15 the sweep whose last
// round is its fullest runs at 0.86 of the HBM peak, the two that end on a one-op leftover round at 0.69 -- but rounds planned
// from the back get their controlled phases BEFORE the butterflies of their bits, where the encoder's deferred-diagonal merging
// (one star per round, RC_QFT2) does not apply, and the extra device ops cost more than the better overlap wins.
std::vector<RoundPlan> plan_rounds(const Sweep& sw, int R)
{
    std::vector<RoundPlan> fwd = plan_rounds_dir(sw, R, false);
    if (g_light_first && fwd.size() >= 2)
    {
        // the last round of a full-size tile shadows the load of the CTA's next tile: when it is a leftover (fewer register
        // bits than the others), plan again with that budget for the FIRST round; same number of rounds or the plan is dropped
        auto cost = [&](const RoundPlan& rp) {
            double w = 0;
            for (int i : rp.ops) w += is_cp(sw.ops[i]) ? 2 : (sw.ops[i].nb == 2 ? 16 : 4); // (rough FP64 cost per element)
            return w;
        };
        const int left = __builtin_popcount(fwd.back().touched);
        bool srn = false;
        for (const TileOp& t : sw.ops) srn = srn || t.cls == CLS_SRN1;
        if (!srn && left >= 1 && left < R)
        {
            std::vector<RoundPlan> lf = plan_rounds_dir(sw, R, false, left);
            size_t n0 = 0, n1 = 0;
            for (const RoundPlan& rp : fwd) n0 += rp.ops.size();
            for (const RoundPlan& rp : lf) n1 += rp.ops.size();
            if (n0 == n1 && lf.size() <= fwd.size() && cost(lf.back()) > cost(fwd.back())) fwd.swap(lf);
        }
    }
    if (!g_heavy_last || fwd.size() < 2) return fwd;
    std::vector<RoundPlan> bwd = plan_rounds_dir(sw, R, true);
    if (bwd.empty() || bwd.size() > fwd.size()) return fwd;
    auto weight = [&](const RoundPlan& rp) {
        double w = 0;
        for (int i : rp.ops) w += is_cp(sw.ops[i]) ? 2 : (sw.ops[i].nb == 2 ? 16 : 4); // (rough FP64 cost per element)
        return w;
    };
    size_t nf = 0, nb = 0;
    for (const RoundPlan& rp : fwd) nf += rp.ops.size();
    for (const RoundPlan& rp : bwd) nb += rp.ops.size();
    if (nf != nb) return fwd; // (cannot happen: both cover every op)
    return (bwd.size() < fwd.size() || weight(bwd.back()) > weight(fwd.back())) ? bwd : fwd;
}

// H gates of a sweep: s*[[1,1],[1,-1]] becomes the payload-free butterfly RC_HAD (2 FP64 instructions per pair and
// component instead of 4) and its scale s moves into the payload of a DENSE op of the same sweep: a real scalar
// commutes with every op (SRN included, it is real-linear), across rounds too.  Only dense payloads carry scales
// (scaling a diagonal table would turn its skipped unit entries into multiplications); when the sweep has none, its
// last H stays a scaled real 2x2 and carries the product.  RC_DENSE1_RR payload = {d0, d1, d2/d0, det/d0}.
void fold_hadamard_scales(std::vector<DevOp>& ops)
{
    auto is_had = [](const DevOp& d) {
        return d.code == RC_DENSE1_RR && d.m[0] != 0.0 && d.m[1] == d.m[0] && d.m[2] == 1.0 && d.m[3] == -2.0 * d.m[0];
    };
    int carrier = -1;
    for (int i = (int)ops.size() - 1; i >= 0 && carrier < 0; i--)
    {
        const int c = ops[i].code;
        if ((c == RC_DENSE1 || c == RC_DENSE2 || c == RC_DENSE2_LU || c == RC_DENSE1_RI || c == RC_DENSE1_RR) && !is_had(ops[i])) carrier = i;
    }
    if (carrier < 0)
        for (int i = (int)ops.size() - 1; i >= 0 && carrier < 0; i--)
            if (is_had(ops[i])) carrier = i;
    if (carrier < 0) return;
    double f = 1.0;
    for (int i = 0; i < (int)ops.size(); i++)
    {
        DevOp& d = ops[i];
        if (i == carrier || !is_had(d)) continue;
        f *= d.m[0];
        d.code = RC_HAD;
        d.aux = 1 << d.pos;
        memset(d.m, 0, sizeof(d.m));
    }
    if (f == 1.0) return;
    DevOp& d = ops[carrier];
    switch (d.code)
    {
    case RC_DENSE1: for (int i = 0; i < 8; i++) d.m[i] *= f; break;
    case RC_DENSE2: for (int i = 0; i < 32; i++) d.m[i] *= f; break;
    case RC_DENSE2_LU: for (int i = 0; i < 20; i++) d.m[i] *= f; break; // (f L U = L (f U))
    default: d.m[0] *= f; d.m[1] *= f; d.m[3] *= f; break; // pivoted forms: d2/d0 is scale free
    }
}

// register bits a device op acts on (conservative for the table diagonals)
unsigned reg_support(const DevOp& d)
{
    static const int hi[6] = {1, 2, 2, 3, 3, 3}, lo[6] = {0, 0, 1, 0, 1, 2};
    switch (d.code)
    {
    case RC_DENSE2: case RC_DENSE2_LU: case RC_PERM2: case RC_CP2: case RC_QFT2: return (1u << hi[d.pos]) | (1u << lo[d.pos]);
    case RC_HAD: case RC_STAR: return (unsigned)d.aux & 15u;
    case RC_DIAGR: case RC_DIAGP: return 15u;
    default: return 1u << d.pos;
    }
}

// Butterflies on different register bits commute with each other and with every op on other bits: an RC_HAD moves back
// over ops that do not touch its bit and joins the previous RC_HAD of its round (one dispatch for up to 4 butterflies).
void merge_butterflies(EncodedSweep& out)
{
    std::vector<DevOp> ops;
    ops.reserve(out.ops.size());
    for (DevRound& rd : out.rounds)
    {
        const size_t begin = ops.size();
        for (int j = 0; j < rd.count; j++)
        {
            const DevOp& d = out.ops[(size_t)rd.first + j];
            bool merged = false;
            if (d.code == RC_HAD)
                for (size_t i = ops.size(); i-- > begin;)
                {
                    if (ops[i].code == RC_HAD) { ops[i].aux |= d.aux; merged = true; break; }
                    if (reg_support(ops[i]) & (unsigned)d.aux) break;
                }
            if (!merged) ops.push_back(d);
        }
        // butterfly(lo) . controlled phase(hi, lo) . butterfly(hi)  ->  one RC_QFT2 (the radix-4 step of a QFT round)
        static const int hi[6] = {1, 2, 2, 3, 3, 3}, lo[6] = {0, 0, 1, 0, 1, 2};
        for (size_t j = begin + 1; j + 1 < ops.size(); j++)
        {
            if (ops[j].code != RC_CP2 || ops[j - 1].code != RC_HAD || ops[j + 1].code != RC_HAD) continue;
            const int bl = 1 << lo[ops[j].pos], bh = 1 << hi[ops[j].pos];
            if (!(ops[j - 1].aux & bl) || !(ops[j + 1].aux & bh) || (ops[j - 1].aux & bh) || (ops[j + 1].aux & bl)) continue;
            ops[j].code = RC_QFT2;
            ops[j - 1].aux &= ~bl;
            ops[j + 1].aux &= ~bh;
            if (!ops[j + 1].aux) ops.erase(ops.begin() + (long)j + 1);
            if (!ops[j - 1].aux) { ops.erase(ops.begin() + (long)j - 1); j--; }
        }
        rd.first = (int32_t)begin;
        rd.count = (int32_t)(ops.size() - begin);
    }
    out.ops.swap(ops);
}

// The diagonal-type ops of a round all commute with each other, and each commutes with every op that touches none
// of its register bits.  They are therefore kept PENDING and only materialise when a non-diagonal op needs one of
// their register bits (or at the end of the round): one device op for many circuit ops.
struct StarPartner
{
    bool tile;  // partner is a tile-local bit (lane / iteration / warp bit), else a physical bit outside the tile
    int bit;
    cplx phi;
};
struct DiagFactor // diagonal over the round's register bits
{
    unsigned support; // register bits it depends on
    cplx e[kRegElems];
};
struct Pending
{
    std::vector<DiagFactor> diag;
    std::vector<StarPartner> star[kRegBits]; // controlled phases whose other bit is NOT a register bit of the round
};
} // namespace

void encode_sweep(const Sweep& sw, EncodedSweep& out)
{
    out.ops.clear();
    out.stream.clear();
    out.rounds.clear();
    out.groups.clear();
    out.stars.clear();
    out.op_mask = 0;
    memset(&out.direct, 0, sizeof(out.direct));
    const int k = sw.k;
    const int mode = sw.swz_mode;
    const int nwb = k >= kWarpBits + kRegBits ? kWarpBits : 0;
    const int R = std::min(kRegBits, k - nwb);
    const std::vector<RoundPlan> plan = plan_rounds(sw, R);
    const size_t nr = plan.size();
    size_t first = 0;
    while (first < nr)
    {
        // ---- group: consecutive rounds that leave kWarpBits tile bits untouched ----
        unsigned used = 0;
        size_t end = first;
        // (kSwzTma: only tile bits 0..5 move the bank group, and lane bits 0..2 need three of them with different
        // residues mod 3 -- a group ends early rather than let its warp bits eat into them)
        const unsigned high_mask = k > 6 ? ((1u << k) - 1u) & ~63u : 0u;
        while (end < nr)
        {
            const unsigned u = used | plan[end].touched;
            if (k - __builtin_popcount(u) < nwb) break;
            if (mode == kSwzTma && end > first && __builtin_popcount(~u & high_mask) < nwb) break;
            used = u;
            end++;
        }
        DevGroup g;
        memset(&g, 0, sizeof(g));
        g.first = (int32_t)out.rounds.size();
        g.n_warps = 1 << nwb;
        std::vector<int> wpos; // the highest untouched bits carry the warp index
        // (last group of a TMA tile: the top tile bit is left to the iteration index when other high bits are free -- the
        // two halves of the tile buffer are then reloaded one after the other during the last round, see DevDirect)
        const bool spare_top = mode == kSwzTma && end == nr && k == kMaxTileBits && g_direct_store &&
                               __builtin_popcount(~used & high_mask & ~(1u << (k - 1))) >= nwb;
        for (int p = k - 1; p >= 0 && (int)wpos.size() < nwb; p--)
            if (!((used >> p) & 1u) && !(spare_top && p == k - 1)) wpos.push_back(p);
        std::sort(wpos.begin(), wpos.end());
        unsigned wmask = 0;
        for (int p : wpos) wmask |= 1u << p;
        for (int w = 0; w < (1 << kWarpBits); w++) g.wtab[w] = (uint16_t)(16u * swz_host(deposit((unsigned)w, wpos), mode));

        for (size_t ri = first; ri < end; ri++)
        {
            const RoundPlan& rp = plan[ri];
            std::vector<int> rb;
            for (int p = 0; p < k; p++)
                if ((rp.touched >> p) & 1u) rb.push_back(p);
            // pad the register bits with free tile bits (lowest first; kSwzTma: highest first, the low ones are lane bits)
            // (last round of a TMA tile: the top tile bit is not used as padding, it becomes the iteration bit -- DevDirect)
            const bool keep_top = mode == kSwzTma && ri + 1 == nr && k == kMaxTileBits && g_direct_store && !((wmask >> (k - 1)) & 1u) &&
                                  k - nwb - (int)rb.size() > R - (int)rb.size();
            for (int q = 0; q < k && (int)rb.size() < R; q++)
            {
                const int p = mode == kSwzTma ? k - 1 - q : q;
                if (keep_top && p == k - 1) continue;
                if (!((wmask >> p) & 1u) && std::find(rb.begin(), rb.end(), p) == rb.end()) rb.push_back(p);
            }
            std::sort(rb.begin(), rb.end());
            unsigned rmask = 0;
            for (int p : rb) rmask |= 1u << p;

            DevRound rd;
            memset(&rd, 0, sizeof(rd));
            rd.first = (int32_t)out.ops.size(); // op INDEX for now; rewritten to the stream offset below
            const int rd_first = rd.first;
            for (int c = 0; c < kRegElems; c++) rd.roff[c] = (uint16_t)(16u * swz_host(deposit((unsigned)c & ((1u << R) - 1u), rb), mode));
            std::vector<int> freep;
            for (int p = 0; p < k; p++)
                if (!(((wmask | rmask) >> p) & 1u)) freep.push_back(p);
            const int nl = std::min(5, (int)freep.size());
            // lane bits 0..2 vary inside one LDS.128 phase: give them positions from three different swizzle
            // classes (p mod 3, see swz_host) whenever one is free, so that the phase is bank-conflict free
            std::vector<int> lanep;
            std::vector<char> taken(k, 0);
            // (kSwzTma: only tile bits 0..5 move the bank group)
            for (int cls = 0; cls < 3 && (int)lanep.size() < nl; cls++)
                for (int p : freep)
                    if (p % 3 == cls && !taken[p] && (mode != kSwzTma || p < 6))
                    {
                        lanep.push_back(p);
                        taken[p] = 1;
                        break;
                    }
            for (int p : freep)
                if ((int)lanep.size() < nl && !taken[p]) { lanep.push_back(p); taken[p] = 1; }
            std::vector<int> iterp;
            for (int p : freep)
                if (!taken[p]) iterp.push_back(p);
            rd.n_iter = 1 << (int)iterp.size();
            rd.n_active = 1 << nl;
            for (int l = 0; l < 32; l++) rd.lane_tab[l] = (uint16_t)(16u * swz_host(deposit((unsigned)l & ((1u << nl) - 1u), lanep), mode));
            for (int it = 0; it < 8; it++)
                rd.iter_tab[it] = (uint16_t)(16u * swz_host(deposit((unsigned)it & ((unsigned)rd.n_iter - 1u), iterp), mode));
            out.rounds.push_back(rd);
            if (ri + 1 == nr)
            {
                // direct store of the sweep's last round: full-size in-place TMA tile whose lane bits 0..2 are the 128-byte run
                DevDirect& dd = out.direct;
                memset(&dd, 0, sizeof(dd));
                const bool run = nl == 5 && sw.out_pos[lanep[0]] < 3 && sw.out_pos[lanep[1]] < 3 && sw.out_pos[lanep[2]] < 3;
                if (g_direct_store && mode == kSwzTma && k == kMaxTileBits && nwb == kWarpBits && R == kRegBits && sw.in_pos == sw.out_pos &&
                    !sw.out_of_place && run && (int)iterp.size() <= 3)
                {
                    dd.enabled = 1;
                    for (int i = 0; i < R; i++) dd.reg_pos[i] = (unsigned char)sw.out_pos[rb[i]];
                    for (int i = 0; i < nl; i++) dd.lane_pos[i] = (unsigned char)sw.out_pos[lanep[i]];
                    for (int i = 0; i < nwb; i++) dd.warp_pos[i] = (unsigned char)sw.out_pos[wpos[i]];
                    for (size_t i = 0; i < iterp.size(); i++) dd.iter_pos[i] = (unsigned char)sw.out_pos[iterp[i]];
                    for (int c = 0; c < kRegElems; c++)
                        for (int i = 0; i < R; i++)
                            if ((c >> i) & 1) dd.reg_off[c] |= 16ull << dd.reg_pos[i];
                    for (int it = 0; it < 8; it++)
                        for (size_t i = 0; i < iterp.size(); i++)
                            if ((it >> i) & 1) dd.iter_off[it] |= 1ull << dd.iter_pos[i];
                    dd.half_enum = -1;
                    int bs[5], bl[5], bnd;
                    const int nbox = tma_box_layout(sw, bs, bl, bnd); // tile bits [nbox, k) are enumerated by the TMA copies
                    if (iterp.size() == 1 && iterp[0] >= nbox) dd.half_enum = iterp[0] - nbox;
                }
            }

            auto reg_pos = [&](int j) { // position of tile bit j among the round's register bits, -1 if none
                const auto it = std::find(rb.begin(), rb.end(), j);
                return it == rb.end() ? -1 : (int)(it - rb.begin());
            };
            Pending pend;
            // materialise the pending diagonal factors / stars that involve a register bit of `need`
            auto flush = [&](unsigned need) {
                cplx e[kRegElems];
                for (auto& x : e) x = cplx(1.0, 0.0);
                bool have = false;
                for (size_t i = 0; i < pend.diag.size();)
                {
                    if (!(pend.diag[i].support & need)) { i++; continue; }
                    for (int c = 0; c < kRegElems; c++) e[c] *= pend.diag[i].e[c];
                    have = true;
                    pend.diag.erase(pend.diag.begin() + (long)i);
                }
                if (have)
                {
                    int skip = 0;
                    unsigned common = (1u << kRegBits) - 1u; // register bits set in every non-unit entry
                    for (int c = 0; c < kRegElems; c++)
                    {
                        // entries within 1e-15 of 1 (e.g. u1(a)*u1(-a) inside a fused controlled phase) are exactly 1
                        if (std::abs(e[c].real() - 1.0) < 1e-15 && std::abs(e[c].imag()) < 1e-15) e[c] = cplx(1.0, 0.0);
                        if (e[c] == cplx(1.0, 0.0)) skip |= 1 << c;
                        else common &= (unsigned)c;
                    }
                    DevOp nd;
                    memset(&nd, 0, sizeof(nd));
                    // one controlled phase between two register bits: exactly the 4 entries with both bits set, equal
                    int cp_hi = -1, cp_lo = -1;
                    if (__builtin_popcount(common) == 2)
                    {
                        cp_lo = __builtin_ctz(common);
                        cp_hi = 31 - __builtin_clz(common);
                        for (int c = 0; c < kRegElems; c++)
                        {
                            const bool in = ((unsigned)c & common) == common;
                            if (in != !((skip >> c) & 1) || (in && e[c] != e[common])) cp_hi = -1;
                        }
                    }
                    if (skip == (1 << kRegElems) - 1) {}
                    else if (cp_hi >= 0)
                    {
                        nd.code = RC_CP2;
                        nd.pos = cp_hi * (cp_hi - 1) / 2 + cp_lo;
                        put(nd, 0, e[common]);
                        out.ops.push_back(nd);
                    }
                    else if (common)
                    {
                        // every non-unit entry has register bit P set (phases controlled by P): an 8-entry table
                        // over the other register bits, so that no instruction is spent on the unit half
                        int P = 0;
                        while (!((common >> P) & 1u)) P++;
                        nd.code = RC_DIAGP;
                        nd.pos = P;
                        int skip8 = 0;
                        for (int j = 0; j < kRegElems / 2; j++)
                        {
                            const int c = ((j >> P) << (P + 1)) | (1 << P) | (j & ((1 << P) - 1));
                            put(nd, j, e[c]);
                            if ((skip >> c) & 1) skip8 |= 1 << j;
                        }
                        nd.aux = skip8;
                        out.ops.push_back(nd);
                    }
                    else
                    {
                        nd.code = RC_DIAGR;
                        for (int c = 0; c < kRegElems; c++) put(nd, c, e[c]);
                        nd.aux = skip;
                        out.ops.push_back(nd);
                    }
                }
                bool any_star = false;
                for (int p = 0; p < kRegBits; p++)
                    if (((need >> p) & 1u) && !pend.star[p].empty()) any_star = true;
                if (any_star)
                {
                    DevOp sd;
                    memset(&sd, 0, sizeof(sd));
                    sd.code = RC_STAR;
                    sd.vid = (int32_t)out.stars.size(); // first slot (host-side meaning of vid for RC_STAR)
                    const int nib = (int)iterp.size();
                    for (int p = 0; p < kRegBits; p++)
                    {
                        if (!((need >> p) & 1u) || pend.star[p].empty()) continue;
                        sd.aux |= 1 << p;
                        DevStar st;
                        memset(&st, 0, sizeof(st));
                        for (int iw = 0; iw < kStarW; iw++)
                        {
                            const unsigned itv = (unsigned)iw & ((1u << nib) - 1u), wv = (unsigned)iw >> nib;
                            const unsigned idx = deposit(itv, iterp) | deposit(wv, wpos);
                            cplx acc(1.0, 0.0);
                            for (const StarPartner& sp : pend.star[p])
                                if (sp.tile && ((idx >> sp.bit) & 1u)) acc *= sp.phi;
                            st.w[2 * iw] = acc.real();
                            st.w[2 * iw + 1] = acc.imag();
                        }
                        for (int l = 0; l < 12; l++) // la[0..7]: lane bits 0..2, lb[0..3]: lane bits 3..4
                        {
                            const unsigned lane = l < 8 ? (unsigned)l : (unsigned)(l - 8) << 3;
                            const unsigned idx = deposit(lane & ((1u << nl) - 1u), lanep);
                            cplx acc(1.0, 0.0);
                            for (const StarPartner& sp : pend.star[p])
                                if (sp.tile && ((idx >> sp.bit) & 1u)) acc *= sp.phi;
                            double* dst = l < 8 ? st.la + 2 * l : st.lb + 2 * (l - 8);
                            dst[0] = acc.real();
                            dst[1] = acc.imag();
                        }
                        // partners outside the tile: one entry per physical bit
                        for (const StarPartner& sp : pend.star[p])
                            if (!sp.tile)
                            {
                                int j = 0;
                                while (j < st.n_out && st.bit[j] != sp.bit) j++;
                                if (j == st.n_out)
                                {
                                    if (st.n_out >= kMaxStarOut) throw std::logic_error("too many outside partners in a star");
                                    st.bit[j] = sp.bit;
                                    st.phi[2 * j] = 1.0;
                                    st.phi[2 * j + 1] = 0.0;
                                    st.n_out++;
                                }
                                const cplx acc = cplx(st.phi[2 * j], st.phi[2 * j + 1]) * sp.phi;
                                st.phi[2 * j] = acc.real();
                                st.phi[2 * j + 1] = acc.imag();
                            }
                        for (int j = st.n_out; j < kMaxStarOut; j++) // padding: bit 63 of an index is never set
                        {
                            st.bit[j] = 63;
                            st.phi[2 * j] = 1.0;
                        }
                        out.stars.push_back(st);
                        pend.star[p].clear();
                    }
                    out.ops.push_back(sd);
                }
            };
            const unsigned all_regs = (1u << kRegBits) - 1u;

            for (int o : rp.ops)
            {
                const TileOp& t = sw.ops[o];
                if (is_cp(t))
                {
                    const cplx phi = t.m[15];
                    const int p0 = reg_pos(t.j0), p1 = t.j1 >= 0 ? reg_pos(t.j1) : -1;
                    if (p0 >= 0 && p1 >= 0)
                    {
                        DiagFactor f;
                        f.support = (1u << p0) | (1u << p1);
                        for (int c = 0; c < kRegElems; c++) f.e[c] = (((c >> p0) & 1) && ((c >> p1) & 1)) ? phi : cplx(1.0, 0.0);
                        pend.diag.push_back(f);
                    }
                    else if (p0 >= 0 || p1 >= 0)
                    {
                        StarPartner sp;
                        sp.phi = phi;
                        if (p0 >= 0) { sp.tile = t.j1 >= 0; sp.bit = t.j1 >= 0 ? t.j1 : t.p1; }
                        else { sp.tile = true; sp.bit = t.j0; }
                        pend.star[p0 >= 0 ? p0 : p1].push_back(sp);
                    }
                    else
                        throw std::logic_error("controlled phase without a register bit in its round");
                    continue;
                }
                const int p0 = reg_pos(t.j0);
                const int p1 = t.nb == 2 ? reg_pos(t.j1) : 0;
                DevOp d = make_reg_op(t, p0, p1);
                if (d.code == RC_DIAG1 || d.code == RC_DIAG2)
                {
                    // every diagonal op becomes a 16-entry diagonal over the round's register bits; the pending
                    // ones are multiplied together on the host: one device op, no position dispatch
                    static const int hi[6] = {1, 2, 2, 3, 3, 3}, lo[6] = {0, 0, 1, 0, 1, 2};
                    DiagFactor f;
                    f.support = d.code == RC_DIAG1 ? (1u << d.pos) : ((1u << hi[d.pos]) | (1u << lo[d.pos]));
                    for (int c = 0; c < kRegElems; c++)
                    {
                        const int idx = d.code == RC_DIAG1 ? ((c >> d.pos) & 1) : 2 * ((c >> hi[d.pos]) & 1) + ((c >> lo[d.pos]) & 1);
                        f.e[c] = cplx(d.m[2 * idx], d.m[2 * idx + 1]);
                    }
                    pend.diag.push_back(f);
                    continue;
                }
                // SRN is real-linear only: it does not commute with complex factors on OTHER bits, so everything
                // pending is materialised before it
                unsigned need = all_regs;
                if (d.code != RC_SRN1) need = t.nb == 2 ? ((1u << p0) | (1u << p1)) : (1u << p0);
                flush(need);
                out.ops.push_back(d);
            }
            flush(all_regs);
            // (first / count are op INDICES until the stream is laid out below; long rounds are split)
            {
                int left = (int32_t)out.ops.size() - rd_first, at = rd_first;
                out.rounds.back().count = std::min(left, kMaxOpsPerRound);
                while ((left -= kMaxOpsPerRound) > 0)
                {
                    DevRound more = out.rounds.back();
                    at += kMaxOpsPerRound;
                    more.first = at;
                    more.count = std::min(left, kMaxOpsPerRound);
                    out.rounds.push_back(more);
                }
            }
        }
        g.count = (int32_t)out.rounds.size() - g.first;
        out.groups.push_back(g);
        first = end;
    }
    if ((int)out.stars.size() > kMaxStarsPerSweep) throw std::logic_error("too many controlled-phase stars in one sweep");
    fold_hadamard_scales(out.ops);
    merge_butterflies(out);

    // FP64 instructions per lane and iteration (kRegElems elements) of every device op: the sweep's algorithmic FP64 work
    out.fp64_per_lane = 0;
    for (const DevOp& d : out.ops)
    {
        const unsigned long long E = kRegElems;
        unsigned long long c = 0;
        switch (d.code)
        {
        case RC_DENSE1: c = 8 * E; break;
        case RC_DENSE1_RR: case RC_DENSE1_RI: c = 4 * E; break;
        case RC_HAD: c = 2 * E * __builtin_popcount((unsigned)d.aux & 15u); break;
        case RC_MONO1: c = ((d.aux >> 12) & 1) ? 0 : 4 * E; break;
        case RC_SRN1: c = 3 * E; break;
        case RC_DENSE2: case RC_DENSE2_LU: c = 16 * E; break;
        case RC_PERM2: c = ((d.aux >> 12) & 1) ? 0 : 4 * E; break;
        case RC_DIAGR: c = 4ull * (E - __builtin_popcount((unsigned)d.aux & ((1u << kRegElems) - 1u))); break;
        case RC_DIAGP: c = 4ull * (E / 2 - __builtin_popcount((unsigned)d.aux & ((1u << (kRegElems / 2)) - 1u))); break;
        case RC_CP2: c = E; break;
        case RC_QFT2: c = 5 * E; break;
        case RC_STAR: c = (4 + 2 * E) * __builtin_popcount((unsigned)d.aux & 15u); break;
        default: break;
        }
        out.fp64_per_lane += c;
    }

    // ---- the device op stream: 16-byte header + the used part of the payload per op, a zero header at the end ----
    std::vector<int> offset16(out.ops.size() + 1, 0);
    for (size_t i = 0; i < out.ops.size(); i++)
    {
        DevOp& d = out.ops[i];
        const int star0 = d.code == RC_STAR ? d.vid : 0;
        DevOpHdr h;
        h.vid = dev_vid(d.code, d.pos, d.aux);
        h.aux = d.aux;
        const int payload = dev_op_payload_bytes(d.code);
        h.size16 = 1 + payload / 16;
        h.star[0] = star0;
        offset16[i] = (int)(out.stream.size() / 16);
        const unsigned char* hb = reinterpret_cast<const unsigned char*>(&h);
        out.stream.insert(out.stream.end(), hb, hb + sizeof(h));
        const unsigned char* pb = reinterpret_cast<const unsigned char*>(d.m);
        out.stream.insert(out.stream.end(), pb, pb + payload);
        out.op_mask |= 1u << d.code;
    }
    offset16[out.ops.size()] = (int)(out.stream.size() / 16);
    out.stream.insert(out.stream.end(), 16, (unsigned char)0);
    for (DevRound& rd : out.rounds)
    {
        memset(rd.vids, 0, sizeof(rd.vids));
        rd.star0 = -1;
        for (int j = 0; j < rd.count; j++)
        {
            const DevOp& d = out.ops[(size_t)rd.first + j];
            rd.vids[j] = (uint8_t)dev_vid(d.code, d.pos, d.aux);
            if (d.code == RC_STAR && rd.star0 < 0) rd.star0 = d.vid; // host-side vid of RC_STAR = its first slot
        }
        if (rd.star0 < 0) rd.star0 = 0;
        rd.first = offset16[rd.first];
    }
}

void fill_sweep_tables(const Sweep& sw, int M, SweepArgs& a)
{
    const int k = sw.k;
    a.k = k;
    a.n_comp = M - k;
    a.n_tiles = 1ull << a.n_comp;
    // load: loop bit i <-> tile-local bit i <-> physical in_pos[i] (ascending by construction)
    // store: enumerate in ascending OUTPUT position, so stores stay coalesced when the sweep permutes bits
    std::vector<int> ord(k);
    for (int i = 0; i < k; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int x, int y) { return sw.out_pos[x] < sw.out_pos[y]; });
    for (int i = 0; i < k && i < kThreadBits; i++)
    {
        a.gin[i] = (unsigned char)sw.in_pos[i];
        a.gout[i] = (unsigned char)sw.out_pos[ord[i]];
        a.sout[i] = (unsigned char)ord[i];
    }
    for (int it = 0; it < kMaxIter; it++)
    {
        unsigned long long hi = 0, ho = 0;
        unsigned hs = 0;
        for (int i = kThreadBits; i < k; i++)
        {
            const unsigned long long bit = (it >> (i - kThreadBits)) & 1;
            hi |= bit << sw.in_pos[i];
            ho |= bit << sw.out_pos[ord[i]];
            hs |= (unsigned)bit << ord[i];
        }
        a.hin[it] = hi << 4; // BYTE offsets (16 B per element)
        a.hout[it] = ho << 4;
        a.hs[it] = (unsigned short)swz_host(hs, sw.swz_mode);
    }
    a.swz_mode = sw.swz_mode;
    a.tma_load = a.tma_store = 0;
    if (sw.swz_mode == kSwzTma && k == kMaxTileBits && sw.in_pos[0] == 0 && sw.in_pos[1] == 1 && sw.in_pos[2] == 2)
    {
        // the box: the lowest tile bits, as runs of consecutive physical bits (dimension 0 = the 128-byte run, <= 8 bits
        // per dimension, <= 5 dimensions, <= tma_box_bits bits); every other tile bit is enumerated by separate copies
        TmaGeom& t = a.tma;
        int nd = 0, start[5], len[5];
        const int nbox = tma_box_layout(sw, start, len, nd);
        for (int d = 0; d < nd; d++) t.start[d] = (unsigned char)start[d];
        for (int d = 0; d < 5; d++)
        {
            if (d >= nd) t.start[d] = (unsigned char)M; // unit dimensions pad the rank to 5
            t.box_log2[d] = (unsigned char)(d < nd ? len[d] : 0);
        }
        for (int d = 0; d < 5; d++) t.span[d] = (unsigned char)((d + 1 < 5 ? t.start[d + 1] : M) - t.start[d]);
        t.n_enum = (unsigned char)(k - nbox);
        t.n_copies = 1 << t.n_enum;
        t.box_bytes = 16 << nbox;
        for (int j = 0; j < t.n_copies; j++)
        {
            unsigned long long off = 0;
            for (int b = 0; b < t.n_enum; b++) off |= (unsigned long long)((j >> b) & 1) << sw.in_pos[nbox + b];
            t.enum_off[j] = off;
        }
        a.tma_load = 1;
        a.tma_store = (sw.in_pos == sw.out_pos && !sw.out_of_place) ? 1 : 0;
    }
    std::vector<char> used_in(M, 0), used_out(M, 0);
    for (int i = 0; i < k; i++) { used_in[sw.in_pos[i]] = 1; used_out[sw.out_pos[i]] = 1; }
    // the positions the tile id enumerates, lowest id bit first: ascending -- except that a remap pack sweep takes the
    // rank-selecting top local bits first (Sweep::spread_top).  Any order is a valid enumeration as long as loading and storing
    // use the same one (a sweep only permutes bits INSIDE its tile: the unused positions are the same set on both sides)
    int ci = 0, co = 0;
    const int top = sw.spread_top > 0 && sw.spread_top < M ? M - sw.spread_top : M;
    for (int pass = 0; pass < 2; pass++)
        for (int p = pass == 0 ? top : 0; p < (pass == 0 ? M : top); p++)
        {
            if (!used_in[p]) a.cin[ci++] = (unsigned char)p;
            if (!used_out[p]) a.cout[co++] = (unsigned char)p;
        }
}

void fill_base_tables(SweepArgs& a)
{
    for (int t = 0; t < 3; t++)
        for (int v = 0; v < 128; v++)
        {
            unsigned long long bi = 0, bo = 0;
            for (int b = 0; b < 7; b++)
            {
                const int i = 7 * t + b;
                if (i >= a.n_comp || !((v >> b) & 1)) continue;
                bi |= 1ull << a.cin[i];
                bo |= 1ull << a.cout[i];
            }
            a.base_in[t][v] = bi;
            a.base_out[t][v] = bo;
        }
}

std::string encoded_to_json(const EncodedSweep& e, const SweepArgs& a)
{
    std::ostringstream o;
    auto arr = [&](const char* name, auto* v, int n, unsigned div = 1) {
        o << "\"" << name << "\":[";
        for (int i = 0; i < n; i++) o << (i ? "," : "") << (unsigned long long)v[i] / div;
        o << "]";
    };
    o << "{\"k\":" << a.k << ",\"n_comp\":" << a.n_comp << ",\"geom\":{\"thread_bits\":" << kThreadBits << ",\"reg_bits\":" << kRegBits
      << ",\"warp_bits\":" << kWarpBits << ",\"star_w\":" << kStarW << "},\"swz\":" << a.swz_mode << ",\"tma_load\":" << a.tma_load
      << ",\"tma_store\":" << a.tma_store << ",";
    if (a.tma_load)
    {
        o << "\"tma\":{\"n_copies\":" << a.tma.n_copies << ",\"box_bytes\":" << a.tma.box_bytes << ",";
        arr("start", a.tma.start, 5); o << ",";
        arr("span", a.tma.span, 5); o << ",";
        arr("box_log2", a.tma.box_log2, 5); o << ",";
        arr("enum_off", a.tma.enum_off, a.tma.n_copies);
        o << "},";
    }
    unsigned long long hin_e[kMaxIter], hout_e[kMaxIter]; // element offsets for the emulator
    for (int it = 0; it < kMaxIter; it++) { hin_e[it] = a.hin[it] >> 4; hout_e[it] = a.hout[it] >> 4; }
    arr("hin", hin_e, kMaxIter); o << ",";
    arr("hout", hout_e, kMaxIter); o << ",";
    arr("hs", a.hs, kMaxIter); o << ",";
    arr("gin", a.gin, 12); o << ",";
    arr("gout", a.gout, 12); o << ",";
    arr("sout", a.sout, 12); o << ",";
    arr("cin", a.cin, a.n_comp); o << ",";
    arr("cout", a.cout, a.n_comp);
    if (e.direct.enabled)
    {
        o << ",\"direct\":{\"half_enum\":" << e.direct.half_enum << ",";
        arr("reg_pos", e.direct.reg_pos, kRegBits); o << ",";
        arr("lane_pos", e.direct.lane_pos, 5); o << ",";
        arr("warp_pos", e.direct.warp_pos, kWarpBits); o << ",";
        arr("iter_pos", e.direct.iter_pos, 1);
        o << "}";
    }
    o << ",\"groups\":[";
    for (size_t g = 0; g < e.groups.size(); g++)
    {
        const DevGroup& G = e.groups[g];
        o << (g ? "," : "") << "{\"first\":" << G.first << ",\"count\":" << G.count << ",\"n_warps\":" << G.n_warps << ",";
        arr("wtab", G.wtab, 16, 16); // (byte offsets on the device, element indices in the JSON)
        o << "}";
    }
    o << "],\"rounds\":[";
    for (size_t r = 0; r < e.rounds.size(); r++)
    {
        const DevRound& D = e.rounds[r];
        o << (r ? "," : "") << "{\"first\":" << D.first << ",\"count\":" << D.count << ",\"n_iter\":" << D.n_iter
          << ",\"n_active\":" << D.n_active << ",";
        arr("lane_tab", D.lane_tab, 32, 16); o << ",";
        arr("iter_tab", D.iter_tab, 8, 16); o << ",";
        arr("roff", D.roff, kRegElems, 16);
        o << "}";
    }
    o << "],\"ops\":[";
    char buf[40];
    auto dbl = [&](const double* v, int n) {
        for (int j = 0; j < n; j++)
        {
            snprintf(buf, sizeof(buf), "%.17g", v[j]);
            o << (j ? "," : "") << buf;
        }
    };
    int off16 = 0;
    for (size_t i = 0; i < e.ops.size(); i++)
    {
        const DevOp& d = e.ops[i];
        // "off": where the op sits in the device stream (DevRound::first refers to it); RC_STAR: "star" = first slot
        o << (i ? "," : "") << "{\"code\":" << d.code << ",\"aux\":" << d.aux << ",\"pos\":" << d.pos << ",\"off\":" << off16
          << ",\"star\":" << (d.code == RC_STAR ? d.vid : -1) << ",\"m\":[";
        dbl(d.m, 32);
        o << "]}";
        off16 += 1 + dev_op_payload_bytes(d.code) / 16;
    }
    o << "],\"stars\":[";
    for (size_t i = 0; i < e.stars.size(); i++)
    {
        const DevStar& st = e.stars[i];
        o << (i ? "," : "") << "{\"w\":[";
        dbl(st.w, 2 * kStarW);
        o << "],\"la\":[";
        dbl(st.la, 16);
        o << "],\"lb\":[";
        dbl(st.lb, 8);
        o << "],\"bit\":[";
        for (int j = 0; j < st.n_out; j++) o << (j ? "," : "") << st.bit[j];
        o << "],\"phi\":[";
        dbl(st.phi, 2 * st.n_out);
        o << "]}";
    }
    o << "]}";
    return o.str();
}
} // namespace dmb
