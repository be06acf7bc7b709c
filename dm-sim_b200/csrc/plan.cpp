// dm-sim_b200/csrc/plan.cpp -- expansion, fusion and tile scheduling (host only; see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace dmb
{
static const double kPI = 3.14159265358979323846;  // reference src/config.hpp:55
static const double kS2I = 0.70710678118654752440; // reference src/config.hpp:57

static const char* kOpNames[] = {"U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S",
                                 "SDG", "T", "TDG", "RX", "RY", "RZ", "CZ", "CY", "SWAP", "CH",
                                 "CCX", "CSWAP", "CRX", "CRY", "CRZ", "CU1", "CU3", "RXX", "RZZ", "RCCX",
                                 "RC3X", "C3X", "C3SQRTX", "C4X", "R", "SRN", "W", "RYY"};

const char* op_name(int op)
{
    if (op >= 0 && op < DMB_OP_COUNT) return kOpNames[op];
    if (op == DMB_OP_C1) return "C1";
    if (op == DMB_OP_C2) return "C2";
    return "?";
}

// ------------------------------------------------------------------------------------------------
// Expansion: every Gate becomes the primitive sequence the reference executes for it.
// Matrices are built from the same real expressions as the reference bodies so that the entries
// agree to the last bit before fusion (reference lines cited per primitive).
// ------------------------------------------------------------------------------------------------
namespace
{
struct Expander
{
    int n;
    std::vector<Block>& out;

    void one(int q, cplx a, cplx b, cplx c, cplx d)
    {
        Block k;
        k.nq = 1;
        k.q[0] = q;
        k.m[0] = a; k.m[1] = b; k.m[2] = c; k.m[3] = d;
        out.push_back(k);
    }
    // CX_GATE :1132-1168: swap (ctrl=1,tgt=0) <-> (ctrl=1,tgt=1); index = 2*bit(ctrl)+bit(tgt)
    void cx(int c, int t)
    {
        Block k;
        k.nq = 2;
        k.q[0] = c; k.q[1] = t;
        for (auto& e : k.m) e = 0;
        k.m[0] = 1; k.m[5] = 1; k.m[11] = 1; k.m[14] = 1;
        out.push_back(k);
    }
    void x(int q) { one(q, 0, 1, 1, 0); }                                   // :1175-1187
    void y(int q) { one(q, 0, cplx(0, -1), cplx(0, 1), 0); }                // :1196-1209
    void z(int q) { one(q, 1, 0, 0, -1); }                                  // :1216-1225
    void h(int q) { one(q, kS2I, kS2I, kS2I, -kS2I); }                      // :1232-1245
    void s(int q) { one(q, 1, 0, 0, cplx(0, 1)); }                          // :1299-1307
    void sdg(int q) { one(q, 1, 0, 0, cplx(0, -1)); }                       // :1314-1322
    void t(int q) { one(q, 1, 0, 0, cplx(kS2I, kS2I)); }                    // :1329-1337
    void tdg(int q) { one(q, 1, 0, 0, cplx(kS2I, -kS2I)); }                 // :1344-1352
    void r(double p, int q) { one(q, 1, 0, 0, cplx(0, p)); }                // R_GATE :1283-1292 (v1 *= i*p)
    void w(int q) { one(q, kS2I, cplx(0, -kS2I), cplx(0, -kS2I), kS2I); }   // :1787-1799
    void u1(double l, int q) { one(q, 1, 0, 0, cplx(cos(l), sin(l))); }     // :1381-1397
    void u2(double p, double l, int q)                                      // :1404-1418
    {
        one(q, kS2I, cplx(-kS2I * cos(l), -kS2I * sin(l)), cplx(kS2I * cos(p), kS2I * sin(p)),
            cplx(kS2I * cos(p + l), kS2I * sin(p + l)));
    }
    void u3(double th, double p, double l, int q)                           // :1425-1440
    {
        one(q, cos(th / 2.), cplx(-cos(l) * sin(th / 2.), -sin(l) * sin(th / 2.)),
            cplx(cos(p) * sin(th / 2.), sin(p) * sin(th / 2.)),
            cplx(cos(p + l) * cos(th / 2.), sin(p + l) * cos(th / 2.)));
    }
    void rx(double th, int q)                                               // :1444-1459
    {
        double c = cos(th / 2.0), ms = -sin(th / 2.0);
        one(q, c, cplx(0, ms), cplx(0, ms), c);
    }
    void ry(double th, int q)                                               // :1463-1481
    {
        double c = cos(th / 2.0), s_ = sin(th / 2.0);
        one(q, c, -s_, s_, c);
    }
    void rz(double p, int q) { u1(p, q); }                                  // :1485-1489 (== U1)
    void srn(int q)                                                         // :1253-1266
    {
        Block k;
        k.nq = 1; k.q[0] = q; k.srn = true;
        out.push_back(k);
    }
    // ---- composites :1493-1780, :1803-1813 ----
    void cz(int a, int b) { h(b); cx(a, b); h(b); }
    void cy(int a, int b) { sdg(b); cx(a, b); s(b); }
    void swap(int a, int b) { cx(a, b); cx(b, a); cx(a, b); }
    void ch(int a, int b)
    {
        h(b); sdg(b); cx(a, b); h(b); t(b); cx(a, b); t(b); h(b); s(b); x(b); s(a);
    }
    void crz(double l, int a, int b) { u1(l / 2, b); cx(a, b); u1(-l / 2, b); cx(a, b); }
    void cu1(double l, int a, int b) { u1(l / 2, a); cx(a, b); u1(-l / 2, b); cx(a, b); u1(l / 2, b); }
    void cu3(double th, double p, double l, int c, int t_)
    {
        double t1 = (l - p) / 2, t2 = th / 2, t3 = -(p + l) / 2;
        u1(-t3, c); u1(t1, t_); cx(c, t_); u3(-t2, 0, t3, t_); cx(c, t_); u3(t2, p, 0, t_);
    }
    void ccx(int a, int b, int c)
    {
        h(c); cx(b, c); tdg(c); cx(a, c); t(c); cx(b, c); tdg(c); cx(a, c);
        t(b); t(c); h(c); cx(a, b); t(a); tdg(b); cx(a, b);
    }
    void cswap(int a, int b, int c) { cx(c, b); ccx(a, b, c); cx(c, b); }
    void crx(double l, int a, int b)
    {
        u1(kPI / 2, b); cx(a, b); u3(-l / 2, 0, 0, b); cx(a, b); u3(l / 2, -kPI / 2, 0, b);
    }
    void cry(double l, int a, int b) { u3(l / 2, 0, 0, b); cx(a, b); u3(-l / 2, 0, 0, b); cx(a, b); }
    void rxx(double th, int a, int b)
    {
        u3(kPI / 2, th, 0, a); h(b); cx(a, b); u1(-th, b); cx(a, b); h(b); u2(-kPI, kPI - th, a);
    }
    void rzz(double th, int a, int b) { cx(a, b); u1(th, b); cx(a, b); }
    void rccx(int a, int b, int c)
    {
        u2(0, kPI, c); u1(kPI / 4, c); cx(b, c); u1(-kPI / 4, c); cx(a, c); u1(kPI / 4, c); cx(b, c);
        u1(-kPI / 4, c); u2(0, kPI, c);
    }
    void rc3x(int a, int b, int c, int d)
    {
        u2(0, kPI, d); u1(kPI / 4, d); cx(c, d); u1(-kPI / 4, d); u2(0, kPI, d); cx(a, d); u1(kPI / 4, d);
        cx(b, d); u1(-kPI / 4, d); cx(a, d); u1(kPI / 4, d); cx(b, d); u1(-kPI / 4, d); u2(0, kPI, d);
        u1(kPI / 4, d); cx(c, d); u1(-kPI / 4, d); u2(0, kPI, d);
    }
    void hcu1h(double ang, int ctl, int d) { h(d); cu1(ang, ctl, d); h(d); }
    void c3x_like(double ang, int a, int b, int c, int d) // C3X :1698-1728, C3SQRTX :1733-1763
    {
        hcu1h(-ang, a, d); cx(a, b); hcu1h(ang, b, d); cx(a, b); hcu1h(-ang, b, d); cx(b, c);
        hcu1h(ang, c, d); cx(a, c); hcu1h(-ang, c, d); cx(b, c); hcu1h(ang, c, d); cx(a, c);
        hcu1h(-ang, c, d);
    }
    void c4x(int a, int b, int c, int d, int e)
    {
        h(e); cu1(-kPI / 2, d, e); h(e);
        c3x_like(kPI / 4, a, b, c, d);
        h(d); cu1(kPI / 4, d, e); h(d);
        c3x_like(kPI / 4, a, b, c, d);
        c3x_like(kPI / 8, a, b, c, e);
    }
    void ryy(double th, int a, int b)
    {
        rx(kPI / 2, a); rx(kPI / 2, b); cx(a, b); rz(th, b); cx(a, b); rx(-kPI / 2, a); rx(-kPI / 2, b);
    }
};

int arity(int op)
{
    switch (op)
    {
    case DMB_OP_CX: case DMB_OP_CZ: case DMB_OP_CY: case DMB_OP_SWAP: case DMB_OP_CH: case DMB_OP_CRX:
    case DMB_OP_CRY: case DMB_OP_CRZ: case DMB_OP_CU1: case DMB_OP_CU3: case DMB_OP_RXX: case DMB_OP_RZZ:
    case DMB_OP_RYY: case DMB_OP_C2:
        return 2;
    case DMB_OP_CCX: case DMB_OP_CSWAP: case DMB_OP_RCCX:
        return 3;
    case DMB_OP_RC3X: case DMB_OP_C3X: case DMB_OP_C3SQRTX:
        return 4;
    case DMB_OP_C4X:
        return 5;
    default:
        return 1;
    }
}
} // namespace

void expand_gates(int n_qubits, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
                  std::vector<Block>& prims)
{
    Expander E{n_qubits, prims};
    for (size_t i = 0; i < n_gates; i++)
    {
        const dmb_gate& g = gates[i];
        const bool known = (g.op >= 0 && g.op < DMB_OP_COUNT) || g.op == DMB_OP_C1 || g.op == DMB_OP_C2;
        if (!known) throw std::invalid_argument("gate " + std::to_string(i) + ": unknown op " + std::to_string(g.op));
        // append() asserts every qb < n_qubits (reference :334-338), used or not
        for (int k = 0; k < 5; k++)
            if (g.qb[k] < 0 || g.qb[k] >= n_qubits)
                throw std::invalid_argument("gate " + std::to_string(i) + " (" + op_name(g.op) + "): qubit index " +
                                            std::to_string(g.qb[k]) + " out of range");
        const int ar = arity(g.op);
        for (int a = 0; a < ar; a++)
            for (int b = a + 1; b < ar; b++)
                if (g.qb[a] == g.qb[b])
                    throw std::invalid_argument("gate " + std::to_string(i) + " (" + op_name(g.op) +
                                                "): repeated qubit operand"); // reference asserts ctrl != qubit (:1051,:1139)
        const int q0 = g.qb[0], q1 = g.qb[1], q2 = g.qb[2], q3 = g.qb[3], q4 = g.qb[4];
        // which Gate field feeds which parameter: *_OP wrappers :1821-2008
        switch (g.op)
        {
        case DMB_OP_U3: E.u3(g.theta, g.phi, g.lambda, q0); break;
        case DMB_OP_U2: E.u2(g.phi, g.lambda, q0); break;
        case DMB_OP_U1: E.u1(g.lambda, q0); break;
        case DMB_OP_CX: E.cx(q0, q1); break;
        case DMB_OP_ID: break;
        case DMB_OP_X: E.x(q0); break;
        case DMB_OP_Y: E.y(q0); break;
        case DMB_OP_Z: E.z(q0); break;
        case DMB_OP_H: E.h(q0); break;
        case DMB_OP_S: E.s(q0); break;
        case DMB_OP_SDG: E.sdg(q0); break;
        case DMB_OP_T: E.t(q0); break;
        case DMB_OP_TDG: E.tdg(q0); break;
        case DMB_OP_RX: E.rx(g.theta, q0); break;
        case DMB_OP_RY: E.ry(g.theta, q0); break;
        case DMB_OP_RZ: E.rz(g.phi, q0); break;
        case DMB_OP_CZ: E.cz(q0, q1); break;
        case DMB_OP_CY: E.cy(q0, q1); break;
        case DMB_OP_SWAP: E.swap(q0, q1); break;
        case DMB_OP_CH: E.ch(q0, q1); break;
        case DMB_OP_CCX: E.ccx(q0, q1, q2); break;
        case DMB_OP_CSWAP: E.cswap(q0, q1, q2); break;
        case DMB_OP_CRX: E.crx(g.lambda, q0, q1); break;
        case DMB_OP_CRY: E.cry(g.lambda, q0, q1); break;
        case DMB_OP_CRZ: E.crz(g.lambda, q0, q1); break;
        case DMB_OP_CU1: E.cu1(g.lambda, q0, q1); break;
        case DMB_OP_CU3: E.cu3(g.theta, g.phi, g.lambda, q0, q1); break;
        case DMB_OP_RXX: E.rxx(g.theta, q0, q1); break;
        case DMB_OP_RZZ: E.rzz(g.theta, q0, q1); break;
        case DMB_OP_RCCX: E.rccx(q0, q1, q2); break;
        case DMB_OP_RC3X: E.rc3x(q0, q1, q2, q3); break;
        case DMB_OP_C3X: E.c3x_like(kPI / 4, q0, q1, q2, q3); break;
        case DMB_OP_C3SQRTX: E.c3x_like(kPI / 8, q0, q1, q2, q3); break;
        case DMB_OP_C4X: E.c4x(q0, q1, q2, q3, q4); break;
        case DMB_OP_R: E.r(g.theta, q0); break;
        case DMB_OP_SRN: E.srn(q0); break;
        case DMB_OP_W: E.w(q0); break;
        case DMB_OP_RYY: E.ryy(g.theta, q0, q1); break;
        case DMB_OP_C1:
        case DMB_OP_C2:
        {
            if (!mats || g.mat < 0 || (size_t)g.mat >= n_mats)
                throw std::invalid_argument("gate " + std::to_string(i) + ": matrix index out of range");
            const double* m = mats + 32 * (size_t)g.mat;
            Block k;
            k.nq = g.op == DMB_OP_C1 ? 1 : 2;
            k.q[0] = q0;
            k.q[1] = q1;
            const int cnt = k.nq == 1 ? 4 : 16;
            for (int e = 0; e < cnt; e++) k.m[e] = cplx(m[2 * e], m[2 * e + 1]);
            prims.push_back(k);
            break;
        }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fusion
// ------------------------------------------------------------------------------------------------
namespace
{
void mat4_mul(const cplx* a, const cplx* b, cplx* c) // c = a*b (4x4)
{
    cplx t[16];
    for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++)
        {
            cplx s = 0;
            for (int k = 0; k < 4; k++) s += a[r * 4 + k] * b[k * 4 + col];
            t[r * 4 + col] = s;
        }
    memcpy(c, t, sizeof(t));
}
void mat2_mul(const cplx* a, const cplx* b, cplx* c)
{
    cplx t[4];
    t[0] = a[0] * b[0] + a[1] * b[2];
    t[1] = a[0] * b[1] + a[1] * b[3];
    t[2] = a[2] * b[0] + a[3] * b[2];
    t[3] = a[2] * b[1] + a[3] * b[3];
    memcpy(c, t, sizeof(t));
}
// 2x2 u acting on the MSB (hi=true) or LSB factor of a 2-qubit index, as a 4x4
void embed1(const cplx* u, bool hi, cplx* o)
{
    for (int i = 0; i < 16; i++) o[i] = 0;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++)
        {
            int rh = r >> 1, rl = r & 1, ch = c >> 1, cl = c & 1;
            if (hi) { if (rl == cl) o[r * 4 + c] = u[rh * 2 + ch]; }
            else    { if (rh == ch) o[r * 4 + c] = u[rl * 2 + cl]; }
        }
}
// same operator with the two qubits' roles exchanged (index bit swap)
void swap_roles(const cplx* m, cplx* o)
{
    static const int p[4] = {0, 2, 1, 3};
    cplx t[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) t[p[r] * 4 + p[c]] = m[r * 4 + c];
    memcpy(o, t, sizeof(t));
}
bool is_identity2(const cplx* u) { return u[0] == cplx(1) && u[3] == cplx(1) && u[1] == cplx(0) && u[2] == cplx(0); }
} // namespace

// Peephole on the primitive list.  With H X H = Z inserted pairwise (H CX(c,t) H = CZ(c,t) for H on the TARGET wire t):
//   H_t  CX(c1,t) ... CX(ck,t)  H_t   ==  CZ(c1,t) ... CZ(ck,t)                 (both Hadamards vanish)
//   H_t  CX(c1,t) ... CX(ck,t)        ==  CZ(c1,t) ... CZ(ck,t)  H_t            (the Hadamard moves behind a fan with
//                                                                                 at least two different controls; off by
//                                                                                 default, option "move_h")
// when nothing else touches t in between.  The CZs are diagonal: they become controlled phases that need only ONE of
// their bits in a tile / register round, instead of register permutations that need both (Bernstein-Vazirani, parity /
// fan-in circuits; the reference builds CZ itself as H CX H, :1493-1499).  Exact operator identities (H H = 1 up to one
// rounding of 2 * S2I^2), valid for any state; circuits with SRN are left alone.
void rewrite_hadamard_cx(int n, std::vector<Block>& prims, bool move_h)
{
    for (const Block& p : prims)
        if (p.srn) return;
    auto is_h = [](const Block& b) {
        return b.nq == 1 && b.m[0] == cplx(kS2I, 0) && b.m[1] == cplx(kS2I, 0) && b.m[2] == cplx(kS2I, 0) && b.m[3] == cplx(-kS2I, 0);
    };
    auto is_cx_onto = [](const Block& b, int t) { // CX exactly as Expander::cx builds it, target t
        if (b.nq != 2 || b.q[1] != t) return false;
        for (int i = 0; i < 16; i++)
            if (b.m[i] != cplx((i == 0 || i == 5 || i == 11 || i == 14) ? 1.0 : 0.0, 0.0)) return false;
        return true;
    };
    std::vector<char> drop(prims.size(), 0);
    std::vector<std::pair<int, Block>> moved;   // Hadamards re-inserted AFTER primitive index .first
    std::vector<int> open_h(n, -1);             // index of a Hadamard on this wire that may start a pattern
    std::vector<std::vector<int>> fan(n);       // the CXs onto the wire since that Hadamard
    auto to_cz = [&](int t, int extra_weight) {
        for (int j : fan[t])
        {
            Block& c = prims[j];
            for (auto& e : c.m) e = 0;
            c.m[0] = c.m[5] = c.m[10] = 1;
            c.m[15] = -1;
            c.weight += extra_weight; // keep the primitive count of the sweep statistics
            extra_weight = 0;
        }
    };
    auto end_pattern = [&](int t) { // something else arrives on wire t (or the circuit ends): second identity
        // The second identity,  H_t CX(c1,t) .. CX(ck,t) == CZ(c1,t) .. CZ(ck,t) H_t,  is implemented (opt.move_h) but off:
        // measured on bv_n15 it trades k register permutations for k single-bit stars at about the same cost (45.2 vs
        // 43.0 ms) and makes the sparse start less effective (the target bit enters the first sweep).
        bool real_fan = false;
        for (size_t j = 1; j < fan[t].size(); j++)
            if (prims[fan[t][j]].q[0] != prims[fan[t][0]].q[0]) real_fan = true;
        if (move_h && open_h[t] >= 0 && real_fan)
        {
            to_cz(t, 0);
            moved.push_back({fan[t].back(), prims[open_h[t]]});
            drop[open_h[t]] = 1;
        }
        open_h[t] = -1;
        fan[t].clear();
    };
    for (size_t i = 0; i < prims.size(); i++)
    {
        Block& b = prims[i];
        if (is_h(b))
        {
            const int t = b.q[0];
            if (open_h[t] >= 0 && !fan[t].empty()) // first identity
            {
                to_cz(t, prims[open_h[t]].weight + b.weight);
                drop[open_h[t]] = 1;
                drop[i] = 1;
                open_h[t] = -1;
                fan[t].clear();
                continue;
            }
            open_h[t] = (int)i;
            fan[t].clear();
            continue;
        }
        for (int k = 0; k < b.nq; k++)
        {
            const int q = b.q[k];
            if (open_h[q] < 0) continue;
            if (is_cx_onto(b, q)) fan[q].push_back((int)i);
            else end_pattern(q); // (also being the CONTROL of a CX: diagonal on this wire, it does not commute with H)
        }
    }
    for (int t = 0; t < n; t++) end_pattern(t);
    std::stable_sort(moved.begin(), moved.end(), [](const std::pair<int, Block>& a, const std::pair<int, Block>& b) { return a.first < b.first; });
    std::vector<Block> out;
    out.reserve(prims.size());
    size_t mi = 0;
    for (size_t i = 0; i < prims.size(); i++)
    {
        if (!drop[i]) out.push_back(prims[i]);
        while (mi < moved.size() && moved[mi].first == (int)i) out.push_back(moved[mi++].second);
    }
    prims.swap(out);
}

void fuse_blocks(int n, const std::vector<Block>& prims, std::vector<Block>& blocks, bool split_cphase)
{
    struct Pend { bool have = false; cplx u[4]; int weight = 0; };
    std::vector<Pend> pend(n);
    std::vector<int> open(n, -1);       // index into work[] of the open 2-qubit block on this qubit
    std::vector<Block> work;

    bool any_srn = false;
    for (const Block& p : prims) any_srn |= p.srn;
    auto close_block = [&](int idx) {
        if (idx < 0) return;
        Block b = work[idx];
        open[b.q[0]] = -1;
        open[b.q[1]] = -1;
        // A diagonal block diag(d0, d1, d2, d3) = d0 * diag_q0(1, d2/d0) * diag_q1(1, d1/d0) * CP(d0 d3 / (d1 d2)): emit
        // the pure controlled phase (the schedulers only need ONE of its bits in a tile, CLS_CPHASE) and leave the
        // 1-qubit phases pending on their qubits -- they commute with it and merge into whatever follows there.
        if (split_cphase && !any_srn && classify(2, b.m, nullptr) == CLS_DIAG2)
        {
            const cplx d0 = b.m[0], d1 = b.m[5], d2 = b.m[10], d3 = b.m[15];
            const double tiny = 1e-8;
            if (std::abs(d0) > tiny && std::abs(d1) > tiny && std::abs(d2) > tiny)
            {
                const cplx u0[4] = {d0, 0, 0, d2}, u1[4] = {1, 0, 0, d1 / d0}; // d0 rides on q0's factor
                const cplx phi = d0 * d3 / (d1 * d2);
                for (auto& e : b.m) e = 0;
                b.m[0] = b.m[5] = b.m[10] = 1;
                b.m[15] = phi;
                const bool trivial = std::abs(phi.real() - 1.0) < 1e-15 && std::abs(phi.imag()) < 1e-15;
                if (!trivial) blocks.push_back(b);
                const cplx* us[2] = {u0, u1};
                for (int side = 0; side < 2; side++)
                {
                    const int q = b.q[side];
                    if (is_identity2(us[side])) continue;
                    if (pend[q].have) mat2_mul(us[side], pend[q].u, pend[q].u);
                    else
                    {
                        pend[q].have = true;
                        memcpy(pend[q].u, us[side], sizeof(cplx) * 4);
                        pend[q].weight = 0;
                    }
                }
                return;
            }
        }
        blocks.push_back(b);
    };
    auto flush_pending = [&](int q) {
        if (!pend[q].have) return;
        Block b;
        b.nq = 1; b.q[0] = q; b.weight = pend[q].weight;
        memcpy(b.m, pend[q].u, sizeof(cplx) * 4);
        blocks.push_back(b);
        pend[q].have = false; pend[q].weight = 0;
    };

    for (const Block& p : prims)
    {
        if (p.srn)
        {
            // SRN conjugates amplitudes (v0' = (v0 + conj(v1))/2), so it does not commute with complex-linear
            // ops on OTHER qubits either: it is a barrier for the whole circuit.
            for (int q = 0; q < n; q++)
            {
                close_block(open[q]);
                flush_pending(q);
            }
            blocks.push_back(p);
            continue;
        }
        if (p.nq == 1)
        {
            const int q = p.q[0];
            // structure-preserving: a dense 1-qubit gate (H, U3, ...) is NOT folded into a diagonal / monomial 2-qubit
            // block (CX, controlled phases): the block would become a dense 4x4 (4x the FP64 work) and diagonal blocks
            // would lose their freedom to commute in the schedulers
            if (open[q] >= 0 && classify(1, p.m, nullptr) == CLS_DENSE1 &&
                classify(2, work[open[q]].m, nullptr) != CLS_DENSE2)
                close_block(open[q]);
            // likewise a monomial 1-qubit gate (X, Y) stays out of a DIAGONAL 2-qubit block: the block would stop being
            // a controlled phase (which needs only one of its bits in a tile)
            if (open[q] >= 0 && classify(1, p.m, nullptr) != CLS_DIAG1 && classify(2, work[open[q]].m, nullptr) == CLS_DIAG2)
                close_block(open[q]);
            if (open[q] >= 0)
            {
                Block& b = work[open[q]];
                cplx e[16];
                embed1(p.m, b.q[0] == q, e);
                mat4_mul(e, b.m, b.m);
                b.weight += p.weight;
            }
            else if (pend[q].have)
            {
                mat2_mul(p.m, pend[q].u, pend[q].u);
                pend[q].weight += p.weight;
            }
            else
            {
                pend[q].have = true;
                memcpy(pend[q].u, p.m, sizeof(cplx) * 4);
                pend[q].weight = p.weight;
            }
            continue;
        }
        const int a = p.q[0], b_ = p.q[1];
        if (open[a] >= 0 && open[a] == open[b_])
        {
            Block& b = work[open[a]];
            cplx e[16];
            if (b.q[0] == a) memcpy(e, p.m, sizeof(e));
            else swap_roles(p.m, e);
            mat4_mul(e, b.m, b.m);
            b.weight += p.weight;
            continue;
        }
        close_block(open[a]);
        close_block(open[b_]);
        Block nb = p;
        const bool nb_dense = classify(2, p.m, nullptr) == CLS_DENSE2, nb_diag = classify(2, p.m, nullptr) == CLS_DIAG2;
        for (int side = 0; side < 2; side++)
        {
            const int q = side == 0 ? a : b_;
            if (!pend[q].have) continue;
            const int pc = classify(1, pend[q].u, nullptr);
            if ((!nb_dense && pc == CLS_DENSE1) || (nb_diag && pc != CLS_DIAG1))
            {
                flush_pending(q); // keep the 1-qubit product as its own block (see above)
                continue;
            }
            if (!is_identity2(pend[q].u))
            {
                cplx e[16];
                embed1(pend[q].u, side == 0, e);
                mat4_mul(nb.m, e, nb.m);
            }
            nb.weight += pend[q].weight;
            pend[q].have = false; pend[q].weight = 0;
        }
        work.push_back(nb);
        open[a] = open[b_] = (int)work.size() - 1;
    }
    for (int q = 0; q < n; q++)
        if (open[q] >= 0) close_block(open[q]); // may leave 1-qubit phases pending on either qubit
    for (int q = 0; q < n; q++) flush_pending(q);
}

int classify(int nb, const cplx* m, int* src_out)
{
    int src[4] = {0, 1, 2, 3};
    const double eps = 1e-15;
    auto nz = [&](cplx v) { return std::abs(v.real()) > eps || std::abs(v.imag()) > eps; };
    const int d = nb == 1 ? 2 : 4;
    bool diag = true, mono = true;
    for (int r = 0; r < d; r++)
    {
        int cnt = 0, where = -1;
        for (int c = 0; c < d; c++)
            if (nz(m[r * d + c]))
            {
                cnt++; where = c;
                if (c != r) diag = false;
            }
        if (cnt > 1) mono = false;
        if (cnt == 0) where = r; // zero row: treated as a zero phase on itself
        src[r] = where;
    }
    if (mono)
    {
        // must be a permutation of columns (a zero row/col would make it singular; treat as dense then)
        int seen = 0;
        for (int r = 0; r < d; r++) seen |= 1 << src[r];
        if (seen != (1 << d) - 1) mono = false;
    }
    if (src_out)
        for (int r = 0; r < d; r++) src_out[r] = src[r];
    if (nb == 1) return diag ? CLS_DIAG1 : (mono ? CLS_MONO1 : CLS_DENSE1);
    return diag ? CLS_DIAG2 : (mono ? CLS_MONO2 : CLS_DENSE2);
}

// ------------------------------------------------------------------------------------------------
// Scheduling on the 2n-bit flat index
// ------------------------------------------------------------------------------------------------
namespace
{
// (the matrices live in a side array: the tile-selection scans walk thousands of these per candidate, and 32-byte
// records keep them in L1 -- with the 256-byte matrix inline a 10^4-gate circuit spent most of its planning time on
// cache misses)
struct FlatOp
{
    int nb;
    int bit[2]; // logical bits of the 2n-bit index; bit[0] carries the matrix MSB
    bool srn;
    int weight;
    int side; // 0 = L (row bits), 1 = R (column bits)
    bool diag = false; // diagonal matrix: commutes with every other diagonal op
    bool cp = false;   // diag(1, 1, 1, phi): needs only ONE of its bits inside the tile (see CLS_CPHASE)
    bool done = false;
};

struct Candidate
{
    std::vector<int> tile_logical; // logical bits in the tile, in order of admission
    std::vector<int> picked;       // indices into ops, in execution order
    long score = 0;
};
} // namespace

Plan make_plan(int n, int world_size, const dmb_gate* gates, size_t n_gates, const double* mats, size_t n_mats,
               const std::vector<int>& start_layout, const PlanOptions& opt, bool conj_state, bool non_hermitian)
{
    if (n < 1 || n > 20) throw std::invalid_argument("n_qubits must be in [1, 20]");
    int g = 0;
    while ((1 << g) < world_size) g++;
    if ((1 << g) != world_size) throw std::invalid_argument("world_size must be a power of two");
    const int N = 2 * n, M = N - g;
    if (g > n) throw std::invalid_argument("world_size must divide 2^n_qubits"); // reference :218-229
    if (g > 0 && M - g < 0) throw std::invalid_argument("too many ranks for this state");

    Plan plan;
    plan.n = n; plan.g = g;
    plan.n_gates = n_gates;
    std::vector<Block> prims, blocks;
    expand_gates(n, gates, n_gates, mats, n_mats, prims);
    plan.n_primitives = prims.size();
    if (opt.cphase) rewrite_hadamard_cx(n, prims, opt.move_h);
    fuse_blocks(n, prims, blocks, opt.cphase);
    plan.n_blocks = blocks.size();

    // mirror: L part on bit q, R part (conjugated) on bit q+n.  SRN is real-linear and self-conjugate.
    std::vector<FlatOp> ops;
    struct Mat16 { cplx m[16]; };
    std::vector<Mat16> op_m; // op_m[i].m = matrix of ops[i] (already conjugated for R parts)
    ops.reserve(blocks.size() * 2);
    op_m.reserve(blocks.size() * 2);
    for (const Block& b : blocks)
        if (b.srn) plan.has_srn = true;
    // With SRN in the circuit the run is strictly "all L parts, then all R parts" in program order
    // (X = F_R F_L M0 with F_R = T F T, T = conjugate transpose), every SRN being a full barrier.
    for (int pass = 0; pass < (plan.has_srn ? 2 : 1); pass++)
    for (const Block& b : blocks)
    {
        for (int side = (plan.has_srn ? pass : 0); side < (plan.has_srn ? pass + 1 : 2); side++)
        {
            FlatOp f;
            Mat16 fm;
            cplx* const f_m = fm.m;
            f.nb = b.nq; f.srn = b.srn; f.weight = b.weight; f.side = side;
            f.bit[0] = b.q[0] + side * n;
            f.bit[1] = b.nq == 2 ? b.q[1] + side * n : 0;
            for (auto& e : fm.m) e = 0;
            const int cnt = b.nq == 1 ? 4 : 16;
            // on a conjugated store (conj_state) E acts as conj(E): conj(E conj(x)) = conj(E) x
            for (int e = 0; e < cnt; e++) f_m[e] = ((side == 1) != conj_state) ? std::conj(b.m[e]) : b.m[e];
            if (!b.srn)
            {
                const int cls = classify(b.nq, f_m, nullptr);
                f.diag = !plan.has_srn && (cls == CLS_DIAG1 || cls == CLS_DIAG2);
                if (f.diag && cls == CLS_DIAG2 && opt.cphase)
                {
                    // entries within 1e-15 of 1 (u1(a) u1(-a) inside a fused controlled phase) are exactly 1
                    auto is_one = [](cplx v) { return std::abs(v.real() - 1.0) < 1e-15 && std::abs(v.imag()) < 1e-15; };
                    if (is_one(f_m[0]) && is_one(f_m[5]) && is_one(f_m[10]))
                    {
                        f.cp = true;
                        f_m[0] = f_m[5] = f_m[10] = cplx(1.0, 0.0);
                    }
                }
            }
            ops.push_back(f);
            op_m.push_back(fm);
        }
    }

    std::vector<int> phys(N);
    if (start_layout.empty())
        for (int l = 0; l < N; l++) phys[l] = l;
    else
    {
        if ((int)start_layout.size() != N) throw std::invalid_argument("start_layout must have 2n entries");
        phys = start_layout;
    }
    plan.start_layout = phys;

    const int kmax = std::max(1, std::min(std::min(opt.tile_bits, 12), M));
    const int lowb = std::max(0, std::min(opt.low_bits, kmax));

    size_t n_pending = ops.size();
    size_t first_pending = 0;
    std::vector<int> logical_at(N);

    auto build_candidate = [&](int strategy) {
        Candidate c;
        std::vector<char> in_tile(N, 0), blocked(N, 0), is_rank(N, 0);
        for (int l = 0; l < N; l++)
        {
            if (phys[l] >= M) is_rank[l] = 1; // rank bits: not addressable inside a shard
            if (phys[l] < lowb) { in_tile[l] = 1; c.tile_logical.push_back(l); }
        }
        int tile_cnt = (int)c.tile_logical.size();
        long picked_cost = 0;
        int picked_cp = 0;
        int n_free_bits = 0;
        for (int l = 0; l < N; l++) n_free_bits += !is_rank[l];
        // blocked[bit]: 0 free, 1 only DIAGONAL ops were skipped on this bit (a later diagonal op commutes with all of
        // them and may still run in this sweep), 2 closed
        auto visit = [&](size_t i) {
            FlatOp& f = ops[i];
            if (f.done) return;
            bool blk = false;
            int add[2], n_add = 0;
            if (f.cp)
            {
                // a controlled phase runs as soon as ONE of its bits is in the tile; the other one is only read
                bool any_in = false;
                for (int b = 0; b < 2; b++)
                {
                    if (blocked[f.bit[b]] == 2) blk = true;
                    if (in_tile[f.bit[b]]) any_in = true;
                }
                if (!blk && !any_in)
                {
                    if (!is_rank[f.bit[0]]) add[n_add++] = f.bit[0];
                    else if (!is_rank[f.bit[1]]) add[n_add++] = f.bit[1];
                    else blk = true;
                }
            }
            else
                for (int b = 0; b < f.nb; b++)
                {
                    const int lvl = blocked[f.bit[b]];
                    if (is_rank[f.bit[b]] || lvl == 2 || (lvl == 1 && !f.diag)) blk = true;
                    else if (!in_tile[f.bit[b]]) add[n_add++] = f.bit[b];
                }
            if (!blk && tile_cnt + n_add <= kmax && picked_cost + (f.cp ? 1 : 8) <= 8 * opt.max_ops &&
                (!f.cp || picked_cp < opt.max_cphase))
            {
                picked_cost += f.cp ? 1 : 8; // controlled phases merge into star ops: they hardly use op-table space
                picked_cp += f.cp ? 1 : 0;   // ... but every one of them may need its own star slot
                for (int b = 0; b < n_add; b++) { in_tile[add[b]] = 1; c.tile_logical.push_back(add[b]); tile_cnt++; }
                c.picked.push_back((int)i);
                c.score += 1000L * f.weight + 1; // zero-weight ops (1-qubit phases split off a controlled phase) still count as progress
            }
            else
                for (int b = 0; b < f.nb; b++)
                {
                    const char lvl = f.diag ? 1 : 2;
                    if (blocked[f.bit[b]] < lvl)
                    {
                        if (lvl == 2 && !is_rank[f.bit[b]]) n_free_bits--;
                        blocked[f.bit[b]] = lvl;
                    }
                }
        };
        // strategy 0: program order (L and R interleaved); 1: all L parts first; 2: all R parts first
        if (plan.has_srn)
        {
            // strict sequence order; an SRN may only run once everything before it has, and fences what follows
            bool all_prior = true;
            for (size_t i = first_pending; i < ops.size() && n_free_bits > 0; i++)
            {
                if (ops[i].done) continue;
                if (ops[i].srn && !all_prior) break;
                const size_t before = c.picked.size();
                visit(i);
                const bool took = c.picked.size() > before;
                if (!took) all_prior = false;
                if (ops[i].srn && !took) break;
            }
        }
        else if (strategy == 0)
        {
            const size_t end = std::min(ops.size(), first_pending + (size_t)opt.scan_window);
            for (size_t i = first_pending; i < end && n_free_bits > 0; i++) visit(i);
        }
        else
        {
            const int first_side = strategy == 1 ? 0 : 1;
            for (int pass = 0; pass < 2; pass++)
            {
                const int side = pass == 0 ? first_side : 1 - first_side;
                const size_t end = std::min(ops.size(), first_pending + (size_t)opt.scan_window);
                for (size_t i = first_pending; i < end && n_free_bits > 0; i++)
                    if (ops[i].side == side) visit(i);
            }
        }
        return c;
    };

    auto emit_sweep = [&](const std::vector<int>& tile_logical_in, const std::vector<int>& picked,
                          const std::vector<std::pair<int, int>>& moves /* logical -> new phys */, bool in_place = false) {
        // pad the tile with the lowest free physical bits: longer contiguous runs, fewer larger tiles --
        // but keep at least 2^min_tiles_log2 tiles when the state is small.
        std::vector<int> tile_logical = tile_logical_in;
        std::vector<char> used(N, 0);
        for (int l : tile_logical) used[phys[l]] = 1;
        for (int l = 0; l < N; l++) logical_at[phys[l]] = l;
        int want = std::max((int)tile_logical.size(), std::min(kmax, M - opt.min_tiles_log2));
        want = std::min(want, M);
        for (int p = 0; p < M && (int)tile_logical.size() < want; p++)
            if (!used[p]) { used[p] = 1; tile_logical.push_back(logical_at[p]); }
        Step st;
        st.kind = 0;
        Sweep& sw = st.sweep;
        sw.k = (int)tile_logical.size();
        std::vector<int> order = tile_logical;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return phys[a] < phys[b]; });
        std::vector<int> local_of(N, -1);
        for (int j = 0; j < sw.k; j++) { local_of[order[j]] = j; sw.in_pos.push_back(phys[order[j]]); }
        // L parts (row bits) and R parts (column bits) commute: run all L parts first, so that each half leaves
        // the other half's tile bits untouched (long barrier-free warp groups in the kernel)
        std::vector<int> ordered;
        for (int side = 0; side < 2; side++)
            for (int i : picked)
                if (ops[i].side == side) ordered.push_back(i);
        for (int i : ordered)
        {
            FlatOp& f = ops[i];
            TileOp t;
            t.nb = f.nb;
            t.j0 = local_of[f.bit[0]];
            t.j1 = f.nb == 2 ? local_of[f.bit[1]] : 0;
            t.weight = f.weight;
            memcpy(t.m, op_m[i].m, sizeof(t.m));
            t.cls = f.srn ? (int)CLS_SRN1 : classify(f.nb, op_m[i].m, nullptr);
            if (f.cp)
            {
                t.cls = CLS_CPHASE; // symmetric in its two bits: keep the in-tile one first
                if (t.j0 < 0) { std::swap(t.j0, t.j1); t.p1 = phys[f.bit[0]]; }
                else if (t.j1 < 0) t.p1 = phys[f.bit[1]];
                if (t.j0 < 0) throw std::logic_error("controlled phase scheduled without a tile bit");
            }
            sw.ops.push_back(t);
            sw.weight += f.weight;
            f.done = true;
            n_pending--;
        }
        for (auto& mv : moves) phys[mv.first] = mv.second;
        for (int j = 0; j < sw.k; j++) sw.out_pos.push_back(phys[order[j]]);
        // (L2-resident states: no TMA -- every sweep of the run then fits the one-launch cooperative executor, capi.cu)
        sw.swz_mode = (opt.tma && sw.k == 12 && lowb >= 3 && M > opt.small_state_bits) ? 1 : 0;
        sw.out_of_place = !moves.empty() && !in_place; // a permutation INSIDE the tile may run in place: every tile is read
                                                       // completely before it is written, to the same set of addresses
        plan.steps.push_back(st);
        plan.n_sweeps++;
        while (first_pending < ops.size() && ops[first_pending].done) first_pending++;
    };

    // ---- gain-driven tile selection --------------------------------------------------------------------------
    // scan_fixed: with the tile bit set FIXED, which pending ops can run (program order per bit, diagonal ops may hop
    // over skipped diagonal ops)?  Returns the total weight; fills picked when asked.
    // (bit sets as 64-bit masks: `tile` = bits in the tile, `b2` = bits blocked by a skipped non-diagonal op, `b1` = by a
    // skipped diagonal op; a visit is a handful of mask operations)
    std::vector<unsigned long long> op_bits(ops.size());
    for (size_t i = 0; i < ops.size(); i++)
        op_bits[i] = (1ull << ops[i].bit[0]) | (ops[i].nb == 2 ? 1ull << ops[i].bit[1] : 0ull);
    auto scan_fixed = [&](const std::vector<char>& in_tile, std::vector<int>* picked) -> long {
        unsigned long long tile = 0, b1 = 0, b2 = 0;
        for (int l = 0; l < N; l++)
            if (in_tile[l]) tile |= 1ull << l;
        long score = 0;
        int n_picked = 0, n_cp = 0;
        // (a sweep holds at most max_ops ops: looking further than a few thousand pending ops ahead only costs time --
        // the scans of a 10^4-gate circuit were quadratic without the window)
        int visited = 0;
        for (size_t i = first_pending; i < ops.size() && (tile & ~b2) && visited < opt.scan_window; i++)
        {
            const FlatOp& f = ops[i];
            if (f.done) continue;
            if (n_picked >= 8 * opt.max_ops) break; // the sweep's op table is full: nothing further can be picked
            visited++;
            const unsigned long long bm = op_bits[i];
            bool ok = n_picked + (f.cp ? 1 : 8) <= 8 * opt.max_ops && (!f.cp || n_cp < opt.max_cphase);
            if (f.cp) ok = ok && !(bm & b2) && (bm & tile);
            else ok = ok && !(bm & ~tile) && !(bm & b2) && (f.diag || !(bm & b1));
            if (ok)
            {
                score += 1000L * f.weight + 1;
                n_picked += f.cp ? 1 : 8;
                n_cp += f.cp ? 1 : 0;
                if (picked) picked->push_back((int)i);
                continue;
            }
            if (f.diag) b1 |= bm;
            else b2 |= bm;
        }
        return score;
    };
    auto build_by_gain = [&]() {
        Candidate c;
        std::vector<char> in_tile(N, 0);
        for (int l = 0; l < N; l++)
            if (phys[l] < lowb) { in_tile[l] = 1; c.tile_logical.push_back(l); }
        long cur = scan_fixed(in_tile, nullptr);
        while ((int)c.tile_logical.size() < kmax)
        {
            const int room = kmax - (int)c.tile_logical.size();
            // candidate additions: the missing bits of the pending ops near the front (deduplicated; bit sets as 64-bit
            // masks in first-seen order -- small vectors here cost a 10^4-gate circuit most of its planning time)
            std::vector<unsigned long long> cands;
            auto add_cand = [&](unsigned long long m) {
                if (std::find(cands.begin(), cands.end(), m) == cands.end()) cands.push_back(m);
            };
            int looked = 0;
            for (size_t i = first_pending; i < ops.size() && looked < 512; i++)
            {
                const FlatOp& f = ops[i];
                if (f.done) continue;
                looked++;
                if (f.cp)
                {
                    // one bit in the tile is enough: each local bit is a candidate on its own
                    if (in_tile[f.bit[0]] || in_tile[f.bit[1]]) continue;
                    for (int b = 0; b < 2; b++)
                        if (phys[f.bit[b]] < M) add_cand(1ull << f.bit[b]);
                    continue;
                }
                unsigned long long miss = 0;
                bool local = true;
                for (int b = 0; b < f.nb; b++)
                {
                    if (phys[f.bit[b]] >= M) local = false;
                    if (!in_tile[f.bit[b]]) miss |= 1ull << f.bit[b];
                }
                if (!local || !miss || __builtin_popcountll(miss) > room) continue;
                add_cand(miss);
            }
            long best_gain = 0;
            double best_rate = 0;
            int best_i = -1;
            for (size_t ci = 0; ci < cands.size(); ci++)
            {
                for (int l = 0; l < N; l++)
                    if ((cands[ci] >> l) & 1ull) in_tile[l] = 1;
                const long gain = scan_fixed(in_tile, nullptr) - cur;
                for (int l = 0; l < N; l++)
                    if ((cands[ci] >> l) & 1ull) in_tile[l] = 0;
                const double rate = (double)gain / (double)__builtin_popcountll(cands[ci]);
                if (gain > 0 && (rate > best_rate || (rate == best_rate && gain > best_gain)))
                {
                    best_rate = rate; best_gain = gain; best_i = (int)ci;
                }
            }
            if (best_i < 0) break;
            for (int l = 0; l < N; l++)
                if ((cands[best_i] >> l) & 1ull) { in_tile[l] = 1; c.tile_logical.push_back(l); }
            cur += best_gain;
        }
        c.score = scan_fixed(in_tile, &c.picked);
        return c;
    };

    while (n_pending > 0)
    {
        Candidate best;
        bool have = false;
        if (!plan.has_srn)
        {
            best = build_by_gain();
            have = true;
        }
        for (int s = 0; s < (plan.has_srn ? 1 : 3); s++)
        {
            Candidate c = build_candidate(s);
            if (!have || c.score > best.score ||
                (c.score == best.score && c.tile_logical.size() < best.tile_logical.size()))
            {
                best = c; have = true;
            }
        }
        if (!best.picked.empty())
        {
            // Hot bits to the low positions.  The lowest `lowb` physical bits belong to EVERY tile (coalescing), so a
            // logical bit sitting there costs the later sweeps nothing.  When this sweep finishes the work of the bit
            // at such a position while another bit of the tile still has ops that need it inside a tile (a shared
            // target, a carry, ...), the two swap positions in this sweep's store phase (in place, no extra pass):
            // bv_n15 4 -> 3 sweeps.  Controlled phases do not count: they need only one of their bits in a tile.
            std::vector<std::pair<int, int>> moves;
            if (opt.hot_low && lowb > 0 && !plan.has_srn)
            {
                std::vector<char> taken(ops.size(), 0);
                for (int i : best.picked) taken[i] = 1;
                std::vector<long> pend_w(N, 0);
                for (size_t i = first_pending; i < ops.size(); i++)
                    if (!ops[i].done && !taken[i] && !ops[i].cp)
                        for (int b = 0; b < ops[i].nb; b++) pend_w[ops[i].bit[b]] += ops[i].weight;
                std::vector<int> hot;
                for (int l : best.tile_logical)
                    if (phys[l] >= lowb && pend_w[l] > 0) hot.push_back(l);
                std::stable_sort(hot.begin(), hot.end(), [&](int a, int b) { return pend_w[a] > pend_w[b]; });
                for (int l = 0; l < N; l++) logical_at[phys[l]] = l;
                size_t h = 0;
                for (int p = 0; p < lowb && h < hot.size(); p++)
                {
                    const int dead = logical_at[p];
                    if (pend_w[dead] > 0) continue;
                    moves.push_back({dead, phys[hot[h]]});
                    moves.push_back({hot[h], p});
                    h++;
                }
            }
            emit_sweep(best.tile_logical, best.picked, moves, true);
            continue;
        }
        // Stuck: every pending op needs a rank bit.  Qubit remap: pick the g local logical bits with the
        // least pending work, move them to physical [M-g, M) (permuting sweep), then exchange.
        if (g == 0) throw std::logic_error("scheduler stuck without rank bits");
        std::vector<long> pending_w(N, 0);
        for (size_t i = first_pending; i < ops.size(); i++)
            if (!ops[i].done)
                for (int b = 0; b < ops[i].nb; b++) pending_w[ops[i].bit[b]] += ops[i].weight;
        // keep the partners of frontier ops that wait on a rank bit local, or the remap cannot unblock them
        {
            std::vector<char> seen(N, 0);
            for (size_t i = first_pending; i < ops.size(); i++)
            {
                if (ops[i].done) continue;
                bool frontier = true, on_rank = false;
                int n_rank = 0;
                for (int b = 0; b < ops[i].nb; b++)
                {
                    if (seen[ops[i].bit[b]]) frontier = false;
                    if (phys[ops[i].bit[b]] >= M) n_rank++;
                }
                on_rank = ops[i].cp ? n_rank == 2 : n_rank > 0; // a controlled phase only needs one local bit
                if (frontier && on_rank)
                    for (int b = 0; b < ops[i].nb; b++) pending_w[ops[i].bit[b]] += (long)1 << 40;
                for (int b = 0; b < ops[i].nb; b++) seen[ops[i].bit[b]] = 1;
            }
        }
        std::vector<int> locals;
        for (int l = 0; l < N; l++)
            if (phys[l] < M) locals.push_back(l);
        std::stable_sort(locals.begin(), locals.end(), [&](int a, int b) {
            if (pending_w[a] != pending_w[b]) return pending_w[a] < pending_w[b];
            return phys[a] > phys[b];
        });
        std::vector<int> chosen(locals.begin(), locals.begin() + g);
        // positions: chosen not already in [M-g, M) pair up with top-local bits not chosen
        for (int l = 0; l < N; l++) logical_at[phys[l]] = l;
        std::vector<int> a_only, b_only;
        std::vector<char> is_chosen(N, 0);
        for (int l : chosen) is_chosen[l] = 1;
        for (int l : chosen)
            if (phys[l] < M - g) a_only.push_back(l);
        for (int p = M - g; p < M; p++)
            if (!is_chosen[logical_at[p]]) b_only.push_back(logical_at[p]);
        if (!a_only.empty())
        {
            std::vector<int> tile;
            for (int l = 0; l < N; l++)
                if (phys[l] < lowb) tile.push_back(l);
            std::vector<std::pair<int, int>> moves;
            for (size_t i = 0; i < a_only.size(); i++)
            {
                for (int l : {a_only[i], b_only[i]})
                    if (std::find(tile.begin(), tile.end(), l) == tile.end()) tile.push_back(l);
                moves.push_back({a_only[i], phys[b_only[i]]});
                moves.push_back({b_only[i], phys[a_only[i]]});
            }
            if ((int)tile.size() > kmax) throw std::logic_error("remap permutation does not fit one tile");
            emit_sweep(tile, {}, moves);
            if (opt.spread_peers) plan.steps.back().sweep.spread_top = g;
        }
        Step ex;
        ex.kind = 1;
        plan.steps.push_back(ex);
        plan.n_exchanges++;
        for (int l = 0; l < N; l++)
        {
            if (phys[l] >= M) phys[l] -= g;
            else if (phys[l] >= M - g) phys[l] += g;
        }
    }
    // Every sim() of the reference ends in its transposed frame: the result it reports is S' = conj(X'^T) where X' is
    // what this plan computes (all L parts and all R parts applied without transposing).  For a Hermitian state the
    // two coincide; for the non-Hermitian states SRN (or an arbitrary dmb_set_dm input) produces they do not.  The
    // transpose costs nothing here: swap the row/column halves of the layout and flip the conjugation flag.
    // (When the state is known to be Hermitian and the circuit has no SRN the frame change is the identity and is
    // skipped, so that repeated runs of one circuit keep one plan / one CUDA graph.)
    plan.conj_start = plan.conj_end = conj_state;
    if (plan.has_srn || non_hermitian)
    {
        for (int l = 0; l < n; l++) std::swap(phys[l], phys[l + n]);
        plan.conj_end = !conj_state;
    }
    plan.end_layout = phys;
    return plan;
}

// ------------------------------------------------------------------------------------------------
std::string plan_to_json(const Plan& p, const std::vector<std::string>* extra)
{
    std::ostringstream o;
    char buf[64];
    auto num = [&](double v) { snprintf(buf, sizeof(buf), "%.17g", v); return std::string(buf); };
    auto ivec = [&](const std::vector<int>& v) {
        std::string s = "[";
        for (size_t i = 0; i < v.size(); i++) s += (i ? "," : "") + std::to_string(v[i]);
        return s + "]";
    };
    o << "{\"n\":" << p.n << ",\"g\":" << p.g << ",\"n_gates\":" << p.n_gates << ",\"n_primitives\":" << p.n_primitives
      << ",\"n_blocks\":" << p.n_blocks << ",\"n_sweeps\":" << p.n_sweeps << ",\"n_exchanges\":" << p.n_exchanges
      << ",\"has_srn\":" << (p.has_srn ? "true" : "false") << ",\"conj_start\":" << (p.conj_start ? "true" : "false")
      << ",\"conj_end\":" << (p.conj_end ? "true" : "false") << ",\"start_layout\":" << ivec(p.start_layout)
      << ",\"end_layout\":" << ivec(p.end_layout) << ",\"steps\":[";
    for (size_t s = 0; s < p.steps.size(); s++)
    {
        const Step& st = p.steps[s];
        if (s) o << ",";
        if (st.kind == 1) { o << "{\"kind\":\"exchange\"}"; continue; }
        const Sweep& sw = st.sweep;
        o << "{\"kind\":\"sweep\",\"k\":" << sw.k << ",\"in_pos\":" << ivec(sw.in_pos) << ",\"out_pos\":" << ivec(sw.out_pos)
          << ",\"out_of_place\":" << (sw.out_of_place ? "true" : "false") << ",\"swz\":" << sw.swz_mode << ",\"weight\":" << sw.weight;
        if (extra && s < extra->size() && !(*extra)[s].empty()) o << ",\"dev\":" << (*extra)[s];
        o << ",\"ops\":[";
        for (size_t i = 0; i < sw.ops.size(); i++)
        {
            const TileOp& t = sw.ops[i];
            if (i) o << ",";
            o << "{\"cls\":" << t.cls << ",\"nb\":" << t.nb << ",\"j0\":" << t.j0 << ",\"j1\":" << t.j1 << ",\"p1\":" << t.p1
              << ",\"m\":[";
            const int cnt = t.nb == 1 ? 4 : 16;
            for (int e = 0; e < cnt; e++) o << (e ? "," : "") << num(t.m[e].real()) << "," << num(t.m[e].imag());
            o << "]}";
        }
        o << "]}";
    }
    o << "]}";
    return o.str();
}
} // namespace dmb
