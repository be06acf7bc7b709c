// dm-sim_b200/csrc/encode.cpp -- turns a planned Sweep into what sweep_kernel consumes: device ops with
// pre-swizzled index tables, warp groups, and the load/store address tables of the kernel parameter block.
// Pure host code (system compiler).
#include "encode.hpp"

#include <algorithm>
#include <cstring>
#include <cstdio>
#include <sstream>

namespace dmb
{
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

// class-specific payload (which matrix entries the device reads, skip masks)
void encode_payload(const TileOp& t, DevOp& d)
{
    const cplx one(1.0, 0.0);
    switch (t.cls)
    {
    case CLS_DENSE2:
        for (int i = 0; i < 16; i++) put(d, i, t.m[i]);
        break;
    case CLS_DENSE1:
        for (int i = 0; i < 4; i++) put(d, i, t.m[i]);
        break;
    case CLS_DIAG2:
    {
        int skip = 0;
        for (int r = 0; r < 4; r++)
        {
            put(d, r, t.m[r * 5]);
            if (t.m[r * 5] == one) skip |= 1 << r;
        }
        d.aux = skip << 8;
        break;
    }
    case CLS_DIAG1:
    {
        int skip = 0;
        for (int r = 0; r < 2; r++)
        {
            put(d, r, t.m[r * 3]);
            if (t.m[r * 3] == one) skip |= 1 << r;
        }
        d.aux = skip << 8;
        break;
    }
    case CLS_MONO2:
    {
        int src[4];
        classify(2, t.m, src);
        int aux = 0, skip = 0;
        bool unit = true;
        for (int r = 0; r < 4; r++)
        {
            const cplx ph = t.m[r * 4 + src[r]];
            put(d, r, ph);
            aux |= src[r] << (2 * r);
            if (ph != one) unit = false;
            if (src[r] == r && ph == one) skip |= 1 << r;
        }
        d.aux = aux | (skip << 8) | ((unit ? 1 : 0) << 12);
        break;
    }
    case CLS_MONO1:
    {
        put(d, 0, t.m[1]);
        put(d, 1, t.m[2]);
        const bool unit = (t.m[1] == one && t.m[2] == one);
        d.aux = (unit ? 1 : 0) << 12;
        break;
    }
    default:
        break;
    }
}
} // namespace

void encode_sweep(const Sweep& sw, EncodedSweep& out)
{
    out.ops.clear();
    out.groups.clear();
    const int k = sw.k;
    const int nwb = k >= kWarpBits + 2 ? kWarpBits : 0;
    const size_t n = sw.ops.size();
    size_t first = 0;
    while (first < n)
    {
        // grow the group while kWarpBits tile bits stay untouched
        unsigned used = 0;
        size_t end = first;
        while (end < n)
        {
            unsigned u = used | (1u << sw.ops[end].j0);
            if (sw.ops[end].nb == 2) u |= 1u << sw.ops[end].j1;
            if (k - __builtin_popcount(u) < nwb) break;
            used = u;
            end++;
        }
        DevGroup g;
        memset(&g, 0, sizeof(g));
        g.first = (int32_t)first;
        g.count = (int32_t)(end - first);
        g.n_warps = 1 << nwb;
        std::vector<int> wpos; // the highest untouched bits carry the warp index
        for (int p = k - 1; p >= 0 && (int)wpos.size() < nwb; p--)
            if (!((used >> p) & 1u)) wpos.push_back(p);
        std::sort(wpos.begin(), wpos.end());
        unsigned wmask = 0;
        for (int p : wpos) wmask |= 1u << p;
        for (int w = 0; w < 8; w++) g.wtab[w] = (uint16_t)swz_host(deposit((unsigned)w, wpos));
        out.groups.push_back(g);

        for (size_t i = first; i < end; i++)
        {
            const TileOp& t = sw.ops[i];
            DevOp d;
            memset(&d, 0, sizeof(d));
            d.cls = t.cls;
            encode_payload(t, d);
            unsigned opmask = 1u << t.j0;
            if (t.nb == 2) opmask |= 1u << t.j1;
            std::vector<int> freep;
            for (int p = 0; p < k; p++)
                if (!(((wmask | opmask) >> p) & 1u)) freep.push_back(p);
            const int nfree = (int)freep.size();
            const int nl = std::min(5, nfree);
            // lane bits 0..2 vary inside one LDS.128 phase: give them positions from three different swizzle
            // classes ({0,3},{1,4},{2,5}) whenever the op leaves one free, so the phase is conflict-free
            std::vector<int> lanep;
            std::vector<char> taken(k, 0);
            for (int cls = 0; cls < 3 && (int)lanep.size() < nl; cls++)
                for (int p : {cls, cls + 3})
                    if (p < k && !taken[p] && std::find(freep.begin(), freep.end(), p) != freep.end())
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
            d.n_iter = 1 << (int)iterp.size();
            d.n_active = 1 << nl;
            for (int l = 0; l < 32; l++) d.lane_tab[l] = (uint16_t)swz_host(deposit((unsigned)l & ((1u << nl) - 1u), lanep));
            for (int it = 0; it < 8; it++)
                d.iter_tab[it] = (uint16_t)swz_host(deposit((unsigned)it & ((unsigned)d.n_iter - 1u), iterp));
            if (t.nb == 2)
            {
                d.off[1] = (uint16_t)swz_host(1u << t.j1);
                d.off[2] = (uint16_t)swz_host(1u << t.j0);
                d.off[3] = d.off[1] ^ d.off[2];
            }
            else
                d.off[1] = (uint16_t)swz_host(1u << t.j0);
            out.ops.push_back(d);
        }
        first = end;
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
    for (int i = 0; i < k && i < 8; i++)
    {
        a.gin[i] = (unsigned char)sw.in_pos[i];
        a.gout[i] = (unsigned char)sw.out_pos[ord[i]];
        a.sout[i] = (unsigned char)ord[i];
    }
    for (int it = 0; it < 16; it++)
    {
        unsigned long long hi = 0, ho = 0;
        unsigned hs = 0;
        for (int i = 8; i < k; i++)
        {
            const unsigned long long bit = (it >> (i - 8)) & 1;
            hi |= bit << sw.in_pos[i];
            ho |= bit << sw.out_pos[ord[i]];
            hs |= (unsigned)bit << ord[i];
        }
        a.hin[it] = hi;
        a.hout[it] = ho;
        a.hs[it] = (unsigned short)swz_host(hs);
    }
    std::vector<char> used_in(M, 0), used_out(M, 0);
    for (int i = 0; i < k; i++) { used_in[sw.in_pos[i]] = 1; used_out[sw.out_pos[i]] = 1; }
    int ci = 0, co = 0;
    for (int p = 0; p < M; p++)
    {
        if (!used_in[p]) a.cin[ci++] = (unsigned char)p;
        if (!used_out[p]) a.cout[co++] = (unsigned char)p;
    }
}

std::string encoded_to_json(const EncodedSweep& e, const SweepArgs& a)
{
    std::ostringstream o;
    auto arr = [&](const char* name, auto* v, int n) {
        o << "\"" << name << "\":[";
        for (int i = 0; i < n; i++) o << (i ? "," : "") << (unsigned long long)v[i];
        o << "]";
    };
    o << "{\"k\":" << a.k << ",\"n_comp\":" << a.n_comp << ",";
    arr("hin", a.hin, 16); o << ",";
    arr("hout", a.hout, 16); o << ",";
    arr("hs", a.hs, 16); o << ",";
    arr("gin", a.gin, 8); o << ",";
    arr("gout", a.gout, 8); o << ",";
    arr("sout", a.sout, 8); o << ",";
    arr("cin", a.cin, a.n_comp); o << ",";
    arr("cout", a.cout, a.n_comp);
    o << ",\"groups\":[";
    for (size_t g = 0; g < e.groups.size(); g++)
    {
        const DevGroup& G = e.groups[g];
        o << (g ? "," : "") << "{\"first\":" << G.first << ",\"count\":" << G.count << ",\"n_warps\":" << G.n_warps << ",";
        arr("wtab", G.wtab, 8);
        o << "}";
    }
    o << "],\"ops\":[";
    char buf[40];
    for (size_t i = 0; i < e.ops.size(); i++)
    {
        const DevOp& d = e.ops[i];
        o << (i ? "," : "") << "{\"cls\":" << d.cls << ",\"aux\":" << d.aux << ",\"n_iter\":" << d.n_iter
          << ",\"n_active\":" << d.n_active << ",";
        arr("lane_tab", d.lane_tab, 32); o << ",";
        arr("iter_tab", d.iter_tab, 8); o << ",";
        arr("off", d.off, 4);
        o << ",\"m\":[";
        for (int j = 0; j < 32; j++)
        {
            snprintf(buf, sizeof(buf), "%.17g", d.m[j]);
            o << (j ? "," : "") << buf;
        }
        o << "]}";
    }
    o << "]}";
    return o.str();
}
} // namespace dmb
