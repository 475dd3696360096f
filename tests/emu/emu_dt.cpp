// tests/emu/emu_dt.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the device headers (g_tessla_b200/csrc/dt_core.cuh, gt_predicates.cuh) driven
// sequentially in the kernel's order (sort -> bottom-up levels -> enumeration), so the merge
// logic and the exact predicates can be checked against the oracle without a GPU.
// Never linked into libgtessla_b200.so; the product has no CPU path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#define GT_STATS 1
#include "../../g_tessla_b200/csrc/dt_core.cuh"
#include "../../g_tessla_b200/csrc/lpf_core.cuh"
namespace gt { Stats g_stats; long g_exact_calls[8]; }

extern "C" {

// per-level work counters of the last emu_triangulate call: [level][0..5] = nodes, orient2d, incircle,
// merge steps, cuts, longest node (steps)
static long g_level[32][6];
void emu_level_stats(long* out) { memcpy(out, g_level, sizeof g_level); }
// exact-stage counters since the last call: orient2d with 1/2/4 words, incircle with 1/2/4 words, total orient2d, total incircle
void emu_exact_stats(long* out) {
    memcpy(out, gt::g_exact_calls, sizeof gt::g_exact_calls);
    out[6] = gt::g_stats.o2d; out[7] = gt::g_stats.icc;
    memset(gt::g_exact_calls, 0, sizeof gt::g_exact_calls); gt::g_stats.o2d = gt::g_stats.icc = 0;
}

// points [n][2] -> triangles in emission order (indices into the input). Returns ntri or -(err bits).
int emu_triangulate(const double* points, int n, int* tri_out) {
    using namespace gt;
    if (n < 2) return -1000;
    memset(g_level, 0, sizeof g_level);
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::sort(perm.begin(), perm.end(), [&](int u, int v) {
        double ux = points[2 * u], uy = points[2 * u + 1], vx = points[2 * v], vy = points[2 * v + 1];
        return ux < vx || (ux == vx && uy < vy);
    });
    // runs of |dx| < 1e-12 that are not exact ties are ordered by y (reference compareVerts), as in the kernel
    {
        auto X = [&](int i) { return points[2 * perm[i]]; };
        auto tie = [&](int a, int b) { double d = X(a) - X(b); return d < 1e-12 && d > -1e-12; };
        for (int i = 0; i + 1 < n;) {
            int e = i + 1; bool inexact = false;
            while (e < n && tie(e, e - 1)) { inexact |= (X(e) != X(e - 1)); ++e; }
            if (inexact) {
                if (!tie(e - 1, i)) return -2000;
                std::stable_sort(perm.begin() + i, perm.begin() + e, [&](int u, int v) { return points[2 * u + 1] < points[2 * v + 1]; });
            }
            i = e;
        }
    }
    std::vector<Pt> pt(n);
    for (int i = 0; i < n; ++i) { pt[i].x = points[2 * perm[i]]; pt[i].y = points[2 * perm[i] + 1]; }
    const int cap_e = 3 * n + 64;
    std::vector<u16> head(n, GT_NIL), dst(2 * cap_e), nxt(2 * cap_e), prv(2 * cap_e);
    uint32_t bump = 0, stack = GT_NIL, err = 0;
    Frame f;
    f.pt = pt.data(); f.head = head.data(); f.dst = dst.data(); f.nxt = nxt.data(); f.prv = prv.data();
    f.bump = &bump; f.stack = &stack; f.err = &err; f.n = n; f.cap_e = cap_e;
    const int D = tree_depth(n);
    for (int d = D; d >= 0; --d) {
        // one "lane" per node, flushed at the end of the level like the kernel does
        std::vector<Lane> lanes(1 << d);
        memset(g_level[d], 0, sizeof g_level[d]);
        for (int k = 0; k < (1 << d); ++k) {
            lanes[k].free_head = GT_NIL;
            Stats s0 = g_stats;
            process_node(f, lanes[k], d, k);
            int ia, ib;
            if (tree_node(n, d, k, &ia, &ib) && ib - ia >= 1) g_level[d][0]++;
            g_level[d][1] += g_stats.o2d - s0.o2d; g_level[d][2] += g_stats.icc - s0.icc;
            g_level[d][3] += g_stats.steps - s0.steps; g_level[d][4] += g_stats.cuts - s0.cuts;
            if (g_stats.steps - s0.steps > g_level[d][5]) g_level[d][5] = g_stats.steps - s0.steps;
        }
        for (int k = 0; k < (1 << d); ++k) flush_lane(f, lanes[k]);
        if (err) return -(int)err;
    }
    int nt = 0;
    for (int i = 0; i < n; ++i)
        vertex_triangles(f, i, [&](int n1, int n2) {
            tri_out[3 * nt] = perm[i]; tri_out[3 * nt + 1] = perm[n1]; tri_out[3 * nt + 2] = perm[n2];
            ++nt;
        });
    return nt;
}

// Same input/output as emu_triangulate, through the per-lane state machine (lpf_core.cuh), one lane.
// iters_out (nullable) receives the number of iterations the lane needed.
int emu_triangulate_lpf(const double* points, int n, int* tri_out, long* iters_out) {
    using namespace gt;
    if (n < 2) return -1000;
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    std::sort(perm.begin(), perm.end(), [&](int u, int v) {
        double ux = points[2 * u], uy = points[2 * u + 1], vx = points[2 * v], vy = points[2 * v + 1];
        return ux < vx || (ux == vx && uy < vy);
    });
    {
        auto X = [&](int i) { return points[2 * perm[i]]; };
        auto tie = [&](int a, int b) { double d = X(a) - X(b); return d < 1e-12 && d > -1e-12; };
        for (int i = 0; i + 1 < n;) {
            int e = i + 1; bool inexact = false;
            while (e < n && tie(e, e - 1)) { inexact |= (X(e) != X(e - 1)); ++e; }
            if (inexact) {
                if (!tie(e - 1, i)) return -2000;
                std::stable_sort(perm.begin() + i, perm.begin() + e, [&](int u, int v) { return points[2 * u + 1] < points[2 * v + 1]; });
            }
            i = e;
        }
    }
    std::vector<LPt> pt(n);
    for (int i = 0; i < n; ++i) { pt[i].x = points[2 * perm[i]]; pt[i].y = points[2 * perm[i] + 1]; pt[i].z = 0.0; pt[i].head = (u16)GT_NIL; pt[i].perm = (u16)perm[i]; pt[i].pad = 0; }
    const int cap_e = 3 * n + 64;
    std::vector<LRec> rec(2 * cap_e);
    LFrame f; f.pt = pt.data(); f.rec = rec.data(); f.n = n; f.cap_e = cap_e;
    LState s; memset(&s, 0, sizeof s); lp_begin(f, s);
    long iters = 0;
    bool more = s.phase != LP_DONE;
    uint32_t w[LS_WORDS];
    ls_pack(s, [&](int i, uint32_t v) { w[i] = v; });
    while (more) {
        // the GPU scheduler keeps only the packed form between iterations and rewrites only the words the
        // phase may modify (LsDirty): go through exactly that here
        LState t; memset(&t, 0xee, sizeof t);
        ls_unpack(t, [&](int i) { return w[i]; });
        auto put = [&](int i, uint32_t v) { w[i] = v; };
        switch (t.phase) {
#define EMU_PH(P) case P: lpf_step_t<P>(f, t); ls_writeback<P>(t, put); break;
        EMU_PH(LP_LEAF) EMU_PH(LP_TINIT) EMU_PH(LP_TAN) EMU_PH(LP_INS) EMU_PH(LP_SCAN) EMU_PH(LP_V) EMU_PH(LP_C)
#undef EMU_PH
        default: return -4000;
        }
        s = t;
        more = t.phase != LP_DONE;
        if (++iters > 400L * n + 10000) return -3000;
    }
    if (iters_out) *iters_out = iters;
    if (s.err) return -(int)s.err;
    int nt = 0;
    for (int i = 0; i < n; ++i)
        lf_vertex_triangles(f, i, [&](int n1, int n2) {
            tri_out[3 * nt] = perm[i]; tri_out[3 * nt + 1] = perm[n1]; tri_out[3 * nt + 2] = perm[n2];
            ++nt;
        });
    return nt;
}

void emu_orient2d_signs(const double* p, int n, int* out) {
    for (int i = 0; i < n; ++i) { const double* c = p + 6 * (size_t)i;
        out[i] = gt::orient2d_sign(c[0], c[1], c[2], c[3], c[4], c[5]); }
}
void emu_incircle_signs(const double* p, int n, int* out) {
    for (int i = 0; i < n; ++i) { const double* c = p + 8 * (size_t)i;
        out[i] = gt::incircle_sign(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]); }
}
// exact stage only (bypasses the FP64 filter)
void emu_orient2d_exact(const double* p, int n, int* out) {
    for (int i = 0; i < n; ++i) { const double* c = p + 6 * (size_t)i;
        out[i] = gt::orient2d_exact(c[0], c[1], c[2], c[3], c[4], c[5]); }
}
void emu_incircle_exact(const double* p, int n, int* out) {
    for (int i = 0; i < n; ++i) { const double* c = p + 8 * (size_t)i;
        out[i] = gt::incircle_exact(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]); }
}
}
