// tests/emu/emu_dt.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the device headers (g_tessla_b200/csrc/dt_core.cuh, gt_predicates.cuh) driven
// sequentially in the kernel's order (sort -> bottom-up levels -> enumeration), so the merge
// logic and the exact predicates can be checked against the oracle without a GPU.
// Never linked into libgtessla_b200.so; the product has no CPU path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#define GT_STATS 1
#include "../../g_tessla_b200/csrc/dt_core.cuh"
#include "../../g_tessla_b200/csrc/lpf_core.cuh"
#include "../../g_tessla_b200/csrc/star_core.cuh"
namespace gt { Stats g_stats; long g_exact_calls[8]; StarStats g_star; }

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

// Same input/output as emu_triangulate, through the walker state machine (lpf_core.cuh) the way the GPU scheduler drives
// it: the frame is split into 2^level subtrees (level = min(want_level, what n allows)), every subtree gets a walker,
// walkers are stepped ONE ITERATION AT A TIME IN A PSEUDO-RANDOM INTERLEAVING (seed), they share the frame's edge pool
// and arrive at their parents through the "ready" bits. The result must not depend on the interleaving.
// stats_out (nullable): [0] iterations in total, [1] edges taken from the pool, [2] longest walker (iterations),
// [3] iterations of the walker that finished the root.
int emu_triangulate_lpf2(const double* points, int n, int want_level, unsigned seed, int* tri_out, long* stats_out) {
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
    LMeta meta = {0, 0, 0};
    LFrame f; f.pt = pt.data(); f.rec = rec.data(); f.m = &meta; f.n = n; f.cap_e = cap_e;
    const int level = lp_split_level(n, want_level);
    struct Walker { uint32_t w[LS_WORDS]; long iters; };
    std::vector<Walker> live(1 << level);
    for (int k = 0; k < (1 << level); ++k) {
        LState s; memset(&s, 0, sizeof s); lp_begin(f, s, level, k);
        ls_pack(s, [&](int i, uint32_t v) { live[k].w[i] = v; });
        live[k].iters = 0;
    }
    long iters = 0, longest = 0, root_iters = 0;
    unsigned long long rng = seed * 6364136223846793005ull + 1442695040888963407ull;
    bool root_done = false;
    while (!live.empty()) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        const size_t pick = seed ? (size_t)((rng >> 33) % live.size()) : 0;     // seed 0: one walker at a time, to its end
        Walker& wk = live[pick];
        // the GPU scheduler keeps only the packed form between iterations and rewrites only the words the
        // phase may modify (LsDirty): go through exactly that here
        LState t; memset(&t, 0xee, sizeof t);
        ls_unpack(t, [&](int i) { return wk.w[i]; });
        auto put = [&](int i, uint32_t v) { wk.w[i] = v; };
        const int ph = t.phase;
        switch (ph) {
#define EMU_PH(P) case P: lpf_step_t<P>(f, t); ls_writeback<P>(t, put); break;
        EMU_PH(LP_LEAF) EMU_PH(LP_TINIT) EMU_PH(LP_TAN) EMU_PH(LP_INS) EMU_PH(LP_SCAN) EMU_PH(LP_V) EMU_PH(LP_C)
#undef EMU_PH
        default: return -4000;
        }
        ++wk.iters;
        if (t.err) {                                       // as gta_lpf.cuh::lpf_wave_error does, from the packed state
            ls_unpack(t, [&](int i) { return wk.w[i]; });
            if (t.phase >= LP_DONE || !(t.err & (GT_ERR_POOL | GT_ERR_RING))) meta.err |= t.err;
            else { lp_abort(f, t); ls_pack(t, put); }
        }
        if (++iters > 400L * n + 10000) return -3000;
        if (t.phase == LP_DONE || t.phase == LP_ROOT) {
            if (wk.iters > longest) longest = wk.iters;
            if (t.phase == LP_ROOT) { if (root_done) return -5000; root_done = true; root_iters = wk.iters; }
            live[pick] = live.back(); live.pop_back();
        }
    }
    if (!root_done) return -6000;
    if (stats_out) { stats_out[0] = iters; stats_out[1] = (long)meta.bump; stats_out[2] = longest; stats_out[3] = root_iters; }
    if (meta.err) return -(int)meta.err;
    int nt = 0;
    for (int i = 0; i < n; ++i)
        lf_vertex_triangles(f, i, [&](int n1, int n2) {
            tri_out[3 * nt] = perm[i]; tri_out[3 * nt + 1] = perm[n1]; tri_out[3 * nt + 2] = perm[n2];
            ++nt;
        });
    return nt;
}
// one walker for the whole frame (the tree in post-order)
int emu_triangulate_lpf(const double* points, int n, int* tri_out, long* iters_out) {
    long st[4] = {0, 0, 0, 0};
    const int r = emu_triangulate_lpf2(points, n, 0, 0u, tri_out, st);
    if (iters_out) *iters_out = st[0];
    return r;
}

// The star path (star_core.cuh) the way the GPU kernel drives it: bin the points into the grid (cell order, input order
// inside a cell), one star per point, owned triangles out (indices into the input, any order). Returns ntri, or
// -(STAR_* bits) when the frame is not certified (ties, duplicates), or -64 when the triangle count is not 2(n-1)-h.
// stats_out (nullable): [0] candidates visited, [1] left of the edge, [2] incircle filters, [3] exact-stage calls,
// [4] window expansions, [5] walk steps, [6] points, [7] hull certificates, [8] hull vertices, [9] gx, [10] gy
int emu_star(const double* points, int n, int w0, double dens, int full, int ccap, int* tri_out, long* stats_out) {
    using namespace gt;
    if (n < 2) return -1000;
    memset(&g_star, 0, sizeof g_star);
    double xmin = points[0], xmax = points[0], ymin = points[1], ymax = points[1];
    for (int i = 1; i < n; ++i) {
        xmin = std::min(xmin, points[2 * i]); xmax = std::max(xmax, points[2 * i]);
        ymin = std::min(ymin, points[2 * i + 1]); ymax = std::max(ymax, points[2 * i + 1]);
    }
    const StarDims d = star_dims(n, xmin, xmax, ymin, ymax, std::max(n, 1), dens);
    const int nc = d.gx * d.gy;
    std::vector<int> cell(n), cstart(nc + 1, 0), id(n);
    for (int i = 0; i < n; ++i) {
        cell[i] = star_cell(points[2 * i + 1], ymin, d.sy, d.gy) * d.gx + star_cell(points[2 * i], xmin, d.sx, d.gx);
        cstart[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; ++c) cstart[c + 1] += cstart[c];
    std::vector<int> cur(cstart.begin(), cstart.end() - 1);
    std::vector<SPt> pt(n);
    for (int i = 0; i < n; ++i) { const int pos = cur[cell[i]]++; pt[pos].x = points[2 * i]; pt[pos].y = points[2 * i + 1]; id[pos] = i; }
    std::vector<SPtF> ptf(n);
    double M = 0.0;
    for (int i = 0; i < n; ++i) { ptf[i].x = (float)pt[i].x; ptf[i].y = (float)pt[i].y; M = std::max(M, std::max(fabs(pt[i].x), fabs(pt[i].y))); }
    StarGrid<int> g;
    g.ptf = ptf.data(); g.kf = (float)(2.5 * 5.9604644775390625e-08 * M) * 1.0001f;
    g.pt = pt.data(); g.cstart = cstart.data(); g.n = n; g.gx = d.gx; g.gy = d.gy;
    g.x0 = xmin; g.y0 = ymin; g.sx = d.sx; g.sy = d.sy; g.xmin = xmin; g.xmax = xmax; g.ymin = ymin; g.ymax = ymax;
    g.cx0 = g.cy0 = 0; g.ggx = d.gx; g.ggy = d.gy; g.open = 0; g.rowx = nullptr;
    int nt = 0, st = 0, nh = 0;
    for (int ip = 0; ip < n; ++ip) {
        bool hull = false;
        // (short lists: the refill of the left list and the overflow of the candidate list are exercised)
        // (passed by value to the passes, like the device's pair of shared-memory pointers)
        int lv[8], lc[64];
        struct { int *v, *c; void put(int k, int pos) { v[k] = pos; } int get(int k) const { return v[k]; }
                 void cput(int k, int pos) { c[k] = pos; } int cget(int k) const { return c[k]; } } lst;
        lst.v = lv; lst.c = lc;
        auto emit = [&](int n1, int n2) {
            tri_out[3 * nt] = id[ip]; tri_out[3 * nt + 1] = id[n1]; tri_out[3 * nt + 2] = id[n2];
            ++nt;
        };
        const long rows0 = g_star.rows, gc0 = g_star.gridcand, ex0 = g_star.expand;
        if (ccap <= 12) {
            if (full) st |= star_point<8, 12, true>(g, ip, true, w0, lst, &hull, emit);
            else st |= star_point<8, 12, false>(g, ip, true, w0, lst, &hull, emit);
        } else {
            if (full) st |= star_point<8, 64, true>(g, ip, true, w0, lst, &hull, emit);
            else st |= star_point<8, 64, false>(g, ip, true, w0, lst, &hull, emit);
        }
        if (getenv("STAR_DEBUG") && (g_star.rows - rows0 > 20000 || g_star.gridcand - gc0 > 20000))
            fprintf(stderr, "heavy point pos %d (%.4f, %.4f) cell (%d, %d): region rows %ld, candidates %ld, scans %ld, hull %d\n", ip, pt[ip].x, pt[ip].y,
                    star_cell(pt[ip].x, xmin, d.sx, d.gx), star_cell(pt[ip].y, ymin, d.sy, d.gy), g_star.rows - rows0, g_star.gridcand - gc0, g_star.expand - ex0, (int)hull);
        nh += hull;
    }
    if (stats_out) {
        stats_out[0] = g_star.cand; stats_out[1] = g_star.left; stats_out[2] = g_star.icc; stats_out[3] = g_star.exact;
        stats_out[4] = g_star.expand; stats_out[5] = g_star.steps; stats_out[6] = g_star.points; stats_out[7] = g_star.hullcert;
        stats_out[8] = nh; stats_out[9] = d.gx; stats_out[10] = d.gy; stats_out[11] = g_star.gridcand; stats_out[12] = g_star.rows;
    }
    if (st) return -st;
    if (!full) {
        // arcs only: the hull size is not known from the walks. When the four corners of the bounding box are points, the hull
        // is that rectangle and its vertices are the points on it (the -corr case); otherwise the count is not checked here.
        int corners = 0; nh = 0;
        for (int i = 0; i < n; ++i) {
            const double x = points[2 * i], y = points[2 * i + 1];
            const bool bx = x == xmin || x == xmax, by = y == ymin || y == ymax;
            nh += bx || by; corners += bx && by;
        }
        if (corners < 4) return nt;
    }
    if (nt != 2 * (n - 1) - nh) return -64;
    return nt;
}

// The star path over TILES of the grid, the way ls_tile_kernel (gta_large_star.cuh) drives it: every tile of tw x tw cells is
// staged with a halo of `halo` cells (StarGrid GL = 2), its own points build their whole stars from the staged cells, and a
// point whose search would leave the staged cells on an open side is deferred and done over the whole grid afterwards.
// Returns ntri (or -(STAR_* bits), -64 count rule); stats_out[0] = deferred points, [1] = tiles.
int emu_star_tiles(const double* points, int n, int tw, int halo, double dens, int* tri_out, long* stats_out) {
    using namespace gt;
    if (n < 3) return -1000;
    memset(&g_star, 0, sizeof g_star);
    double xmin = points[0], xmax = points[0], ymin = points[1], ymax = points[1];
    for (int i = 1; i < n; ++i) {
        xmin = std::min(xmin, points[2 * i]); xmax = std::max(xmax, points[2 * i]);
        ymin = std::min(ymin, points[2 * i + 1]); ymax = std::max(ymax, points[2 * i + 1]);
    }
    const StarDims d = star_dims(n, xmin, xmax, ymin, ymax, std::max(n, 1), dens);
    const int nc = d.gx * d.gy, gx = d.gx, gy = d.gy;
    std::vector<int> cell(n), cstart(nc + 1, 0), id(n);
    for (int i = 0; i < n; ++i) {
        cell[i] = star_cell(points[2 * i + 1], ymin, d.sy, d.gy) * d.gx + star_cell(points[2 * i], xmin, d.sx, d.gx);
        cstart[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; ++c) cstart[c + 1] += cstart[c];
    std::vector<int> cur(cstart.begin(), cstart.end() - 1);
    std::vector<SPt> pt(n);
    for (int i = 0; i < n; ++i) { const int pos = cur[cell[i]]++; pt[pos].x = points[2 * i]; pt[pos].y = points[2 * i + 1]; id[pos] = i; }
    std::vector<SPtF> ptf(n);
    double M = 0.0;
    for (int i = 0; i < n; ++i) { ptf[i].x = (float)pt[i].x; ptf[i].y = (float)pt[i].y; M = std::max(M, std::max(fabs(pt[i].x), fabs(pt[i].y))); }
    const float kf = (float)(2.5 * 5.9604644775390625e-08 * M) * 1.0001f;
    int nt = 0, st = 0, nh = 0;
    int lv[8], lc[64];
    struct { int *v, *c; void put(int k, int pos) { v[k] = pos; } int get(int k) const { return v[k]; }
             void cput(int k, int pos) { c[k] = pos; } int cget(int k) const { return c[k]; } } lst;
    lst.v = lv; lst.c = lc;
    std::vector<int> deferred;
    long ntiles = 0;
    for (int ty = 0; ty * tw < gy; ++ty) for (int tx = 0; tx * tw < gx; ++tx) {
        ++ntiles;
        const int cc0 = tx * tw, cc1 = std::min(gx, cc0 + tw) - 1, cr0 = ty * tw, cr1 = std::min(gy, cr0 + tw) - 1;
        const int C0 = std::max(0, cc0 - halo), C1 = std::min(gx - 1, cc1 + halo), R0 = std::max(0, cr0 - halo), R1 = std::min(gy - 1, cr1 + halo);
        const int W = C1 - C0 + 1, Hh = R1 - R0 + 1;
        std::vector<int> rowg0(Hh), rowbase(Hh + 1, 0), lcs(W * Hh + 1);
        for (int r = 0; r < Hh; ++r) { rowg0[r] = cstart[(R0 + r) * gx + C0]; rowbase[r + 1] = rowbase[r] + cstart[(R0 + r) * gx + C1 + 1] - rowg0[r]; }
        const int nloc = rowbase[Hh];
        for (int r = 0; r < Hh; ++r) for (int c = 0; c < W; ++c) lcs[r * W + c] = cstart[(R0 + r) * gx + C0 + c] - rowg0[r] + rowbase[r];
        lcs[W * Hh] = nloc;
        std::vector<SPt> lpt(std::max(nloc, 1)); std::vector<SPtF> lptf(std::max(nloc, 1));
        for (int r = 0; r < Hh; ++r) for (int j = 0; j < rowbase[r + 1] - rowbase[r]; ++j) { lpt[rowbase[r] + j] = pt[rowg0[r] + j]; lptf[rowbase[r] + j] = ptf[rowg0[r] + j]; }
        StarGrid<int, 2> g;
        g.pt = lpt.data(); g.ptf = lptf.data(); g.cstart = lcs.data(); g.kf = kf; g.n = nloc; g.gx = W; g.gy = Hh;
        g.x0 = xmin; g.y0 = ymin; g.sx = d.sx; g.sy = d.sy; g.xmin = xmin; g.xmax = xmax; g.ymin = ymin; g.ymax = ymax;
        g.cx0 = C0; g.cy0 = R0; g.ggx = gx; g.ggy = gy;
        g.open = (C0 > 0 ? 1 : 0) | (C1 < gx - 1 ? 2 : 0) | (R0 > 0 ? 4 : 0) | (R1 < gy - 1 ? 8 : 0);
        auto gpos = [&](int lp) { const int r = star_row(g, lpt[lp].y); return rowg0[r] + lp - rowbase[r]; };
        for (int r = cr0 - R0; r <= cr1 - R0; ++r) for (int lp = lcs[r * W + (cc0 - C0)]; lp < lcs[r * W + (cc1 - C0) + 1]; ++lp) {
            const int gip = gpos(lp);
            if (nloc < 3) { deferred.push_back(gip); continue; }
            bool hull = false;
            const int nt0 = nt;
            const int s1 = star_point<8, 12, true>(g, lp, true, 2, lst, &hull, [&](int n1, int n2) {
                tri_out[3 * nt] = id[gip]; tri_out[3 * nt + 1] = id[gpos(n1)]; tri_out[3 * nt + 2] = id[gpos(n2)];
                ++nt;
            });
            if (s1 & STAR_DEFER) { nt = nt0; deferred.push_back(gip); }
            else { st |= s1; nh += hull; }
        }
    }
    StarGrid<int, 1> g;                                       // (GL = 1, as on the device: the region walk that queues four rows)
    g.ptf = ptf.data(); g.kf = kf; g.pt = pt.data(); g.cstart = cstart.data(); g.n = n; g.gx = gx; g.gy = gy;
    g.x0 = xmin; g.y0 = ymin; g.sx = d.sx; g.sy = d.sy; g.xmin = xmin; g.xmax = xmax; g.ymin = ymin; g.ymax = ymax;
    g.cx0 = g.cy0 = 0; g.ggx = gx; g.ggy = gy; g.open = 0;
    std::vector<double> rowx(2 * (size_t)gy);
    for (int r = 0; r < gy; ++r) {
        double lo = INFINITY, hi = -INFINITY;
        for (int i = cstart[r * gx]; i < cstart[(r + 1) * gx]; ++i) { lo = std::min(lo, pt[i].x); hi = std::max(hi, pt[i].x); }
        rowx[2 * r] = lo; rowx[2 * r + 1] = hi;
    }
    g.rowx = getenv("STAR_NOROWX") ? nullptr : rowx.data();
    const long rows_t = g_star.rows, gc_t = g_star.gridcand, ex_t = g_star.expand, exact_t = g_star.exact, icc_t = g_star.icc;
    for (int ip : deferred) {
        bool hull = false;
        st |= star_point<8, 12, true>(g, ip, true, 2, lst, &hull, [&](int n1, int n2) {
            tri_out[3 * nt] = id[ip]; tri_out[3 * nt + 1] = id[n1]; tri_out[3 * nt + 2] = id[n2];
            ++nt;
        });
        nh += hull;
    }
    if (stats_out) { stats_out[0] = (long)deferred.size(); stats_out[1] = ntiles; stats_out[2] = nh; stats_out[3] = g_star.rows; }
    if (getenv("STAR_DEBUG")) fprintf(stderr, "deferred pass: rows %ld, rejected by the row extents %ld, grid candidates %ld, scans %ld, exact %ld, icc %ld\n", g_star.rows - rows_t, g_star.screened, g_star.gridcand - gc_t, g_star.expand - ex_t, g_star.exact - exact_t, g_star.icc - icc_t);
    if (st) return -st;
    if (nt != 2 * (n - 1) - nh) return -64;
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
