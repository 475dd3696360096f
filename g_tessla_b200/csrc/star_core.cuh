// star_core.cuh -- the Delaunay star of ONE point, built from a uniform grid: no rings, no pointer chasing.
//
// The reference's divide and conquer (src/delaunay_tri.c:511-601) is a chain of data-dependent ring edits; the
// walker kernel (lpf_core.cuh) follows it decision for decision and is bound by the latency of that chain. But what
// the path has to deliver is the reference's RESULT (BASELINE.json: "the triangle index set must be bit-exact after
// canonical sorting"), and for a point set whose Delaunay triangulation is unique -- no four cocircular points on an
// empty circle -- every correct algorithm with exact predicates delivers that same set. So the fast path computes every
// point's star independently (one thread per point, the frame in shared memory) and CERTIFIES what it did:
//   * start edge: the exact nearest neighbour q0 of p (a Gabriel edge, hence Delaunay). Nearest by rounded squared
//     distance; a runner-up within the rounding error of the winner is "not decided" (tie).
//   * walk: given a Delaunay edge p->q, the triangle on its left is (p, q, r) with r the left point that no other
//     left point beats, s beats r <=> incircle(p, q, r, s) > 0 (exact sign: FP filter of the reference,
//     predicates.c:3238-3267, with the exact stage of gt_predicates.cuh behind it). Counter-clockwise from q0 until the
//     star closes; if it runs into the hull, clockwise from q0 (the same code on mirrored coordinates).
//   * completeness: candidates come from a window of grid cells around p. A triangle counts only when its whole
//     circumdisc lies inside the window (conservative FP bound, monotone cell map), a hull edge only when all four
//     corners of the point set's bounding box are on its right or on it (exact orient2d) -- otherwise the window grows
//     ring by ring (only the new ring is scanned) up to the whole grid, where the answer is exact by construction.
//   * uniqueness: an incircle of exactly 0 against the final r (four cocircular points on an empty circle), an
//     undecided nearest neighbour or a (near-)duplicate point make the frame TIED: the caller sends it through the
//     divide-and-conquer path, which keeps the reference's tie-breaking (north_star (2)). So does a frame whose
//     triangle count is not 2(n-1) - h (delaunay_tri.c:608).
// Triangles are owned by their lexicographically smallest vertex and emitted as (p, n1, n2) with n2 the
// counter-clockwise successor of n1 around p: the vertex roles of the reference's convertTrisFreeAdj
// (delaunay_tri.c:606-635), so that area_tri sees the same operands in the same order; only the order of the sum over
// a frame's triangles differs (|rel| ~ 1e-16 against the 1e-12 contract).
//
// Host-compilable (tests/emu only), like the other *_core headers.
#pragma once
#include "gt_predicates.cuh"

namespace gt {

struct alignas(16) SPt { double x, y; };

enum { STAR_TIE = 1, STAR_DUP = 2, STAR_GUARD = 4, STAR_RANGE = 8 };

// The grid of one frame: points in cell order (cell = row * gx + col), cell c holds positions cstart[c] .. cstart[c+1]-1.
template <class Idx>
struct StarGrid {
    const SPt* pt;             // [n] cell-sorted coordinates
    const Idx* cstart;         // [gx * gy + 1]
    int n, gx, gy;
    double x0, y0, sx, sy;     // col = trunc((x - x0) * sx) clamped to [0, gx-1]; row likewise: monotone in x / y
    double xmin, xmax, ymin, ymax;   // exact extremes of the point set (its bounding box)
};

// work counters of the emulation build
#if defined(GT_STATS) && !defined(__CUDA_ARCH__)
struct StarStats { long cand, left, icc, exact, expand, steps, points, hullcert; };
extern StarStats g_star;
#define STAR_STAT(x, v) (g_star.x += (v))
#else
#define STAR_STAT(x, v) ((void)0)
#endif

GT_INLINE int star_cell(double v, double o, double s, int g) {
    double t = (v - o) * s;
    const double hi = (double)(g - 1);
    t = t > 0.0 ? t : 0.0;                                   // (also NaN -> 0)
    t = t < hi ? t : hi;
    return (int)t;
}

// window of cells [c0, c1] x [r0, r1] (inclusive)
struct StarWin { int c0, c1, r0, r1; };
template <class Idx>
GT_INLINE bool star_win_all(const StarGrid<Idx>& g, const StarWin& w) { return w.c0 == 0 && w.r0 == 0 && w.c1 == g.gx - 1 && w.r1 == g.gy - 1; }
template <class Idx>
GT_INLINE StarWin star_win_grow(const StarGrid<Idx>& g, const StarWin& w, int by) {
    StarWin o;
    o.c0 = w.c0 - by > 0 ? w.c0 - by : 0; o.r0 = w.r0 - by > 0 ? w.r0 - by : 0;
    o.c1 = w.c1 + by < g.gx - 1 ? w.c1 + by : g.gx - 1; o.r1 = w.r1 + by < g.gy - 1 ? w.r1 + by : g.gy - 1;
    return o;
}
// is the axis-parallel box [xa, xb] x [ya, yb] inside the window? (a bound beyond the point set's extent clamps to the
// outermost cell, which is what makes the window's outer edges "infinite")
template <class Idx>
GT_INLINE bool star_covers(const StarGrid<Idx>& g, const StarWin& w, double xa, double xb, double ya, double yb) {
    if (!(xa <= xb && ya <= yb)) return false;               // NaN
    return star_cell(xa, g.x0, g.sx, g.gx) >= w.c0 && star_cell(xb, g.x0, g.sx, g.gx) <= w.c1 &&
           star_cell(ya, g.y0, g.sy, g.gy) >= w.r0 && star_cell(yb, g.y0, g.sy, g.gy) <= w.r1;
}
// Cursor over the positions of window `w` that are not in window `in` (in.c0 > in.c1: no inner window), one contiguous
// run of cell-sorted positions ("segment") at a time.
struct StarCursor { int pos, end, s, ns, ring; StarWin w, in; };
template <class Idx>
GT_INLINE bool star_cur_next(const StarGrid<Idx>& g, StarCursor& c) {        // next non-empty segment; false: window exhausted
    while (c.s < c.ns) {
        int row, ca, cb;
        if (!c.ring) { row = c.w.r0 + c.s; ca = c.w.c0; cb = c.w.c1; }
        else {
            row = c.w.r0 + (c.s >> 1);
            const bool half = (c.s & 1) != 0, split = row >= c.in.r0 && row <= c.in.r1;
            if (!split) { ca = c.w.c0; cb = half ? -1 : c.w.c1; }
            else if (!half) { ca = c.w.c0; cb = c.in.c0 - 1; }
            else { ca = c.in.c1 + 1; cb = c.w.c1; }
        }
        ++c.s;
        if (ca > cb) continue;
        const int base = row * g.gx;
        c.pos = (int)g.cstart[base + ca]; c.end = (int)g.cstart[base + cb + 1];
        if (c.pos < c.end) return true;
    }
    return false;
}
template <class Idx>
GT_INLINE bool star_cur_begin(const StarGrid<Idx>& g, StarCursor& c, const StarWin& w, const StarWin& in) {
    c.w = w; c.in = in; c.ring = in.c0 <= in.c1; c.s = 0; c.ns = (w.r1 - w.r0 + 1) << (c.ring ? 1 : 0); c.pos = c.end = 0;
    return star_cur_next(g, c);
}

// The exact stages behind the two inlined filters, out of line (rare; registers of the hot loop matter more).
GT_NOINLINE int star_orient_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    STAR_STAT(exact, 1);
    return orient2d_sign<true>(ax, ay, bx, by, cx, cy);
}
GT_NOINLINE int star_incircle_exact(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy) {
    STAR_STAT(exact, 1);
    return incircle_sign<true>(ax, ay, bx, by, cx, cy, dx, dy);
}

// On the device the 32 lanes of a warp build 32 stars in LOCKSTEP: every loop below runs until no lane of the warp has
// work left in it (a vote), so that the lanes meet again after every candidate, every scan and every walk step instead of
// drifting apart through the nest of data-dependent loops (the first version, written as plain nested loops, ran at 4.4
// active lanes of 32). All 32 lanes must call star_point together; lanes without a point pass active = false.
#if defined(__CUDA_ARCH__)
#define STAR_ANY(x) (__any_sync(0xffffffffu, (x)) != 0)
#else
#define STAR_ANY(x) (x)
#endif

// Star of the point at position `ip`. emit(n1_pos, n2_pos) for every triangle (ip, n1, n2) it owns; *hull = the star is
// open (ip is a hull vertex). `w0` = half-width of the first window in cells. Returns STAR_* bits (0: certified).
template <class Idx, class Emit>
GT_INLINE int star_point(const StarGrid<Idx>& g, int ip, bool active, int w0, bool* hull, Emit&& emit) {
    *hull = false;
    if (g.n < 3) { *hull = active; return 0; }                // (uniform: n is the frame's)
    if (!active) ip = 0;
    const SPt P = g.pt[ip];
    const double px = P.x, py = P.y;
    if (active) STAR_STAT(points, 1);
    StarWin none; none.c0 = 1; none.c1 = 0; none.r0 = 1; none.r1 = 0;
    StarWin win;
    win.c0 = win.c1 = star_cell(px, g.x0, g.sx, g.gx); win.r0 = win.r1 = star_cell(py, g.y0, g.sy, g.gy);
    win = star_win_grow(g, win, w0);
    int err = 0;
    StarCursor cu;

    // ---------------------------------------------------------------- nearest neighbour
    int q0 = -1; double l1 = 1.7976931348623157e308, l2 = l1, q0dx = 0.0, q0dy = 0.0;
    {
        bool searching = active;
        bool busy = searching && star_cur_begin(g, cu, win, none);
        for (;;) {
            while (STAR_ANY(busy)) {
                if (busy) {
                    const int pos = cu.pos;
                    const SPt S = g.pt[pos];
                    const double dx = S.x - px, dy = S.y - py, l = dx * dx + dy * dy;
                    if (pos != ip) {
                        if (l < l1) { l2 = l1; l1 = l; q0 = pos; q0dx = dx; q0dy = dy; }
                        else if (l < l2) l2 = l;
                    }
                    if (++cu.pos >= cu.end) busy = star_cur_next(g, cu);
                }
            }
            if (searching) {
                bool done = star_win_all(g, win);
                if (!done && q0 >= 0) {
                    const double d = sqrt(l1) * (1.0 + 1e-9), m = 8.0 * GT_EPS * (fabs(px) + fabs(py) + d);
                    done = star_covers(g, win, px - d - m, px + d + m, py - d - m, py + d + m);
                }
                if (done) searching = false;
                else {
                    const StarWin in = win; win = star_win_grow(g, win, 1);
                    STAR_STAT(expand, 1);
                    busy = star_cur_begin(g, cu, win, in);
                }
            }
            if (!STAR_ANY(searching)) break;
        }
    }
    bool walking = active;
    if (walking && q0 < 0) { *hull = true; walking = false; }                 // (n >= 3 and no other point: cannot happen)
    if (walking && l1 < 4e-24) { err |= STAR_DUP; walking = false; }          // within 1e-12 in x and y: the reference's duplicate rule (delaunay_tri.c:440-452)
    // The rounded distances may not decide which point is nearest (the -corr points of a box edge are equidistant from
    // their two neighbours): then the start edge p -> q0 is verified instead -- it is Delaunay iff the best point on its
    // right is not inside the circle through the best point on its left (modes 0 and 1 below; a hull edge is Delaunay).
    const bool ambiguous = !(l2 > l1 * (1.0 + 16.0 * GT_EPS));

    // ---------------------------------------------------------------- the walks
    // mode 0 / 1: one step from q0 to its left / right (verification of an ambiguous start edge, nothing is emitted)
    // mode 2: counter-clockwise walk from q0 until the star closes or reaches the hull
    // mode 3: clockwise walk from q0 (open stars only) = the same code on mirrored y
    const int guard_max = 4 * g.n + 64;
    int guard = 0;
    int mode = ambiguous ? 0 : 2;
    int vl = -1;                                              // best left point of the start edge (mode 0)
    if (q0 < 0) q0 = 0;
    int cur = q0; double yd = 1.0, qdx = q0dx, qdy = q0dy, ql = l1;
    while (STAR_ANY(walking)) {
        // ---- one step for every walking lane: best left candidate of p -> cur; ties against the best so far are remembered
        int r = -1; double rdx = 0.0, rdy = 0.0, rl = 0.0, mc = 0.0, mcp = 0.0; bool tie = false;
        const SPt Q = g.pt[cur];
        bool scanning = walking;
        bool found = false;
        bool busy = scanning && star_cur_begin(g, cu, win, none);
        if (walking) { STAR_STAT(steps, 1); if (++guard > guard_max) { err |= STAR_GUARD; walking = scanning = busy = false; } }
        for (;;) {
            while (STAR_ANY(busy)) {
                if (busy) {
                    const int pos = cu.pos;
                    const SPt S = g.pt[pos];
                    const double sdx = S.x - px, sdy = yd * (S.y - py);
                    // left of p -> cur?  orient2d(cur, s, p): detleft = qdx * sdy, detright = qdy * sdx (predicates.c:1628-1651).
                    // One comparison decides: when the two products do not have the same strict sign, |det| = |dl| + |dr| = dsum.
                    const double dl = qdx * sdy, dr = qdy * sdx, det = dl - dr, dsum = fabs(dl) + fabs(dr);
                    int sg = (det > 0.0 && pos != cur) ? 1 : 0;                     // (s = cur: det is exactly 0; s = p: all zero)
                    if (!(fabs(det) >= GT_CCW_ERRBOUND_A * dsum) && pos != cur) {
                        sg = star_orient_exact(Q.x, yd * Q.y, S.x, yd * S.y, px, yd * py);
                        if (sg == GT_RANGE_ERR) { err |= STAR_RANGE; sg = 0; }
                    }
                    STAR_STAT(cand, 1);
                    if (sg > 0) {
                        STAR_STAT(left, 1);
                        const double sl = sdx * sdx + sdy * sdy;
                        int ig = 1;
                        if (r >= 0) {
                            // s inside circle(p, cur, r)?  incircle(a = r, b = cur, c = s, d = p), predicates.c:3238-3267
                            STAR_STAT(icc, 1);
                            const double cdxady = sdx * rdy, adxcdy = rdx * sdy;
                            const double ta = rl * det, tb = ql * (cdxady - adxcdy), tc = sl * mc;
                            const double idet = (ta + tb) + tc;
                            const double perm = dsum * rl + (fabs(cdxady) + fabs(adxcdy)) * ql + mcp * sl;
                            const double eb = GT_ICC_ERRBOUND_A * perm;
                            ig = idet > eb ? 1 : -1;
                            if (!(fabs(idet) > eb)) {
                                const SPt R = g.pt[r];
                                ig = star_incircle_exact(R.x, yd * R.y, Q.x, yd * Q.y, S.x, yd * S.y, px, yd * py);
                                if (ig == GT_RANGE_ERR) { err |= STAR_RANGE; ig = -1; }
                            }
                        }
                        if (ig > 0) { r = pos; rdx = sdx; rdy = sdy; rl = sl; mc = dr - dl; mcp = dsum; tie = false; }
                        else if (ig == 0) tie = true;
                    }
                    if (++cu.pos >= cu.end) busy = star_cur_next(g, cu);
                }
            }
            // ---- the scan of the window (or of its newest ring) is over: is the answer final?
            if (scanning) {
                bool done = false;
                if (star_win_all(g, win)) { found = r >= 0; done = true; }
                else if (r >= 0) {
                    // circumdisc of (0, q', r') inside the window?  D = 2 cross(q', r') = -2 mc > 0
                    if (fabs(mc) * 1048576.0 >= mcp) {
                        const double D = -2.0 * mc, iD = 1.0 / D;
                        const double ax = rdy * ql, bx = qdy * rl, ay = qdx * rl, by = rdx * ql;
                        const double ux = (ax - bx) * iD, uy = (ay - by) * iD;
                        const double R = sqrt(ux * ux + uy * uy);
                        const double relD = 4.0 * GT_EPS * mcp / fabs(mc) + 8.0 * GT_EPS;
                        const double ex = 4.0 * GT_EPS * (fabs(ax) + fabs(bx)) * fabs(iD) + fabs(ux) * relD;
                        const double ey = 4.0 * GT_EPS * (fabs(ay) + fabs(by)) * fabs(iD) + fabs(uy) * relD;
                        const double m = 2.0 * (ex + ey) + 1e-12 * R + 8.0 * GT_EPS * (fabs(px) + fabs(py) + R + fabs(ux) + fabs(uy));
                        const double cy = py + yd * uy;
                        if (star_covers(g, win, (px + ux) - R - m, (px + ux) + R + m, cy - R - m, cy + R + m)) { found = true; done = true; }
                    }
                } else {
                    // no point on the left inside the window: p -> cur is a hull edge if the whole bounding box is on its right
                    STAR_STAT(hullcert, 1);
                    bool right = true;
#pragma unroll 1
                    for (int c = 0; c < 4 && right; ++c) {
                        const double bx = (c & 1) ? g.xmax : g.xmin, by = (c & 2) ? g.ymax : g.ymin;
                        const int sg = star_orient_exact(Q.x, yd * Q.y, bx, yd * by, px, yd * py);
                        if (sg == GT_RANGE_ERR) { err |= STAR_RANGE; right = false; }
                        else if (sg > 0) right = false;
                    }
                    if (right) { found = false; done = true; }
                }
                if (done) scanning = false;
                else {
                    const StarWin in = win; win = star_win_grow(g, win, 1);
                    STAR_STAT(expand, 1);
                    busy = star_cur_begin(g, cu, win, in);
                }
            }
            if (!STAR_ANY(scanning)) break;
        }
        // ---- transition
        if (walking) {
            bool next_mode = false;
            if (err) walking = false;
            else if (found && tie) { err |= STAR_TIE; walking = false; }
            else if (mode < 2) {
                // verification of the start edge (hull edges are Delaunay)
                if (mode == 0) vl = found ? r : -1;
                else if (found && vl >= 0) {
                    const SPt L = g.pt[vl], R = g.pt[r];
                    const int ig = star_incircle_exact(px, py, Q.x, Q.y, L.x, L.y, R.x, R.y);     // (p, q0, left) is counter-clockwise
                    if (ig == GT_RANGE_ERR) { err |= STAR_RANGE; walking = false; }
                    else if (ig >= 0) { err |= STAR_TIE; walking = false; }     // cocircular, or q0 is not a neighbour after all: not decided here
                }
                next_mode = true;
            } else if (!found) {
                *hull = true;                                     // hull edge: this walk is over
                if (mode == 3) walking = false; else next_mode = true;
            } else {
                // triangle (p, cur, r), counter-clockwise in walk coordinates
                const bool own_q = qdx > 0.0 || (qdx == 0.0 && yd * qdy > 0.0), own_r = rdx > 0.0 || (rdx == 0.0 && yd * rdy > 0.0);
                if (own_q && own_r) { if (mode == 3) emit(r, cur); else emit(cur, r); }
                cur = r; qdx = rdx; qdy = rdy; ql = rl;
                if (r == q0) {
                    if (mode == 3) err |= STAR_GUARD;              // an open star cannot come back to its start
                    walking = false;                               // closed
                }
            }
            if (next_mode && walking) {
                ++mode;
                yd = (mode & 1) ? -1.0 : 1.0;
                cur = q0; qdx = q0dx; qdy = yd * q0dy; ql = l1;
            }
        }
    }
    return err;
}

// Grid dimensions for n points in the bounding box [xmin, xmax] x [ymin, ymax]: about `dens` points per cell, cells about
// square, at most max_cells of them. (Both the kernel and the emulation call this, so they bin identically.)
struct StarDims { int gx, gy; double sx, sy; };
GT_INLINE StarDims star_dims(int n, double xmin, double xmax, double ymin, double ymax, int max_cells, double dens) {
    StarDims d;
    const double w = xmax - xmin, h = ymax - ymin;
    double cells = (double)n / dens;
    if (cells > (double)max_cells) cells = (double)max_cells;
    if (cells < 1.0) cells = 1.0;
    int gx = 1, gy = 1;
    if (w > 0.0 && h > 0.0) {
        gx = (int)(sqrt(cells * w / h) + 0.5);
        gx = gx < 1 ? 1 : (gx > max_cells ? max_cells : gx);
        gy = (int)(cells / (double)gx);
        gy = gy < 1 ? 1 : gy;
    } else if (w > 0.0) gx = (int)cells;
    else if (h > 0.0) gy = (int)cells;
    while (gx * gy > max_cells) { if (gy > 1) --gy; else --gx; }
    d.gx = gx; d.gy = gy;
    d.sx = w > 0.0 ? (double)gx / w : 0.0;
    d.sy = h > 0.0 ? (double)gy / h : 0.0;
    return d;
}

}  // namespace gt
