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
// Cursor over the cell-sorted positions of a window, or of a window without an inner window (the ring a grown window adds):
// a queue of at most four rectangles of cells, walked row by row -- a row of a rectangle is one contiguous run of
// positions. The per-candidate cost is one increment and one compare; a new row costs two shared-memory loads.
struct StarCursor { int pos, end, ci, rows, width, part; };
template <class Idx>
GT_INLINE bool star_cur_row(const StarGrid<Idx>& g, StarCursor& c) {
    while (c.rows > 0) {
        const int a = (int)g.cstart[c.ci], b = (int)g.cstart[c.ci + c.width];
        c.ci += g.gx; --c.rows;
        if (a < b) { c.pos = a; c.end = b; return true; }
    }
    return false;
}
template <class Idx>
GT_INLINE bool star_cur_next(const StarGrid<Idx>& g, StarCursor& c, const StarWin& w, const StarWin& in) {
    if (star_cur_row(g, c)) return true;
    while (c.part < 4) {
        // the ring w \ in: rows above, rows below, columns left, columns right of the inner window
        int r0, r1, c0, c1;
        const int k = c.part++;
        if (k == 0) { r0 = w.r0; r1 = in.r0 - 1; c0 = w.c0; c1 = w.c1; }
        else if (k == 1) { r0 = in.r1 + 1; r1 = w.r1; c0 = w.c0; c1 = w.c1; }
        else if (k == 2) { r0 = in.r0; r1 = in.r1; c0 = w.c0; c1 = in.c0 - 1; }
        else { r0 = in.r0; r1 = in.r1; c0 = in.c1 + 1; c1 = w.c1; }
        if (r0 > r1 || c0 > c1) continue;
        c.ci = r0 * g.gx + c0; c.rows = r1 - r0 + 1; c.width = c1 - c0 + 1;
        if (star_cur_row(g, c)) return true;
    }
    return false;
}
// start on the whole window w (in.c0 > in.c1) or on the ring w \ in
template <class Idx>
GT_INLINE bool star_cur_begin(const StarGrid<Idx>& g, StarCursor& c, const StarWin& w, const StarWin& in) {
    c.pos = c.end = 0;
    if (in.c0 > in.c1) { c.part = 4; c.ci = w.r0 * g.gx + w.c0; c.rows = w.r1 - w.r0 + 1; c.width = w.c1 - w.c0 + 1; }
    else { c.part = 0; c.rows = 0; c.ci = 0; c.width = 0; }
    return star_cur_next(g, c, w, in);
}
// v or -v (m = 0 or the sign bit of the high word): mirrors a coordinate without a trip through the FP64 pipe
GT_INLINE double star_flip(double v, uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v));
#else
    return m ? -v : v;
#endif
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
// `lst` = the lane's list of LCAP candidate positions (put(k, pos) / get(k)): shared memory on the device.
template <int LCAP, bool FULL, class Idx, class List, class Emit>
GT_INLINE int star_point(const StarGrid<Idx>& g, int ip, bool active, int w0, List&& lst, bool* hull, Emit&& emit) {
    *hull = false;
    if (g.n < 3) { *hull = active; return 0; }                // (uniform: n is the frame's)
    if (!active) ip = 0;
    const SPt P = g.pt[ip];
    const double px = P.x, py = P.y;
    if (active) STAR_STAT(points, 1);
    StarWin none; none.c0 = 1; none.c1 = 0; none.r0 = 1; none.r1 = 0;
    StarWin win, in = none;
    win.c0 = win.c1 = star_cell(px, g.x0, g.sx, g.gx); win.r0 = win.r1 = star_cell(py, g.y0, g.sy, g.gy);
    win = star_win_grow(g, win, w0);
    int err = 0;
    StarCursor cu;

    // ---------------------------------------------------------------- nearest neighbour (and the two runners-up)
    const double BIG = 1.7976931348623157e308;
    int q0 = -1, b2 = -1; double l1 = BIG, l2 = BIG, l3 = BIG, q0dx = 0.0, q0dy = 0.0;
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
                        if (l < l1) { l3 = l2; l2 = l1; b2 = q0; l1 = l; q0 = pos; q0dx = dx; q0dy = dy; }
                        else if (l < l2) { l3 = l2; l2 = l; b2 = pos; }
                        else if (l < l3) l3 = l;
                    }
                    if (++cu.pos >= cu.end) busy = star_cur_next(g, cu, win, in);
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
                    in = win; win = star_win_grow(g, win, 1);
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
    if (walking && !(l2 > l1 * (1.0 + 16.0 * GT_EPS))) {
        // The rounded distances do not decide which of q0 and b2 is nearer (the -corr points of a box edge are equidistant
        // from their two neighbours). p -> q0 is still a Delaunay (Gabriel) edge if b2 is outside the circle with diameter
        // p q0 -- every other point is strictly farther from p than q0 and therefore outside it: (b2 - p).(b2 - q0) > 0.
        bool okay = l3 > l1 * (1.0 + 16.0 * GT_EPS);
        if (okay) {
            const SPt B = g.pt[b2], Q0 = g.pt[q0];
            const double t1 = (B.x - px) * (B.x - Q0.x), t2 = (B.y - py) * (B.y - Q0.y);
            okay = t1 + t2 > 1e-12 * (fabs(t1) + fabs(t2));
        }
        if (!okay) { err |= STAR_TIE; walking = false; }
    }

    // ---------------------------------------------------------------- the walks
    // mir = 0: counter-clockwise, mir = 1: clockwise = the same code on mirrored y.
    // FULL: the whole star -- counter-clockwise from q0 until it closes or reaches the hull, then (open stars) clockwise from
    //   q0; *hull says whether the star is open.
    // !FULL: only the arc of neighbours that are lexicographically greater than p -- a triangle is owned by its smallest
    //   vertex, so that arc holds every triangle p has to report (2 of the ~6 around an interior point). From a greater
    //   q0: counter-clockwise while the neighbours stay greater, then clockwise from q0 likewise. From another q0: towards
    //   the arc by the shorter way (q0 above p: clockwise), through it, and if the hull came first, the other way round.
    const int guard_max = 4 * g.n + 64;
    int guard = 0, mir = 0;
    uint32_t ym = 0;                                              // sign-bit mask of the mirrored walk
    if (q0 < 0) q0 = 0;
    const bool q0_greater = q0dx > 0.0 || (q0dx == 0.0 && q0dy > 0.0);
    bool in_arc = q0_greater, second = false;                     // (!FULL) cur is in the arc; this is the second direction
    if (!FULL && !q0_greater && q0dy > 0.0) { mir = 1; ym = 0x80000000u; }
    int cur = q0; double qdx = q0dx, qdy = star_flip(q0dy, ym), ql = l1;
    while (STAR_ANY(walking)) {
        // ---- one step for every walking lane: best left candidate of p -> cur; ties against the best so far are remembered
        int r = -1; double rdx = 0.0, rdy = 0.0, rl = 0.0, mc = 0.0, mcp = 0.0; bool tie = false;
        bool scanning = walking;
        bool found = false;
        in = none;
        bool busy = scanning && star_cur_begin(g, cu, win, in);
        if (walking) { STAR_STAT(steps, 1); if (++guard > guard_max) { err |= STAR_GUARD; walking = scanning = busy = false; } }
        for (;;) {
            // ---- pass 1: which candidates are left of p -> cur? They go to the lane's list. (Two passes instead of one loop that
            // does everything for a candidate: a candidate is left for half the lanes and beats the best so far for a tenth, and
            // a warp pays for every branch any of its lanes takes -- the one-loop version ran its incircle part at 9 lanes and
            // its update at 4. A full list ends the pass for the lane; it resumes after pass 2.)
            int cnt = 0;
            while (STAR_ANY(busy && cnt < LCAP)) {
                if (busy && cnt < LCAP) {
                    const int pos = cu.pos;
                    const SPt S = g.pt[pos];
                    const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
                    // orient2d(cur, s, p): detleft = qdx * sdy, detright = qdy * sdx (predicates.c:1628-1651). One comparison
                    // decides: when the two products do not have the same strict sign, |det| = |dl| + |dr| = dsum.
                    const double dl = qdx * sdy, dr = qdy * sdx, det = dl - dr, dsum = fabs(dl) + fabs(dr);
                    bool left = det > 0.0 && pos != cur;                             // (s = cur: det is exactly 0; s = p: all zero)
                    if (!(fabs(det) >= GT_CCW_ERRBOUND_A * dsum) && pos != cur) {
                        const SPt Q = g.pt[cur];
                        const int sg = star_orient_exact(Q.x, star_flip(Q.y, ym), S.x, star_flip(S.y, ym), px, star_flip(py, ym));
                        if (sg == GT_RANGE_ERR) err |= STAR_RANGE;
                        left = sg == 1;
                    }
                    STAR_STAT(cand, 1);
                    if (left) { lst.put(cnt, pos); ++cnt; STAR_STAT(left, 1); }
                    if (++cu.pos >= cu.end) busy = star_cur_next(g, cu, win, in);
                }
            }
            // ---- pass 2: the best of the list (and of what earlier rounds of this step found): s beats r when it is inside
            // circle(p, cur, r) = incircle(a = r, b = cur, c = s, d = p) > 0, predicates.c:3238-3267
            int k = 0;
            if (cnt > 0 && r < 0) {                                                  // the first left candidate of the step
                const int pos = lst.get(0);
                const SPt S = g.pt[pos];
                const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
                const double dl = qdx * sdy, dr = qdy * sdx;
                r = pos; rdx = sdx; rdy = sdy; rl = sdx * sdx + sdy * sdy; mc = dr - dl; mcp = fabs(dl) + fabs(dr); tie = false;
                k = 1;
            }
            while (STAR_ANY(k < cnt)) {
                if (k < cnt) {
                    const int pos = lst.get(k); ++k;
                    const SPt S = g.pt[pos];
                    const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
                    const double dl = qdx * sdy, dr = qdy * sdx, det = dl - dr, dsum = fabs(dl) + fabs(dr);
                    const double sl = sdx * sdx + sdy * sdy;
                    STAR_STAT(icc, 1);
                    const double cdxady = sdx * rdy, adxcdy = rdx * sdy;
                    const double ta = rl * det, tb = ql * (cdxady - adxcdy), tc = sl * mc;
                    const double idet = (ta + tb) + tc;
                    const double perm = dsum * rl + (fabs(cdxady) + fabs(adxcdy)) * ql + mcp * sl;
                    const double eb = GT_ICC_ERRBOUND_A * perm;
                    int ig = idet > eb ? 1 : -1;
                    if (!(fabs(idet) > eb)) {
                        const SPt R = g.pt[r], Q = g.pt[cur];
                        ig = star_incircle_exact(R.x, star_flip(R.y, ym), Q.x, star_flip(Q.y, ym), S.x, star_flip(S.y, ym), px, star_flip(py, ym));
                        if (ig == GT_RANGE_ERR) { err |= STAR_RANGE; ig = -1; }
                    }
                    if (ig > 0) { r = pos; rdx = sdx; rdy = sdy; rl = sl; mc = dr - dl; mcp = dsum; tie = false; }
                    else if (ig == 0) tie = true;
                }
            }
            if (STAR_ANY(busy)) continue;                                            // a lane's list was full: it goes on scanning
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
                        const double cy = py + star_flip(uy, ym);
                        if (star_covers(g, win, (px + ux) - R - m, (px + ux) + R + m, cy - R - m, cy + R + m)) { found = true; done = true; }
                    }
                } else {
                    // no point on the left inside the window: p -> cur is a hull edge if the whole bounding box is on its right
                    STAR_STAT(hullcert, 1);
                    const SPt Q = g.pt[cur];
                    bool right = true;
#pragma unroll 1
                    for (int c = 0; c < 4 && right; ++c) {
                        const double bx = (c & 1) ? g.xmax : g.xmin, by = (c & 2) ? g.ymax : g.ymin;
                        const int sg = star_orient_exact(Q.x, star_flip(Q.y, ym), bx, star_flip(by, ym), px, star_flip(py, ym));
                        if (sg == GT_RANGE_ERR) { err |= STAR_RANGE; right = false; }
                        else if (sg > 0) right = false;
                    }
                    if (right) { found = false; done = true; }
                }
                if (done) scanning = false;
                else {
                    in = win; win = star_win_grow(g, win, 1);
                    STAR_STAT(expand, 1);
                    busy = star_cur_begin(g, cu, win, in);
                }
            }
            if (!STAR_ANY(scanning)) break;
        }
        // ---- transition
        if (walking) {
            // triangle (p, cur, r) is counter-clockwise in walk coordinates; p owns it when cur and r are greater than p
            const bool own_q = qdx > 0.0 || (qdx == 0.0 && star_flip(qdy, ym) > 0.0), own_r = rdx > 0.0 || (rdx == 0.0 && star_flip(rdy, ym) > 0.0);
            bool turn = false;                                     // this direction is over
            if (err) walking = false;
            else if (found && tie) { err |= STAR_TIE; walking = false; }
            else if (!found) { *hull = true; turn = true; }        // hull edge
            else if (FULL) {
                if (own_q && own_r) { if (mir) emit(r, cur); else emit(cur, r); }
                cur = r; qdx = rdx; qdy = rdy; ql = rl;
                if (r == q0) {
                    if (mir) err |= STAR_GUARD;                    // an open star cannot come back to its start
                    walking = false;                               // closed
                }
            } else {
                if (in_arc && !own_r) turn = true;                 // left the arc: nothing owned beyond
                else {
                    if (in_arc) { if (mir) emit(r, cur); else emit(cur, r); }
                    in_arc = own_r;
                    cur = r; qdx = rdx; qdy = rdy; ql = rl;
                    if (r == q0) { err |= STAR_GUARD; walking = false; }      // (a closed star has neighbours on both sides of p)
                }
            }
            if (turn && walking) {
                // the other direction from q0 -- unless this was it, or (!FULL) the arc has been walked through already
                // (a walk leaves the arc the moment it steps out of it: in_arc == false here means it never got in)
                const bool again = !second && (FULL || q0_greater || !in_arc);
                if (!again) walking = false;
                else {
                    second = true; mir ^= 1; ym ^= 0x80000000u; in_arc = q0_greater;
                    cur = q0; qdx = q0dx; qdy = star_flip(q0dy, ym); ql = l1;
                }
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
