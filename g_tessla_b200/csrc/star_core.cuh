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
struct alignas(8) SPtF { float x, y; };                       // the same point rounded to float (screening only)

enum { STAR_TIE = 1, STAR_DUP = 2, STAR_GUARD = 4, STAR_RANGE = 8, STAR_DEFER = 16 };

// The grid of one frame: points in cell order (cell = row * gx + col), cell c holds positions cstart[c] .. cstart[c+1]-1.
// GL = 0: the frame's arrays are shared memory (the batched kernel); GL = 1: global memory (one large frame);
// GL = 2: a TILE of a large frame's grid staged in shared memory (gta_large_star.cuh): gx x gy are the staged cells, the
// staged cell (0, 0) is cell (cx0, cy0) of the frame's ggx x ggy grid, and a side of the tile that is not a side of the
// frame's grid is OPEN: there are points beyond it. A search that would have to look beyond an open side gives up with
// STAR_DEFER (the point goes to the kernel that reads the whole frame from global memory); nothing is ever concluded
// from having scanned "everything" unless everything really was staged.
template <class Idx, int GL = 0>
struct StarGrid {
    const SPt* pt;             // [n] cell-sorted coordinates
    const SPtF* ptf;           // [n] the same, rounded to float
    float kf;                  // 2.5 * 2^-24 * (largest |coordinate| of the frame): what the rounding to float can move a coordinate difference by, with slack
    const Idx* cstart;         // [gx * gy + 1]
    uint32_t pt_s, ptf_s, cstart_s;   // device, frame in shared memory: the three arrays' addresses in the shared window (ld.shared, not generic loads)
    int n, gx, gy;
    double x0, y0, sx, sy;     // col = trunc((x - x0) * sx) clamped to [0, gx-1]; row likewise: monotone in x / y
    double xmin, xmax, ymin, ymax;   // exact extremes of the point set (its bounding box)
    const double* rowx;              // GL = 1 only (nullable): [2 r], [2 r + 1] = smallest and largest x of the points in grid row r (+inf, -inf: none)
    int cx0, cy0, ggx, ggy, open;    // GL = 2 only: the tile's place in the frame's grid; open = 1 left | 2 right | 4 bottom | 8 top
};

// Reads of the frame's arrays. On the device they are shared memory and the loads say so (the passes are out-of-line
// functions: behind a generic pointer the compiler issues generic loads, which wait on the long scoreboard).
#if defined(__CUDA_ARCH__)
template <class Idx, int GL> GT_INLINE SPt star_pt(const StarGrid<Idx, GL>& g, int pos) {
    if (GL == 1) return g.pt[pos];
    SPt v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(g.pt_s + 16u * (uint32_t)pos)); return v;
}
template <class Idx, int GL> GT_INLINE SPtF star_ptf(const StarGrid<Idx, GL>& g, int pos) {
    if (GL == 1) return g.ptf[pos];
    SPtF v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(g.ptf_s + 8u * (uint32_t)pos)); return v;
}
template <class Idx, int GL> GT_INLINE int star_cstart(const StarGrid<Idx, GL>& g, int c) {
    if (GL == 1) return (int)g.cstart[c];
    uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(g.cstart_s + 4u * (uint32_t)c)); return (int)v;
}
#else
template <class Idx, int GL> GT_INLINE SPt star_pt(const StarGrid<Idx, GL>& g, int pos) { return g.pt[pos]; }
template <class Idx, int GL> GT_INLINE SPtF star_ptf(const StarGrid<Idx, GL>& g, int pos) { return g.ptf[pos]; }
template <class Idx, int GL> GT_INLINE int star_cstart(const StarGrid<Idx, GL>& g, int c) { return (int)g.cstart[c]; }
#endif

// work counters of the emulation build
#if defined(GT_STATS) && !defined(__CUDA_ARCH__)
struct StarStats { long cand, left, icc, exact, expand, steps, points, hullcert, gridcand, rows, screened; };
extern StarStats g_star;
#define STAR_STAT(x, v) (g_star.x += (v))
#else
#define STAR_STAT(x, v) ((void)0)
#endif

// On the device the 32 lanes of a warp build 32 stars in LOCKSTEP: every loop below runs until no lane of the warp has
// work left in it (a vote), so that the lanes meet again after every candidate, every scan and every walk step instead of
// drifting apart through the nest of data-dependent loops (the first version, written as plain nested loops, ran at 4.4
// active lanes of 32). All 32 lanes must call star_point together; lanes without a point pass active = false.
#if defined(__CUDA_ARCH__)
#define STAR_ANY(x) (__any_sync(0xffffffffu, (x)) != 0)
#define STAR_BALLOT(x) __ballot_sync(0xffffffffu, (x))
#define STAR_FFS(m) (__ffs((int)(m)) - 1)
#define STAR_LANES 32
__device__ __forceinline__ int star_lane_id() { uint32_t v; asm("mov.u32 %0, %%laneid;" : "=r"(v)); return (int)v; }
__device__ __forceinline__ int star_bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double star_bcast(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}
#else
#define STAR_ANY(x) (x)
#define STAR_BALLOT(x) ((x) ? 1u : 0u)
#define STAR_FFS(m) 0
#define STAR_LANES 1
inline int star_lane_id() { return 0; }
inline int star_bcast(int v, int) { return v; }
inline double star_bcast(double v, int) { return v; }
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
// column / row of a coordinate in the grid's own numbering. A tile (GL = 2) bins exactly like the frame's grid (same map,
// shifted by whole cells); the _u forms may lie outside the tile -- that is how "beyond the staged cells" shows.
template <class Idx, int GL> GT_INLINE int star_col_u(const StarGrid<Idx, GL>& g, double x) { return GL == 2 ? star_cell(x, g.x0, g.sx, g.ggx) - g.cx0 : star_cell(x, g.x0, g.sx, g.gx); }
template <class Idx, int GL> GT_INLINE int star_row_u(const StarGrid<Idx, GL>& g, double y) { return GL == 2 ? star_cell(y, g.y0, g.sy, g.ggy) - g.cy0 : star_cell(y, g.y0, g.sy, g.gy); }
template <class Idx, int GL> GT_INLINE int star_col(const StarGrid<Idx, GL>& g, double x) { int c = star_col_u(g, x); if (GL == 2) { c = c > 0 ? c : 0; c = c < g.gx - 1 ? c : g.gx - 1; } return c; }
template <class Idx, int GL> GT_INLINE int star_row(const StarGrid<Idx, GL>& g, double y) { int r = star_row_u(g, y); if (GL == 2) { r = r > 0 ? r : 0; r = r < g.gy - 1 ? r : g.gy - 1; } return r; }
// does the window reach an open side of a tile? (its cells end there, the points do not)
template <class Idx, int GL> GT_INLINE bool star_win_open(const StarGrid<Idx, GL>& g, const StarWin& w) {
    if (GL != 2) return false;
    return ((g.open & 1) && w.c0 <= 0) || ((g.open & 2) && w.c1 >= g.gx - 1) || ((g.open & 4) && w.r0 <= 0) || ((g.open & 8) && w.r1 >= g.gy - 1);
}
template <class Idx, int GL>
GT_INLINE bool star_win_all(const StarGrid<Idx, GL>& g, const StarWin& w) { return w.c0 == 0 && w.r0 == 0 && w.c1 == g.gx - 1 && w.r1 == g.gy - 1 && (GL != 2 || g.open == 0); }
template <class Idx, int GL>
GT_INLINE StarWin star_win_grow(const StarGrid<Idx, GL>& g, const StarWin& w, int by) {
    StarWin o;
    o.c0 = w.c0 - by > 0 ? w.c0 - by : 0; o.r0 = w.r0 - by > 0 ? w.r0 - by : 0;
    o.c1 = w.c1 + by < g.gx - 1 ? w.c1 + by : g.gx - 1; o.r1 = w.r1 + by < g.gy - 1 ? w.r1 + by : g.gy - 1;
    return o;
}
// is the axis-parallel box [xa, xb] x [ya, yb] inside the window? (a bound beyond the point set's extent clamps to the
// outermost cell, which is what makes the window's outer edges "infinite")
template <class Idx, int GL>
GT_INLINE bool star_covers(const StarGrid<Idx, GL>& g, const StarWin& w, double xa, double xb, double ya, double yb) {
    if (!(xa <= xb && ya <= yb)) return false;               // NaN
    return star_col_u(g, xa) >= w.c0 && star_col_u(g, xb) <= w.c1 && star_row_u(g, ya) >= w.r0 && star_row_u(g, yb) <= w.r1;
}
// Cursor over the cell-sorted positions of a window, or of a window without an inner window (the ring a grown window adds):
// a queue of at most four rectangles of cells, walked row by row -- a row of a rectangle is one contiguous run of
// positions. The per-candidate cost is one increment and one compare; a new row costs two shared-memory loads.
struct StarCursor { int pos, end, ci, rows, width, part; };

// What has to be scanned when a step's answer cannot be certified from the window: a part of the half-plane on the walk's
// left of p -> cur, scan-converted against the rows of the grid -- per row one run of cells, minus cells scanned before --
// so that the cost follows the region's area, not the distance it reaches to. star_certify picks the regions:
//   no left candidate yet   the half-plane inside a box around p that doubles until one is found (or the box is everything:
//                           a hull edge); the hull side of a point near the hull is empty for a long way
//   a long edge p - cur     first a band of cells along the segment: its apex lies in the circumdisc of the first guess, and
//                           that disc holds ~|p cur|^2 points; a point near the segment subtends a wide angle, so the disc
//                           through it has a thin cap on the left
//   finally                 the left part of the circumdisc of (p, cur, r): whatever beats r lies in it, and the disc of a
//                           better r lies inside it -- this scan settles the step
struct StarRegion {
    int disc, band;               // clip to the circumdisc / to a band around the segment p - cur
    double px, py, A, B;          // left of p -> cur in real coordinates: B (x - px) < A (y - py)
    double slope, hrow;           // A / B (B != 0) and the height of a row of cells, computed once per region
    double cx, cy, R;             // the disc, centre and (conservative) radius in real coordinates
    double bw, bx0, bx1;          // the band: |x - X(y)| <= bw on the line's x at height y, and bx0 <= x <= bx1
    int row, row_end, seg, sb0, sb1;
    StarWin lim, skip;            // only cells of `lim`; the cells of `skip` have been scanned before
};
// cells [ca, cb] of `row` that may hold points of the region (false: none). Conservative: the bounds are widened by far
// more than their rounding, and the cell map is monotone.
template <class Idx, int GL>
GT_INLINE bool star_region_cols(const StarGrid<Idx, GL>& g, const StarRegion& rg, int row, int* ca, int* cb) {
    double y0 = g.ymin, y1 = g.ymax;
    if (g.sy > 0.0) {
        const double h = rg.hrow;
        const int grow = GL == 2 ? row + g.cy0 : row;                   // (the row's number in the frame's grid)
        y0 = g.y0 + grow * h; y1 = g.y0 + (grow + 1) * h;
        const double my = 1e-9 * (fabs(y0) + fabs(y1) + h);
        y0 -= my; y1 += my;
        if (grow == 0) y0 = fmin(y0, g.ymin);
        if (grow == (GL == 2 ? g.ggy : g.gy) - 1) y1 = fmax(y1, g.ymax);
    }
    double xa = g.xmin, xb = g.xmax;
    const double A = rg.A, B = rg.B, t0 = y0 - rg.py, t1 = y1 - rg.py;
    if (B != 0.0) {
        const double sl = rg.slope, X0 = rg.px + sl * t0, X1 = rg.px + sl * t1;
        const double mx = 1e-9 * (fabs(rg.px) + fabs(sl * t0) + fabs(sl * t1));
        if (X0 == X0 && X1 == X1 && mx == mx) {                   // (an overflowing slope bounds nothing)
            if (B > 0.0) xb = fmin(xb, fmax(X0, X1) + mx); else xa = fmax(xa, fmin(X0, X1) - mx);
            if (rg.band) { xa = fmax(xa, fmin(X0, X1) - mx - rg.bw); xb = fmin(xb, fmax(X0, X1) + mx + rg.bw); }
        }
    } else if (A > 0.0) { if (!(t1 > 0.0)) return false; }
    else if (A < 0.0) { if (!(t0 < 0.0)) return false; }
    if (rg.band) { xa = fmax(xa, rg.bx0); xb = fmin(xb, rg.bx1); }
    if (GL == 1 && g.rowx) {
        // One large frame: the far searches -- a hull edge, the disc of a skinny triangle along the hull -- walk thousands of
        // rows along a side of the point set, between the set and its bounding box, where a row holds a cell or two of the
        // region and no point of it. The row's own extent in x says so before the disc's square root, the cell map and the
        // two loads of cell starts a whole grid row apart are paid.
        xa = fmax(xa, g.rowx[2 * (size_t)row]); xb = fmin(xb, g.rowx[2 * (size_t)row + 1]);
        if (!(xa <= xb)) { STAR_STAT(screened, 1); return false; }
    }
    if (rg.disc) {
        const double cy = rg.cy, R = rg.R;
        if (y1 < cy - R || y0 > cy + R) return false;
        const double dy = (cy >= y0 && cy <= y1) ? 0.0 : fmin(fabs(y0 - cy), fabs(y1 - cy));
        const double w2 = R * R - dy * dy;
        const double w = (w2 > 0.0 ? sqrt(w2) : 0.0) * (1.0 + 1e-9) + 1e-9 * R;
        xa = fmax(xa, rg.cx - w); xb = fmin(xb, rg.cx + w);
    }
    if (!(xa <= xb)) return false;
    int a = star_col(g, xa), b = star_col(g, xb);
    a = a > rg.lim.c0 ? a : rg.lim.c0; b = b < rg.lim.c1 ? b : rg.lim.c1;
    if (a > b) return false;
    *ca = a; *cb = b;
    return true;
}
template <class Idx, int GL>
GT_INLINE bool star_region_next(const StarGrid<Idx, GL>& g, StarCursor& c, StarRegion& rg) {
    for (;;) {
        int ca, cb;
        if (rg.seg == 0) {
            if (rg.row > rg.row_end) return false;
            STAR_STAT(rows, 1);
            if (!star_region_cols(g, rg, rg.row, &ca, &cb)) { ++rg.row; continue; }
            rg.sb0 = 1; rg.sb1 = 0;                                // second run of the row: none
            if (rg.row >= rg.skip.r0 && rg.row <= rg.skip.r1) {    // cells scanned before are skipped
                rg.sb0 = ca > rg.skip.c1 + 1 ? ca : rg.skip.c1 + 1; rg.sb1 = cb;
                cb = cb < rg.skip.c0 - 1 ? cb : rg.skip.c0 - 1;
            }
            rg.seg = 1;
        } else {
            ca = rg.sb0; cb = rg.sb1;
            rg.seg = 0; ++rg.row;
            if (ca > cb) continue;
            const int base = (rg.row - 1) * g.gx;
            c.pos = star_cstart(g, base + ca); c.end = star_cstart(g, base + cb + 1);
            if (c.pos < c.end) return true;
            continue;
        }
        if (ca > cb) continue;
        const int base = rg.row * g.gx;
        c.pos = star_cstart(g, base + ca); c.end = star_cstart(g, base + cb + 1);
        if (c.pos < c.end) return true;
    }
}
// One large frame in global memory (GL = 1): the far searches -- a hull edge, the disc of a skinny triangle along the hull
// -- walk thousands of rows that hold a cell or two of the region and nothing in it, and a row is a chain of ~40 dependent
// FP64 operations and two loads a whole grid row apart: ~1,000 cycles for the one lane that walks. So the walk from one
// non-empty row to the next is done by the WARP: the lanes take the region of the lane that needs a row (`leader`) and
// look at 32 consecutive rows at once. The leader then stands on the first row that holds a candidate (and goes through it
// with star_region_next, which finds it non-empty at once), or the region is exhausted.
// All 32 lanes call this together; returns the row (all lanes) or -1.
template <class Idx, int GL>
GT_INLINE int star_region_find_row(const StarGrid<Idx, GL>& g, const StarRegion& mine, int leader) {
    StarRegion R;
    R.disc = star_bcast(mine.disc, leader); R.band = star_bcast(mine.band, leader);
    R.px = star_bcast(mine.px, leader); R.py = star_bcast(mine.py, leader); R.A = star_bcast(mine.A, leader); R.B = star_bcast(mine.B, leader);
    R.slope = star_bcast(mine.slope, leader); R.hrow = star_bcast(mine.hrow, leader);
    R.cx = star_bcast(mine.cx, leader); R.cy = star_bcast(mine.cy, leader); R.R = star_bcast(mine.R, leader);
    R.bw = star_bcast(mine.bw, leader); R.bx0 = star_bcast(mine.bx0, leader); R.bx1 = star_bcast(mine.bx1, leader);
    R.row = star_bcast(mine.row, leader); R.row_end = star_bcast(mine.row_end, leader);
    R.lim.c0 = star_bcast(mine.lim.c0, leader); R.lim.c1 = star_bcast(mine.lim.c1, leader); R.lim.r0 = star_bcast(mine.lim.r0, leader); R.lim.r1 = star_bcast(mine.lim.r1, leader);
    R.skip.c0 = star_bcast(mine.skip.c0, leader); R.skip.c1 = star_bcast(mine.skip.c1, leader); R.skip.r0 = star_bcast(mine.skip.r0, leader); R.skip.r1 = star_bcast(mine.skip.r1, leader);
    R.seg = 0; R.sb0 = 1; R.sb1 = 0;
    const int lane = star_lane_id();
    for (int row0 = R.row; row0 <= R.row_end; row0 += STAR_LANES) {
        const int row = row0 + lane;
        bool has = false;
        int ca, cb;
        if (row <= R.row_end && star_region_cols(g, R, row, &ca, &cb)) {
            STAR_STAT(rows, 1);
            const int base = row * g.gx;
            int a1 = 1, b1 = 0;
            if (row >= R.skip.r0 && row <= R.skip.r1) {               // cells scanned before are skipped (as in star_region_next)
                a1 = ca > R.skip.c1 + 1 ? ca : R.skip.c1 + 1; b1 = cb;
                cb = cb < R.skip.c0 - 1 ? cb : R.skip.c0 - 1;
            }
            if (ca <= cb) has = star_cstart(g, base + ca) < star_cstart(g, base + cb + 1);
            if (!has && a1 <= b1) has = star_cstart(g, base + a1) < star_cstart(g, base + b1 + 1);
        }
        const uint32_t m = STAR_BALLOT(has);
        if (m) return row0 + STAR_FFS(m);
    }
    return -1;
}
// the second run of the row the cursor has just finished (seg == 1), if it has one and it is not empty
template <class Idx, int GL>
GT_INLINE bool star_region_second(const StarGrid<Idx, GL>& g, StarCursor& c, StarRegion& rg) {
    const int ca = rg.sb0, cb = rg.sb1;
    rg.seg = 0; ++rg.row;
    if (ca > cb) return false;
    const int base = (rg.row - 1) * g.gx;
    c.pos = star_cstart(g, base + ca); c.end = star_cstart(g, base + cb + 1);
    return c.pos < c.end;
}
template <class Idx, int GL>
GT_INLINE bool star_cur_row(const StarGrid<Idx, GL>& g, StarCursor& c) {
    while (c.rows > 0) {
        const int a = star_cstart(g, c.ci), b = star_cstart(g, c.ci + c.width);
        c.ci += g.gx; --c.rows;
        if (a < b) { c.pos = a; c.end = b; return true; }
    }
    return false;
}
template <class Idx, int GL>
GT_INLINE bool star_cur_next(const StarGrid<Idx, GL>& g, StarCursor& c, const StarWin& w, const StarWin& in) {
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
template <class Idx, int GL>
GT_INLINE bool star_cur_begin(const StarGrid<Idx, GL>& g, StarCursor& c, const StarWin& w, const StarWin& in) {
    c.pos = c.end = 0;
    if (in.c0 > in.c1) { c.part = 4; c.ci = w.r0 * g.gx + w.c0; c.rows = w.r1 - w.r0 + 1; c.width = w.c1 - w.c0 + 1; }
    else { c.part = 0; c.rows = 0; c.ci = 0; c.width = 0; }
    return star_cur_next(g, c, w, in);
}
// v or -v (m = 0 or the sign bit of the high word): mirrors a coordinate without a trip through the FP64 pipe
GT_INLINE float star_flipf(float v, uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(__float_as_int(v) ^ (int)m);
#else
    return m ? -v : v;
#endif
}
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

// One lane's state. It lives in local memory between the passes of a step; every pass is its own out-of-line function
// that loads what it needs into registers and writes back what it changed. (Written as one function, the compiler kept
// some fifty values live across the two hot loops and spilled in the middle of them: 10 % of all instructions were local
// loads and stores. This way the loops run in a small register set and the state moves once per pass.)
struct StarLane {
    int ip;                                     // the point (position in cell order)
    StarWin win;                                // the window its candidates come from
    int m, overflow;                            // candidates in the lane's list; the window holds more than the list
    int q0;                                     // nearest neighbour: the walks start at p -> q0
    int cur; uint32_t ym;                       // current edge p -> cur; sign-bit mask of the walk (clockwise walks run on mirrored y)
    int r, tie;                                 // best left candidate of the step so far, and whether something tied with it
    int err;
};
// Everything else a pass needs is a function of these positions and is recomputed from the shared-memory arrays when the
// pass starts (a few loads and FP64 operations per pass, instead of keeping ~25 more words per lane in local memory: the
// lanes' state has to stay in the part of L1 the shared memory leaves).
struct StarEdge { double px, py, qdx, qdy, ql; };
template <class Idx, int GL>
GT_INLINE StarEdge star_edge(const StarGrid<Idx, GL>& g, const StarLane& L) {
    StarEdge e;
    const SPt P = star_pt(g, L.ip), Q = star_pt(g, L.cur);
    e.px = P.x; e.py = P.y;
    e.qdx = Q.x - P.x; e.qdy = star_flip(Q.y - P.y, L.ym);
    e.ql = e.qdx * e.qdx + e.qdy * e.qdy;
    return e;
}
struct StarBest { double rdx, rdy, rl, mc, mcp; };           // r relative to p, |r - p|^2, cross(r', q') and the sum of its two products' magnitudes
template <class Idx, int GL>
GT_INLINE StarBest star_best(const StarGrid<Idx, GL>& g, const StarLane& L, const StarEdge& e) {
    StarBest b;
    const SPt R = star_pt(g, L.r >= 0 ? L.r : L.cur);
    b.rdx = R.x - e.px; b.rdy = star_flip(R.y - e.py, L.ym);
    b.rl = b.rdx * b.rdx + b.rdy * b.rdy;
    const double dl = e.qdx * b.rdy, dr = e.qdy * b.rdx;
    b.mc = dr - dl; b.mcp = fabs(dl) + fabs(dr);
    return b;
}


// Nearest neighbour of the lane's point (and the two runners-up); the window's points become the lane's candidate list.
template <int CCAP, class Idx, int GL, class List>
GT_NOINLINE void star_nn(const StarGrid<Idx, GL>& g, List lst, StarLane& L, bool active, int w0) {
    const int ip = L.ip;
    const SPt P = star_pt(g, ip);
    const double px = P.x, py = P.y;
    StarWin none; none.c0 = 1; none.c1 = 0; none.r0 = 1; none.r1 = 0;
    StarWin win, in = none;
    win.c0 = win.c1 = star_col(g, px); win.r0 = win.r1 = star_row(g, py);
    win = star_win_grow(g, win, w0);
    const double BIG = 1.7976931348623157e308;
    int q0 = -1, b2 = -1; double l1 = BIG, l2 = BIG, l3 = BIG;
    int m = 0; bool overflow = false;
    StarCursor cu;
    bool searching = active;
    bool busy = searching && star_cur_begin(g, cu, win, none);
    for (;;) {
        while (STAR_ANY(busy)) {
            if (busy) {
                const int pos = cu.pos;
                const SPt S = star_pt(g, pos);
                const double dx = S.x - px, dy = S.y - py, l = dx * dx + dy * dy;
                // the three smallest distances, without branches (as an if / else-if chain the compiler shuffled ~40 registers
                // per candidate): l1 <= l2 <= l3 always, so l < l1 implies l < l2
                const bool other = pos != ip;
                const double lx = other ? l : BIG;
                const bool c1 = lx < l1, c2 = lx < l2, c3 = lx < l3;
                l3 = c2 ? l2 : (c3 ? lx : l3);
                l2 = c1 ? l1 : (c2 ? lx : l2);
                b2 = c1 ? q0 : (c2 ? pos : b2);
                l1 = c1 ? lx : l1; q0 = c1 ? pos : q0;
                if (other) { if (m < CCAP) { lst.cput(m, pos); ++m; } else overflow = true; }
                if (++cu.pos >= cu.end) busy = star_cur_next(g, cu, win, in);
            }
        }
        if (searching) {
            bool done = star_win_all(g, win);
            if (!done && q0 >= 0) {
                const double d = sqrt(l1) * (1.0 + 1e-9), mg = 8.0 * GT_EPS * (fabs(px) + fabs(py) + d);
                done = star_covers(g, win, px - d - mg, px + d + mg, py - d - mg, py + d + mg);
            }
            if (done) searching = false;
            else if (star_win_open(g, win)) { L.err |= STAR_DEFER; searching = false; }     // (a tile: the next ring is not staged)
            else {
                in = win; win = star_win_grow(g, win, 1);
                STAR_STAT(expand, 1);
                busy = star_cur_begin(g, cu, win, in);
            }
        }
        if (!STAR_ANY(searching)) break;
    }
    L.win = win; L.m = m; L.overflow = overflow ? 1 : 0;
    L.q0 = q0;
    if (!active || q0 < 0) return;
    if (l1 < 4e-24) { L.err |= STAR_DUP; return; }            // within 1e-12 in x and y: the reference's duplicate rule (delaunay_tri.c:440-452)
    if (!(l2 > l1 * (1.0 + 16.0 * GT_EPS))) {
        // The rounded distances do not decide which of q0 and b2 is nearer (the -corr points of a box edge are equidistant
        // from their two neighbours). p -> q0 is still a Delaunay (Gabriel) edge if b2 is outside the circle with diameter
        // p q0 -- every other point is strictly farther from p than q0 and therefore outside it: (b2 - p).(b2 - q0) > 0.
        bool okay = l3 > l1 * (1.0 + 16.0 * GT_EPS);
        if (okay) {
            const SPt B = star_pt(g, b2), Q0 = star_pt(g, q0);
            const double t1 = (B.x - px) * (B.x - Q0.x), t2 = (B.y - py) * (B.y - Q0.y);
            okay = t1 + t2 > 1e-12 * (fabs(t1) + fabs(t2));
        }
        if (!okay) L.err |= STAR_TIE;
    }
}

// Pass 1 of a walk step, from the lane's candidate list: which candidates are left of p -> cur? They go to the lane's list
// of left candidates. Screened in float: orient2d(cur, s, p) = a d - b c on float coordinates is off by at most
// kf (|a| + |b|) + 5 2^-24 (|a d| + |b c|) from the real determinant; beyond that its sign is the real one. Closer to the
// line: the reference's own filter in double (predicates.c:1628-1651), then the exact stage.
// kk = next entry of the candidate list; returns kk | cnt << 16 (cnt = entries of the left list).
template <int LCAP, class Idx, int GL, class List>
GT_NOINLINE int star_pass1_list(const StarGrid<Idx, GL>& g, List lst, StarLane& L, int kk) {
    const int m = L.m, cur = L.cur;
    const uint32_t ym = L.ym;
    const StarEdge e = star_edge(g, L);
    const SPtF PF = star_ptf(g, L.ip);
    const float pxf = PF.x, pyf = PF.y;
    const float af = (float)e.qdx, bf = (float)e.qdy, kfab = g.kf * (fabsf(af) + fabsf(bf));
    int cnt = 0;
    // the float screen of one candidate, and what follows when it is not sure
    auto slow = [&](int pos) -> bool {
        const SPt S = star_pt(g, pos);
        const double px = e.px, py = e.py, qdx = e.qdx, qdy = e.qdy;
        const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
        const double dl = qdx * sdy, dr = qdy * sdx, det = dl - dr, dsum = fabs(dl) + fabs(dr);
        if (fabs(det) >= GT_CCW_ERRBOUND_A * dsum) return det > 0.0;
        const SPt Q = star_pt(g, cur);
        const int sg = star_orient_exact(Q.x, star_flip(Q.y, ym), S.x, star_flip(S.y, ym), px, star_flip(py, ym));
        if (sg == GT_RANGE_ERR) L.err |= STAR_RANGE;
        return sg == 1;
    };
    // two candidates per iteration: their loads and float chains are independent, and a lane has only three other warps
    // on its scheduler to hide them behind
    while (STAR_ANY(kk < m && cnt + 1 < LCAP)) {
        if (kk < m && cnt + 1 < LCAP) {
            const bool two = kk + 1 < m;
            const int pos0 = lst.cget(kk), pos1 = lst.cget(two ? kk + 1 : kk);
            kk += two ? 2 : 1;
            const SPtF F0 = star_ptf(g, pos0), F1 = star_ptf(g, pos1);
            const float c0 = F0.x - pxf, d0 = star_flipf(F0.y - pyf, ym), c1 = F1.x - pxf, d1 = star_flipf(F1.y - pyf, ym);
            const float dl0 = af * d0, dr0 = bf * c0, det0 = dl0 - dr0, dl1 = af * d1, dr1 = bf * c1, det1 = dl1 - dr1;
            bool left0 = det0 > 0.0f, left1 = det1 > 0.0f;
            const bool sure0 = fabsf(det0) > kfab + 6e-7f * (fabsf(dl0) + fabsf(dr0)) || pos0 == cur;
            const bool sure1 = fabsf(det1) > kfab + 6e-7f * (fabsf(dl1) + fabsf(dr1)) || pos1 == cur;
            if (!sure0) left0 = slow(pos0);
            if (!sure1 && two) left1 = slow(pos1);
            STAR_STAT(cand, two ? 2 : 1);
            if (left0 && pos0 != cur) { lst.put(cnt, pos0); ++cnt; STAR_STAT(left, 1); }
            if (two && left1 && pos1 != cur) { lst.put(cnt, pos1); ++cnt; STAR_STAT(left, 1); }
        }
    }
    return kk | (cnt << 16);
}

// Pass 1 from the grid: a window with more points than the candidate list holds, or the region a failed certificate
// asks for (star_certify). Rare. Returns cnt.
template <int LCAP, class Idx, int GL, class List>
GT_NOINLINE int star_pass1_grid(const StarGrid<Idx, GL>& g, List lst, StarLane& L, StarCursor& cu, StarRegion& rg, bool* busy_io, int cnt) {
    const int cur = L.cur, ip = L.ip, rbest = L.r;                 // (regions may overlap: r met again would tie with itself)
    const uint32_t ym = L.ym;
    const StarEdge e = star_edge(g, L);
    const double px = e.px, py = e.py, qdx = e.qdx, qdy = e.qdy;
    const StarWin win = L.win;
    StarWin none; none.c0 = 1; none.c1 = 0; none.r0 = 1; none.r1 = 0;
    bool busy = *busy_io;
    // (GL = 1) a region that has just been set up, or whose cursor ran out while the left list was full: the lane is busy but
    // stands on nothing -- it needs the next non-empty row, which the warp finds (star_region_find_row)
    bool need_row = GL == 1 && busy && cu.part == 5 && cu.pos >= cu.end;
    while (STAR_ANY(busy && (cnt < LCAP || need_row))) {
        if (busy && cnt < LCAP && !need_row) {
            const int pos = cu.pos;
            const SPt S = star_pt(g, pos);
            const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
            // orient2d(cur, s, p): detleft = qdx * sdy, detright = qdy * sdx. One comparison decides: when the two products do
            // not have the same strict sign, |det| = |dl| + |dr| = dsum.
            const double dl = qdx * sdy, dr = qdy * sdx, det = dl - dr, dsum = fabs(dl) + fabs(dr);
            const bool skip = pos == cur || pos == ip || pos == rbest;      // (s = cur: det is exactly 0; s = p: all zero)
            bool left = det > 0.0 && !skip;
            if (!(fabs(det) >= GT_CCW_ERRBOUND_A * dsum) && !skip) {
                const SPt Q = star_pt(g, cur);
                const int sg = star_orient_exact(Q.x, star_flip(Q.y, ym), S.x, star_flip(S.y, ym), px, star_flip(py, ym));
                if (sg == GT_RANGE_ERR) L.err |= STAR_RANGE;
                left = sg == 1;
            }
            STAR_STAT(cand, 1); STAR_STAT(gridcand, 1);
            if (left) { lst.put(cnt, pos); ++cnt; STAR_STAT(left, 1); }
            if (++cu.pos >= cu.end) {
                if (cu.part != 5) busy = star_cur_next(g, cu, win, none);
                else if (GL != 1) busy = star_region_next(g, cu, rg);
                else if (rg.seg == 0 || !star_region_second(g, cu, rg)) need_row = true;      // (the warp finds the next row, below)
            }
        }
        if (GL == 1) {
            for (uint32_t m = STAR_BALLOT(need_row); m != 0u; m = STAR_BALLOT(need_row)) {
                const int leader = STAR_FFS(m);
                const int row = star_region_find_row(g, rg, leader);
                if (star_lane_id() == leader) {
                    need_row = false;
                    if (row < 0) { busy = false; rg.row = rg.row_end + 1; }
                    else { rg.row = row; rg.seg = 0; busy = star_region_next(g, cu, rg); }
                }
            }
        }
    }
    *busy_io = busy;
    return cnt;
}

// Pass 2: the best of the left list (and of what earlier rounds of this step found): s beats r when it is inside
// circle(p, cur, r) = incircle(a = r, b = cur, c = s, d = p) > 0: the reference's filter (predicates.c:3238-3267), then exact.
template <class Idx, int GL, class List>
GT_NOINLINE void star_pass2(const StarGrid<Idx, GL>& g, List lst, StarLane& L, int cnt) {
    const uint32_t ym = L.ym;
    const StarEdge e = star_edge(g, L);
    const double px = e.px, py = e.py, qdx = e.qdx, qdy = e.qdy, ql = e.ql;
    const StarBest b0 = star_best(g, L, e);                   // (what an earlier round of this step found; unused while r < 0)
    int r = L.r; double rdx = b0.rdx, rdy = b0.rdy, rl = b0.rl, mc = b0.mc, mcp = b0.mcp; bool tie = L.tie != 0;
    int k = 0;
    if (cnt > 0 && r < 0) {                                                  // the first left candidate of the step
        const int pos = lst.get(0);
        const SPt S = star_pt(g, pos);
        const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
        const double dl = qdx * sdy, dr = qdy * sdx;
        r = pos; rdx = sdx; rdy = sdy; rl = sdx * sdx + sdy * sdy; mc = dr - dl; mcp = fabs(dl) + fabs(dr); tie = false;
        k = 1;
    }
    // s against the best so far: +1 inside circle(p, cur, r), 0 on it, -1 outside
    auto icc = [&](int pos, double sdx, double sdy, double det, double dsum, double sl) -> int {
        STAR_STAT(icc, 1);
        const double cdxady = sdx * rdy, adxcdy = rdx * sdy;
        const double ta = rl * det, tb = ql * (cdxady - adxcdy), tc = sl * mc;
        const double idet = (ta + tb) + tc;
        const double perm = dsum * rl + (fabs(cdxady) + fabs(adxcdy)) * ql + mcp * sl;
        const double eb = GT_ICC_ERRBOUND_A * perm;
        if (fabs(idet) > eb) return idet > 0.0 ? 1 : -1;
        const SPt R = star_pt(g, r), Q = star_pt(g, L.cur), S = star_pt(g, pos);
        const int ig = star_incircle_exact(R.x, star_flip(R.y, ym), Q.x, star_flip(Q.y, ym), S.x, star_flip(S.y, ym), px, star_flip(py, ym));
        if (ig == GT_RANGE_ERR) { L.err |= STAR_RANGE; return -1; }
        return ig;
    };
    // (two candidates per iteration, both against the same r and the second one tried again when the first wins, was
    // measured slower: 621 k against 690 k frames/s -- the second chain's registers spill)
    while (STAR_ANY(k < cnt)) {
        if (k < cnt) {
            const int pos = lst.get(k); ++k;
            const SPt S = star_pt(g, pos);
            const double sdx = S.x - px, sdy = star_flip(S.y - py, ym);
            const double dl = qdx * sdy, dr = qdy * sdx, det = dl - dr, dsum = fabs(dl) + fabs(dr);
            const double sl = sdx * sdx + sdy * sdy;
            const int ig = icc(pos, sdx, sdy, det, dsum, sl);
            if (ig > 0) { r = pos; rdx = sdx; rdy = sdy; rl = sl; mc = dr - dl; mcp = dsum; tie = false; }
            else if (ig == 0) tie = true;
        }
    }
    L.r = r; L.tie = tie ? 1 : 0;
}

// A scan is over (the window's, or a region's): is the answer final? 1: yes, r is the next neighbour; 2: yes, p -> cur
// is a hull edge; 0: not yet -- `rg` describes the next region to scan and the cursor stands on it (*busy: it is not empty).
// The search goes outwards from p in boxes that double (window, +4 cells, +8, +16, ...), always clipped to the half-plane on
// the left of p -> cur and, once there is a candidate r, to the circumdisc of (p, cur, r) -- whatever beats r lies in that
// disc, and the disc of a better r lies inside it. It ends when the disc fits into what has been scanned, or everything
// has been. A long edge first gets a band of cells along its segment: the disc of a first guess near p holds ~|p cur|^2
// points, while a point near the segment subtends a wide angle and leaves a thin cap.
// *rs = the step's progress: bits 0-7 the box level, STAR_RS_BAND the band has been scanned.
enum { STAR_RS_BAND = 0x100 };
// the rare part of star_certify: describe the next region. Returns 1 / 2 (final after all) or 0, | the step's new progress << 4.
template <class Idx, int GL>
GT_NOINLINE int star_certify_region(const StarGrid<Idx, GL>& g, StarLane& L, StarCursor& cu, StarRegion& rg, int rs, bool* busy,
                                    int disc, double dcx, double dcy, double dR) {
    const StarWin win = L.win;
    const uint32_t ym = L.ym;
    const int r = L.r;
    const int level = rs & 0xff;
    const StarWin seen = level == 0 ? win : star_win_grow(g, win, level < 24 ? (2 << level) : (1 << 26));      // scanned so far
    const StarEdge e = star_edge(g, L);
    const double px = e.px, py = e.py;
    StarWin all; all.c0 = 0; all.r0 = 0; all.c1 = g.gx - 1; all.r1 = g.gy - 1;
    // left of p -> cur in real coordinates (a clockwise walk's left is the real right)
    const double sgn = ym ? -1.0 : 1.0;
    rg.px = px; rg.py = py; rg.A = sgn * e.qdx; rg.B = sgn * star_flip(e.qdy, ym);
    rg.slope = rg.B != 0.0 ? rg.A / rg.B : 0.0; rg.hrow = g.sy > 0.0 ? 1.0 / g.sy : 0.0;
    rg.disc = disc; rg.cx = dcx; rg.cy = dcy; rg.R = dR;
    rg.band = 0; rg.bw = rg.bx0 = rg.bx1 = 0.0;
    rg.lim = all; rg.skip = seen; rg.seg = 0; rg.sb0 = 1; rg.sb1 = 0;
    if (r >= 0) {
        // (a sliver too flat for a trustworthy centre: no disc, the half-plane alone bounds the search)
        const double qx = e.qdx, qy = star_flip(e.qdy, ym);                   // cur - p, real coordinates
        const double cell = (g.sx > 0.0 && g.sy > 0.0) ? fmax(1.0 / g.sx, 1.0 / g.sy) : 0.0;
        if (!(rs & STAR_RS_BAND) && cell > 0.0 && (!rg.disc || rg.R > 6.0 * cell) && e.ql > 36.0 * cell * cell) {
            STAR_STAT(expand, 1);
            rs |= STAR_RS_BAND;
            rg.band = 1; rg.bw = 1.5 * cell;
            rg.bx0 = fmin(px, px + qx) - rg.bw; rg.bx1 = fmax(px, px + qx) + rg.bw;
            rg.row = star_row(g, fmin(py, py + qy) - rg.bw); rg.row_end = star_row(g, fmax(py, py + qy) + rg.bw);
            if (GL == 2) {                                           // a tile: the band must lie inside the staged cells
                StarWin bb; bb.c0 = star_col_u(g, rg.bx0); bb.c1 = star_col_u(g, rg.bx1);
                bb.r0 = star_row_u(g, fmin(py, py + qy) - rg.bw); bb.r1 = star_row_u(g, fmax(py, py + qy) + rg.bw);
                if (((g.open & 1) && bb.c0 < 0) || ((g.open & 2) && bb.c1 > g.gx - 1) || ((g.open & 4) && bb.r0 < 0) || ((g.open & 8) && bb.r1 > g.gy - 1)) {
                    L.err |= STAR_DEFER; *busy = false; return 1 | (rs << 4);
                }
            }
            cu.part = 5; cu.pos = cu.end = 0;                        // (part 5: the rows come from the lane's StarRegion)
            *busy = GL == 1 ? true : star_region_next(g, cu, rg);    // (GL = 1: the warp finds the first row, star_pass1_grid)
            return rs << 4;
        }
    } else if (level == 0) {
        // no point on the left in the window: p -> cur is a hull edge if the whole bounding box is on its right (exact)
        STAR_STAT(hullcert, 1);
        const SPt Q = star_pt(g, L.cur);
        bool right = true;
#pragma unroll 1
        for (int c = 0; c < 4 && right; ++c) {
            const double bx = (c & 1) ? g.xmax : g.xmin, by = (c & 2) ? g.ymax : g.ymin;
            const int sg = star_orient_exact(Q.x, star_flip(Q.y, ym), bx, star_flip(by, ym), px, star_flip(py, ym));
            if (sg == GT_RANGE_ERR) { L.err |= STAR_RANGE; right = false; }
            else if (sg > 0) right = false;
        }
        if (right) return 2 | (rs << 4);
    }
    // the next box
    STAR_STAT(expand, 1);
    if (GL == 2) {                                                   // a tile: the box must lie inside the staged cells
        const int by = level < 23 ? (4 << level) : (1 << 26);
        if (((g.open & 1) && win.c0 - by < 0) || ((g.open & 2) && win.c1 + by > g.gx - 1) || ((g.open & 4) && win.r0 - by < 0) || ((g.open & 8) && win.r1 + by > g.gy - 1)) {
            L.err |= STAR_DEFER; *busy = false; return 1 | (rs << 4);
        }
    }
    rg.lim = star_win_grow(g, win, level < 23 ? (4 << level) : (1 << 26));
    rg.row = rg.lim.r0; rg.row_end = rg.lim.r1;
    if (rg.disc) {
        const int d0 = star_row(g, rg.cy - rg.R), d1 = star_row(g, rg.cy + rg.R);
        rg.row = rg.row > d0 ? rg.row : d0; rg.row_end = rg.row_end < d1 ? rg.row_end : d1;
    }
    rs = (rs & ~0xff) | (level + 1);
    cu.part = 5; cu.pos = cu.end = 0;                        // (part 5: the rows come from the lane's StarRegion)
    *busy = GL == 1 ? true : star_region_next(g, cu, rg);    // (GL = 1: the warp finds the first row, star_pass1_grid)
    return rs << 4;
}
// Returns the answer (0 / 1 / 2) | the step's new progress << 4.
template <class Idx, int GL>
GT_NOINLINE int star_certify(const StarGrid<Idx, GL>& g, StarLane& L, StarCursor& cu, StarRegion& rg, int rs, bool* busy) {
    const StarWin win = L.win;
    const uint32_t ym = L.ym;
    const int r = L.r;
    const int level = rs & 0xff;
    const StarWin seen = level == 0 ? win : star_win_grow(g, win, level < 24 ? (2 << level) : (1 << 26));      // scanned so far
    if (star_win_all(g, seen)) return (r >= 0 ? 1 : 2) | (rs << 4);
    int disc = 0; double dcx = 0.0, dcy = 0.0, dR = 0.0;
    if (r >= 0) {
        // circumdisc of (0, q', r') inside what has been scanned?  D = 2 cross(q', r') = -2 mc > 0
        const StarEdge e = star_edge(g, L);
        const double px = e.px, py = e.py;
        const StarBest b = star_best(g, L, e);
        const double mc = b.mc, mcp = b.mcp;
        if (fabs(mc) * 1048576.0 >= mcp) {
            const double qdx = e.qdx, qdy = e.qdy, ql = e.ql, rdx = b.rdx, rdy = b.rdy, rl = b.rl;
            const double D = -2.0 * mc, iD = 1.0 / D;
            const double ax = rdy * ql, bx = qdy * rl, ay = qdx * rl, by = rdx * ql;
            const double ux = (ax - bx) * iD, uy = (ay - by) * iD;
            const double R = sqrt(ux * ux + uy * uy);
            const double relD = 4.0 * GT_EPS * mcp / fabs(mc) + 8.0 * GT_EPS;
            const double ex = 4.0 * GT_EPS * (fabs(ax) + fabs(bx)) * fabs(iD) + fabs(ux) * relD;
            const double ey = 4.0 * GT_EPS * (fabs(ay) + fabs(by)) * fabs(iD) + fabs(uy) * relD;
            const double mg = 2.0 * (ex + ey) + 1e-12 * R + 8.0 * GT_EPS * (fabs(px) + fabs(py) + R + fabs(ux) + fabs(uy));
            const double cx = px + ux, cy = py + star_flip(uy, ym);
            if (star_covers(g, seen, cx - R - mg, cx + R + mg, cy - R - mg, cy + R + mg)) return 1 | (rs << 4);
            if (R + mg < 1.7976931348623157e308) { disc = 1; dcx = cx; dcy = cy; dR = R + mg; }
        }
    }
    return star_certify_region(g, L, cu, rg, rs, busy, disc, dcx, dcy, dR);
}

// Star of the point at position `ip`. emit(n1_pos, n2_pos) for every triangle (ip, n1, n2) it owns; *hull = the star is
// open (ip is a hull vertex; FULL only). `w0` = half-width of the first window in cells. Returns STAR_* bits (0: certified).
// `lst` = the lane's two lists of positions in shared memory: its candidates (cput / cget, CCAP entries) and the left
// candidates of the current step (put / get, LCAP entries).
template <int LCAP, int CCAP, bool FULL, class Idx, int GL, class List, class Emit>
GT_INLINE int star_point(const StarGrid<Idx, GL>& g, int ip, bool active, int w0, List lst, bool* hull, Emit&& emit) {
    *hull = false;
    if (g.n < 3) { *hull = active; return 0; }                // (uniform: n is the frame's)
    if (!active) ip = 0;
    StarLane L;
    L.ip = ip; L.err = 0; L.cur = 0; L.r = -1; L.tie = 0; L.ym = 0u;
    if (active) STAR_STAT(points, 1);
    star_nn<CCAP>(g, lst, L, active, w0);                       // (also: duplicates and an undecided nearest neighbour -> L.err)
    bool walking = active && L.err == 0;
    if (walking && L.q0 < 0) { *hull = true; walking = false; }               // (n >= 3 and no other point: cannot happen)

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
    if (L.q0 < 0) L.q0 = 0;
    const int q0 = L.q0;
    bool q0_greater, q0_above;
    {
        const SPt P = star_pt(g, ip), Q0 = star_pt(g, q0);
        const double dx = Q0.x - P.x, dy = Q0.y - P.y;
        q0_greater = dx > 0.0 || (dx == 0.0 && dy > 0.0); q0_above = dy > 0.0;
    }
    bool in_arc = q0_greater, second = false;                     // (!FULL) cur is in the arc; this is the second direction
    if (!FULL && !q0_greater && q0_above) mir = 1;
    L.ym = mir ? 0x80000000u : 0u;
    L.cur = q0;
    StarWin none; none.c0 = 1; none.c1 = 0; none.r0 = 1; none.r1 = 0;
    StarCursor cu;
    StarRegion rg;
    while (STAR_ANY(walking)) {
        // ---- one step for every walking lane: best left candidate of p -> cur; ties against the best so far are remembered
        L.r = -1; L.tie = 0;
        bool scanning = walking;
        if (walking) { STAR_STAT(steps, 1); if (++guard > guard_max) { L.err |= STAR_GUARD; walking = scanning = false; } }
        // the candidates come from the lane's list; a lane whose window holds more than the list walks the grid instead
        int kk = (scanning && !L.overflow) ? 0 : L.m;
        bool busy = scanning && L.overflow && star_cur_begin(g, cu, L.win, none);
        int rs = 0;                                                  // progress through the step's regions (star_certify)
        int found = 0;
        for (;;) {
            // Two passes instead of one loop that does everything for a candidate: a candidate is left for half the lanes and
            // beats the best so far for a tenth, and a warp pays for every branch any of its lanes takes -- the one-loop version
            // ran its incircle part at 9 lanes and its update at 4. A full left list ends pass 1 for the lane; it resumes after pass 2.
            const int pc = star_pass1_list<LCAP>(g, lst, L, kk);
            kk = pc & 0xffff;
            int cnt = pc >> 16;
            if (STAR_ANY(busy)) cnt = star_pass1_grid<LCAP>(g, lst, L, cu, rg, &busy, cnt);
            star_pass2(g, lst, L, cnt);
            if (STAR_ANY(busy || kk < L.m)) continue;                                // a lane's left list was full: it goes on
            if (scanning) {
                const int cr = star_certify(g, L, cu, rg, rs, &busy);
                found = cr & 15; rs = cr >> 4;
                if (found) scanning = false;
            }
            if (!STAR_ANY(scanning)) break;
        }
        // ---- transition
        if (walking) {
            const int r = L.r, cur = L.cur;
            // triangle (p, cur, r) is counter-clockwise in walk coordinates; p owns it when cur and r are greater than p
            bool own_q, own_r;
            {
                const SPt P = star_pt(g, ip), Q = star_pt(g, cur), R = star_pt(g, r >= 0 ? r : cur);
                const double qx = Q.x - P.x, qy = Q.y - P.y, rx = R.x - P.x, ry = R.y - P.y;
                own_q = qx > 0.0 || (qx == 0.0 && qy > 0.0); own_r = rx > 0.0 || (rx == 0.0 && ry > 0.0);
            }
            bool turn = false;                                     // this direction is over
            if (L.err) walking = false;
            else if (found == 1 && L.tie) { L.err |= STAR_TIE; walking = false; }
            else if (found != 1) { *hull = true; turn = true; }    // hull edge
            else if (FULL) {
                if (own_q && own_r) { if (mir) emit(r, cur); else emit(cur, r); }
                L.cur = r;
                if (r == q0) {
                    if (mir) L.err |= STAR_GUARD;                  // an open star cannot come back to its start
                    walking = false;                               // closed
                }
            } else {
                if (in_arc && !own_r) turn = true;                 // left the arc: nothing owned beyond
                else {
                    if (in_arc) { if (mir) emit(r, cur); else emit(cur, r); }
                    in_arc = own_r;
                    L.cur = r;
                    if (r == q0) { L.err |= STAR_GUARD; walking = false; }    // (a closed star has neighbours on both sides of p)
                }
            }
            if (turn && walking) {
                // the other direction from q0 -- unless this was it, or (!FULL) the arc has been walked through already
                // (a walk leaves the arc the moment it steps out of it: in_arc == false here means it never got in)
                const bool again = !second && (FULL || q0_greater || !in_arc);
                if (!again) walking = false;
                else {
                    second = true; mir ^= 1; L.ym ^= 0x80000000u; in_arc = q0_greater;
                    L.cur = q0;
                }
            }
        }
    }
    return L.err;
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
