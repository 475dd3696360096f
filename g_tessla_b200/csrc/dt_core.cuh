// dt_core.cuh -- per-frame Delaunay machinery shared by the batched sm_100a kernel.
//
// What it reproduces (decision for decision) from the reference, src/delaunay_tri.c:
//   ring insertion scan        insertNode        :210-245
//   ring removal / head rule   deleteNode        :247-268
//   first / pred / succ        :285-333  (here O(1): half-edge handles + twins, no searches)
//   lower / upper tangents     lct / uct         :337-406
//   merge loop                 ord_dtriangulate  :535-599
//   2- and 3-point base cases  ord_dtriangulate  :516-531
//   triangle enumeration       convertTrisFreeAdj :606-635
// What is different (B200-first): the reference's recursion becomes a bottom-up, level-
// synchronous sweep over the SAME recursion tree (same split index mid = (ia+ib)/2), one lane
// per independent merge; the malloc'd doubly-linked vertNode lists become a pool of 16-bit
// half-edge records in shared memory (edge e = half-edges 2e and 2e+1, twins adjacent), so
// pred/succ/cut never search a ring; freed edges are recycled through a lane-private list
// and a block-shared stack.
//
// Host-compilable (tests/emu only) -- see gt_predicates.cuh.
#pragma once
#include "gt_predicates.cuh"

namespace gt {

// The merge walks stay inlined in their kernels: an out-of-line merge receives the Frame by reference, which
// parks its pointers in local memory and costs a local load per ring access (measured: 126k -> 82k frames/s).
#define GT_MERGE_LINKAGE GT_INLINE

// Filters of the vertex predicates: one shared out-of-line copy (0, default: measured 126k frames/s on cfg 3)
// or inlined at every call site (1: larger code, more spills, measured 51k frames/s with out-of-line merges).
#ifndef GT_PRED_INLINE
#define GT_PRED_INLINE 0
#endif
#if GT_PRED_INLINE
#define GT_PRED_LINKAGE GT_INLINE
#else
#define GT_PRED_LINKAGE GT_NOINLINE
#endif

typedef uint16_t u16;
#define GT_NIL 0xFFFFu      // empty marker of the 16-bit frame type (FrameT<u16>::NIL)

// Host-only work counters for the emulation build (tests/emu, -DGT_STATS); no-ops in the product.
#if defined(GT_STATS) && !defined(__CUDA_ARCH__)
struct Stats { long o2d, icc, steps, cuts, edges; };
extern Stats g_stats;
#define GT_STAT(x) (++g_stats.x)
#else
#define GT_STAT(x) ((void)0)
#endif

struct Pt { double x, y; };

// Per-frame working set. Idx = 16-bit indices for the batched kernel (every pointer is shared memory,
// n <= ~3.5k points) or 32-bit indices for the single-large-frame path (global memory).
template <class Idx>
struct FrameT {
    typedef Idx idx;
    enum : uint32_t { NIL = (uint32_t)(Idx)~(Idx)0 };     // empty marker (enum: usable in device code without a definition)
    Pt* pt;          // [n] sorted vertex coordinates
    Idx* head;       // [n] ring entry half-edge per vertex ("first", reference :285-289)
    Idx* dst;        // [2*cap_e] destination vertex of half-edge h
    Idx* nxt;        // [2*cap_e] CCW link in the origin's ring   (reference vertNode.next)
    Idx* prv;        // [2*cap_e] CW link in the origin's ring    (reference vertNode.prev)
    uint32_t* bump;  // number of edges ever taken from the pool
    uint32_t* stack; // shared free-edge stack head (edge index, NIL if empty); nullptr = no shared recycling
    uint32_t* err;   // sticky error bits (GT_ERR_*)
    int n;
    uint32_t cap_e;  // edge capacity of the pool
};
typedef FrameT<u16> Frame;
enum { GT_ERR_POOL = 1, GT_ERR_RANGE = 2, GT_ERR_RING = 4 };

// Lane state that survives across merges: a private list of free edges (linked through nxt[2e]);
// the empty list is F::NIL of the frame type in use.
struct Lane { uint32_t free_head; };

#if defined(__CUDA_ARCH__)
GT_INLINE uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
GT_INLINE uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
GT_INLINE void atomic_or_u32(uint32_t* p, uint32_t v) { atomicOr(p, v); }
#else
GT_INLINE uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
GT_INLINE uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
GT_INLINE void atomic_or_u32(uint32_t* p, uint32_t v) { *p |= v; }
#endif

// ---------------------------------------------------------------- predicates on vertices
// Out of line on the device: the merge code calls them from ~20 sites, and one shared copy keeps
// the kernel's register count and instruction-cache footprint down (the walk is latency bound,
// a call costs less than the i-cache misses of 20 inlined incircle bodies).
GT_PRED_LINKAGE bool ccw_pts(const Pt* pt, uint32_t* err, int a, int b, int c) {
    GT_STAT(o2d);
    Pt pa = pt[a], pb = pt[b], pc = pt[c];
    int s = orient2d_sign(pa.x, pa.y, pb.x, pb.y, pc.x, pc.y);
    if (s == GT_RANGE_ERR) { atomic_or_u32(err, GT_ERR_RANGE); return false; }
    return s > 0;
}
GT_PRED_LINKAGE bool in_circle_pts(const Pt* pt, uint32_t* err, int a, int b, int c, int d) {
    // incircle with a repeated point is exactly 0 (two equal rows) -- the commonest
    // degenerate call of the merge loop (r2 == li when ri has two neighbours); skip the math.
    if (a == d || b == d || c == d || a == b || a == c || b == c) return false;
    GT_STAT(icc);
    Pt pa = pt[a], pb = pt[b], pc = pt[c], pd = pt[d];
    int s = incircle_sign(pa.x, pa.y, pb.x, pb.y, pc.x, pc.y, pd.x, pd.y);
    if (s == GT_RANGE_ERR) { atomic_or_u32(err, GT_ERR_RANGE); return false; }
    return s > 0;
}
template <class F>
GT_INLINE bool ccw(const F& f, int a, int b, int c) { return ccw_pts(f.pt, f.err, a, b, c); }
template <class F>
GT_INLINE bool right_of(const F& f, int x, int ea, int eb) { return ccw(f, x, eb, ea); }   // reference :141-145
template <class F>
GT_INLINE bool left_of(const F& f, int x, int ea, int eb) { return ccw(f, x, ea, eb); }    // reference :147-151
template <class F>
GT_INLINE bool in_circle(const F& f, int a, int b, int c, int d) { return in_circle_pts(f.pt, f.err, a, b, c, d); }

// ---------------------------------------------------------------- edge pool
template <class F>
GT_INLINE int edge_alloc(const F& f, Lane& ln) {
    if (ln.free_head != F::NIL) {
        int e = (int)ln.free_head;
        ln.free_head = f.nxt[2 * e];
        return e;
    }
    // block-shared stack: pops only happen inside a merge phase, pushes only between phases
    // (flush_lane), so there is no ABA window.
    uint32_t top = f.stack ? *(volatile uint32_t*)f.stack : F::NIL;
    while (top != F::NIL) {
        uint32_t next = f.nxt[2 * top];
        uint32_t seen = atomic_cas_u32(f.stack, top, next);
        if (seen == top) return (int)top;
        top = seen;
    }
    GT_STAT(edges);
    uint32_t e = atomic_add_u32(f.bump, 1u);
    if (e >= f.cap_e) { atomic_or_u32(f.err, GT_ERR_POOL); return -1; }
    return (int)e;
}
template <class F>
GT_INLINE void edge_free(const F& f, Lane& ln, int e) {
    f.nxt[2 * e] = (typename F::idx)ln.free_head;
    ln.free_head = (uint32_t)e;
}
// Return a lane's private free edges to the shared stack (call between merge phases only).
template <class F>
GT_INLINE void flush_lane(const F& f, Lane& ln) {
    if (!f.stack) { ln.free_head = F::NIL; return; }          // large-frame mode: freed edges are not recycled across merges
    while (ln.free_head != F::NIL) {
        uint32_t e = ln.free_head;
        ln.free_head = f.nxt[2 * e];
        uint32_t top = *(volatile uint32_t*)f.stack;
        for (;;) {
            f.nxt[2 * e] = (typename F::idx)top;
            uint32_t seen = atomic_cas_u32(f.stack, top, e);
            if (seen == top) break;
            top = seen;
        }
    }
}

// ---------------------------------------------------------------- rings
template <class F>
GT_INLINE void link_after(const F& f, int at, int h) {     // reference insertNodeAfter :202-208
    int nx = f.nxt[at];
    f.nxt[at] = (typename F::idx)h; f.prv[nx] = (typename F::idx)h; f.prv[h] = (typename F::idx)at; f.nxt[h] = (typename F::idx)nx;
}

// Insert half-edge h (origin `parent`, destination `in`) into parent's CCW ring at the position
// the reference's scan finds (insertNode :210-245), including its "first" update rule.
template <class F>
GT_INLINE void ring_insert(const F& f, int parent, int in, int h) {
    f.dst[h] = (typename F::idx)in;
    int hd = f.head[parent];
    if (hd == F::NIL) { f.head[parent] = (typename F::idx)h; f.prv[h] = (typename F::idx)h; f.nxt[h] = (typename F::idx)h; return; }
    if (right_of(f, in, parent, f.dst[hd])) {
        int cur = f.prv[hd];
        while (cur != hd && right_of(f, in, parent, f.dst[cur])) cur = f.prv[cur];
        if (cur == hd) {                       // right of every ray: new hull successor
            f.head[parent] = (typename F::idx)h;
            link_after(f, f.prv[cur], h);
        } else {
            link_after(f, cur, h);
        }
    } else {
        int cur = f.nxt[hd];
        while (cur != hd && left_of(f, in, parent, f.dst[cur])) cur = f.nxt[cur];
        if (f.dst[cur] == in) { atomic_or_u32(f.err, GT_ERR_RING); }   // duplicate edge: never on contract input
        link_after(f, f.prv[cur], h);
    }
}

// insertNode with the position known from the topology of the merge. The reference finds the slot
// by orientation scans from "first" (:216-238). Its first test (is the new neighbour right of the
// "first" ray?) also decides whether "first" moves, so it is kept, and so is the clockwise scan that
// follows a positive answer (:217-224). When the answer is negative the reference walks CCW until the
// new ray is no longer left of the current one, i.e. to the ray's angular slot; inside a merge that
// slot is known (next to the previous cross edge, or in the hull gap), so the walk and its orient2d
// calls are skipped. -DGT_VERIFY_HINT re-runs the walk and flags any disagreement (host fuzz only).
template <class F>
GT_INLINE void ring_insert_hint(const F& f, int parent, int in, int h, int after) {
    f.dst[h] = (typename F::idx)in;
    int hd = f.head[parent];
    if (right_of(f, in, parent, f.dst[hd])) {
        int cur = f.prv[hd];
        while (cur != hd && right_of(f, in, parent, f.dst[cur])) cur = f.prv[cur];
        if (cur == hd) {
            f.head[parent] = (typename F::idx)h;
            link_after(f, f.prv[cur], h);
        } else {
            link_after(f, cur, h);
        }
    } else {
#ifdef GT_VERIFY_HINT
        int cur = f.nxt[hd];
        while (cur != hd && left_of(f, in, parent, f.dst[cur])) cur = f.nxt[cur];
        if (f.prv[cur] != after) atomic_or_u32(f.err, GT_ERR_RING);
#endif
        link_after(f, after, h);
    }
}

// connectVerts (reference :270-275): returns the half-edge a->b (its twin b->a is h^1).
template <class F>
GT_INLINE int connect(const F& f, Lane& ln, int a, int b) {
    int e = edge_alloc(f, ln);
    if (e < 0) return -1;
    ring_insert(f, a, b, 2 * e);
    ring_insert(f, b, a, 2 * e + 1);
    return 2 * e;
}

template <class F>
GT_INLINE void ring_unlink(const F& f, int origin, int h) {   // reference deleteNode :247-268
    int p = f.prv[h], nx = f.nxt[h];
    f.nxt[p] = (typename F::idx)nx; f.prv[nx] = (typename F::idx)p;
    if (f.head[origin] == h) f.head[origin] = (nx == h) ? (typename F::idx)F::NIL : (typename F::idx)nx;
}
// cutVerts (reference :277-282) on the half-edge h = (a->b).
template <class F>
GT_INLINE void cut(const F& f, Lane& ln, int h) {
    int a = f.dst[h ^ 1], b = f.dst[h];
    ring_unlink(f, a, h);
    ring_unlink(f, b, h ^ 1);
    edge_free(f, ln, h >> 1);
    GT_STAT(cuts);
}

// ---------------------------------------------------------------- base cases (reference :516-531)
template <class F>
GT_INLINE void leaf(const F& f, Lane& ln, int ia, int ib) {
    if (ib - ia == 1) { connect(f, ln, ia, ib); return; }
    connect(f, ln, ia, ia + 1);
    connect(f, ln, ia + 1, ib);
    if (ccw(f, ia, ia + 1, ib) || ccw(f, ia, ib, ia + 1)) connect(f, ln, ia, ib);
}

// ---------------------------------------------------------------- merge (reference :535-599)
// Left half = sorted positions ia..mid, right half = mid+1..ib, both already triangulated.
template <class F>
GT_MERGE_LINKAGE void merge(const F& f, Lane& ln, int mid) {
    // lower common tangent (reference lct :337-370): walk starts at the rightmost vertex of the
    // left half (position mid) and the leftmost of the right half (mid+1).
    int lx = mid, ly = mid + 1;
    {
        int hr = f.head[ly];                    // ly -> first(ly)
        int hl = f.prv[f.head[lx]];             // lx -> pred(lx, first(lx))
        for (;;) {
            int rfast = f.dst[hr], lfast = f.dst[hl];
            if (right_of(f, rfast, lx, ly)) { hr = f.nxt[hr ^ 1]; ly = rfast; }        // succ(rfast, y)
            else if (right_of(f, lfast, lx, ly)) { hl = f.prv[hl ^ 1]; lx = lfast; }   // pred(lfast, x)
            else break;
        }
    }
    // upper common tangent (reference uct :373-406), from the same start, before any edit
    int ux = mid, uy = mid + 1;
    {
        int hl = f.head[ux];                    // ux -> first(ux)
        int hr = f.prv[f.head[uy]];             // uy -> pred(uy, first(uy))
        for (;;) {
            int rfast = f.dst[hr], lfast = f.dst[hl];
            if (left_of(f, rfast, ux, uy)) { hr = f.prv[hr ^ 1]; uy = rfast; }         // pred(rfast, y)
            else if (left_of(f, lfast, ux, uy)) { hl = f.nxt[hl ^ 1]; ux = lfast; }    // succ(lfast, x)
            else break;
        }
    }
    // First cross edge = lower tangent: both ends receive it in their hull gap (before "first").
    int li = lx, ri = ly;
    int e;                                      // e = li->ri, e^1 = ri->li
    {
        int ne = edge_alloc(f, ln);
        if (ne < 0) return;
        e = 2 * ne;
        ring_insert_hint(f, li, ri, e, f.prv[f.head[li]]);
        ring_insert_hint(f, ri, li, e ^ 1, f.prv[f.head[ri]]);
    }
    int guard = 8 * f.n + 64;
    while (li != ux || ri != uy) {
        if (--guard < 0) { atomic_or_u32(f.err, GT_ERR_RING); return; }
        GT_STAT(steps);
        bool a = false, b = false;
        int hr1 = f.prv[e ^ 1];                 // ri -> pred(ri, li)
        int r1 = f.dst[hr1];
        if (left_of(f, r1, li, ri)) {
            int hr2 = f.prv[hr1];
            int r2 = f.dst[hr2];
            while (in_circle(f, r1, li, ri, r2)) {
                cut(f, ln, hr1);
                hr1 = hr2; r1 = r2;
                hr2 = f.prv[hr1]; r2 = f.dst[hr2];
            }
        } else a = true;
        int hl1 = f.nxt[e];                     // li -> succ(li, ri)
        int l1 = f.dst[hl1];
        if (right_of(f, l1, ri, li)) {
            int hl2 = f.nxt[hl1];
            int l2 = f.dst[hl2];
            while (in_circle(f, li, ri, l1, l2)) {
                cut(f, ln, hl1);
                hl1 = hl2; l1 = l2;
                hl2 = f.nxt[hl1]; l2 = f.dst[hl2];
            }
        } else b = true;
        // ties prefer the right candidate (:588-593)
        const bool go_left = a ? true : (b ? false : in_circle(f, li, ri, r1, l1));
        int ne = edge_alloc(f, ln);
        if (ne < 0) return;
        if (go_left) {
            // new edge l1 -> ri: at l1 it follows (CCW) the edge back to the old li, at ri it
            // precedes the old cross edge ri -> li
            ring_insert_hint(f, l1, ri, 2 * ne, hl1 ^ 1);
            ring_insert_hint(f, ri, l1, 2 * ne + 1, f.prv[e ^ 1]);
            li = l1;
        } else {
            // new edge li -> r1: at li it follows the old cross edge li -> ri, at r1 it precedes
            // the edge back to the old ri
            ring_insert_hint(f, li, r1, 2 * ne, e);
            ring_insert_hint(f, r1, li, 2 * ne + 1, f.prv[hr1 ^ 1]);
            ri = r1;
        }
        e = 2 * ne;
    }
}

// ---------------------------------------------------------------- recursion tree
// The reference splits positions [ia, ib] at mid = (ia+ib)/2 until 2 or 3 remain (:516-539).
// Node k of depth d is found by walking k's bits from the top. Returns false if the walk hits a
// leaf before depth d (the node does not exist).
GT_INLINE bool tree_node(int n, int depth, int k, int* ia, int* ib) {
    int a = 0, b = n - 1;
    for (int lvl = depth - 1; lvl >= 0; --lvl) {
        if (b - a < 3) return false;
        int mid = (a + b) / 2;
        if ((k >> lvl) & 1) a = mid + 1; else b = mid;
    }
    *ia = a; *ib = b;
    return true;
}
GT_INLINE int tree_depth(int n) {       // deepest level that holds a node
    int d = 0, sz = n;                  // largest node size at depth d is ceil(n / 2^d)
    while (sz > 3) { sz = (sz + 1) / 2; ++d; }
    return d;
}
// One node of one level: base case or merge of its two finished children.
template <class F>
GT_INLINE void process_node(const F& f, Lane& ln, int depth, int k) {
    int ia, ib;
    if (!tree_node(f.n, depth, k, &ia, &ib)) return;
    if (ib - ia < 1) return;
    if (ib - ia < 3) leaf(f, ln, ia, ib);
    else merge(f, ln, (ia + ib) / 2);
}

// ---------------------------------------------------------------- enumeration (reference :606-635)
// Triangles owned by sorted vertex i = CCW-consecutive neighbour pairs (n1, n2) of its ring,
// starting at "first", with both neighbours sorting after i (the reference's "adjacency list not
// freed yet"), except the wrap-around pair when it spans the hull gap (:619-621).
// visit(n1, n2) is called in emission order; returns the count.
template <class F, class Visit>
GT_INLINE int vertex_triangles(const F& f, int i, Visit&& visit) {
    int hd = f.head[i];
    if (hd == F::NIL || f.nxt[hd] == hd) return 0;
    int cnt = 0, h = hd;
    do {
        int hn = f.nxt[h];
        int n1 = f.dst[h], n2 = f.dst[hn];
        if (n1 > i && n2 > i) {
            if (hn == hd && !right_of(f, n1, i, n2)) break;
            visit(n1, n2);
            ++cnt;
        }
        h = hn;
    } while (h != hd);
    return cnt;
}

// area_tri (reference include/gta_tri.h:91-99; GROMACS rvec_sub, cprod, norm with real = double).
GT_INLINE double tri_area3(double ax, double ay, double az, double bx, double by, double bz,
                           double cx, double cy, double cz) {
    double abx = bx - ax, aby = by - ay, abz = bz - az;
    double acx = cx - ax, acy = cy - ay, acz = cz - az;
    double px = aby * acz - abz * acy;
    double py = abz * acx - abx * acz;
    double pz = abx * acy - aby * acx;
    return sqrt(px * px + py * py + pz * pz) / 2.0;
}

}  // namespace gt
