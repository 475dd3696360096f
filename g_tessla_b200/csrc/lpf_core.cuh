// lpf_core.cuh -- the Delaunay divide-and-conquer as a per-frame state machine ("lane per frame").
//
// dt_core.cuh::merge is a nest of data-dependent loops: when the lanes of a warp walk different merges
// they diverge and the warp runs them one after another. Here the same algorithm -- the same predicates
// on the same operands, hence the same rings as the reference (src/delaunay_tri.c:511-601, tangents
// :337-406, insertNode :210-245, deleteNode :247-268) -- is cut into ITERATIONS. One iteration of one
// frame runs up to TWO independent chains
//     [one edge = its two half-edge records] -> [one point (+ the ring entry of that vertex)]
//     -> one predicate (orient2d or incircle)
// followed by a short transition that edits the rings and names the phase of the next iteration.
// The two chains are the two sides of the reference's sequential code that cannot influence each other:
//   LP_TAN   upper and lower common tangent walk side by side (read-only walks over untouched halves)
//   LP_V     validity of the right and of the left candidate        (:556, :569)
//   LP_C     the right and the left incircle deletion chain          (:559-567, :572-580; they edit disjoint rings)
//   LP_INS   the two ring insertions of a new cross edge            (connectVerts :270-275; two different rings)
//   LP_SCAN  their clockwise scans, when "right of first" held      (:217-224)
// The choice between the candidates (:582-593) needs no load and is taken at the end of the LP_V / LP_C
// iteration that settles both chains. Each phase is a template instance lpf_step_t<PH>: straight-line
// code. The GPU scheduler (gta_lpf.cuh) bins the frames in flight by phase every round and lets a warp run
// ONE phase for 32 different frames: no divergence between lanes, and 64 dependent-load chains in flight per
// warp. Per frame the persistent state is LS_WORDS words (ls_pack / ls_unpack); the frame itself lives in
// global memory:
//   point  16 B {x, y}   record 8 B {dst, nxt, prv, -} per half-edge, twins adjacent (edge = 16 B, one load)
//   (the ring head of a vertex lives in its LPt record)
// Edge pool, free list and the cursor over the recursion tree are private to the frame: no atomics.
//
// Host-compilable: tests/emu drives the steps sequentially (through the packed state) and compares with the oracle.
#pragma once
#include "dt_core.cuh"

namespace gt {

// The predicate filters of the walker phases: inlined at their 15 call sites (default), or ONE out-of-line copy of each
// (-DGT_LPF_PRED_OL=1: 10 KB less hot code and shorter chunks, but the rounds got longer: 324 ms instead of 308 ms per
// 100k frames, measured with the merged insert / scan phase).
#ifndef GT_LPF_PRED_OL
#define GT_LPF_PRED_OL 0
#endif
#if GT_LPF_PRED_OL
template <bool AP> GT_NOINLINE int lp_orient2d(double ax, double ay, double bx, double by, double cx, double cy) { return orient2d_sign<AP>(ax, ay, bx, by, cx, cy); }
template <bool AP> GT_NOINLINE int lp_incircle(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy) { return incircle_sign<AP>(ax, ay, bx, by, cx, cy, dx, dy); }
#else
template <bool AP> GT_INLINE int lp_orient2d(double ax, double ay, double bx, double by, double cx, double cy) { return orient2d_sign<AP>(ax, ay, bx, by, cx, cy); }
template <bool AP> GT_INLINE int lp_incircle(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy) { return incircle_sign<AP>(ax, ay, bx, by, cx, cy, dx, dy); }
#endif

struct LRec { u16 dst, nxt, prv, pad; };                 // one half-edge
struct LPair { LRec a, b; };                             // a = the half-edge asked for, b = its twin

// Everything the path keeps per sorted vertex, in ONE 32-byte sector: a step that needs a neighbour's coordinates
// and its ring head gets both with one load (they were two arrays: 18 % more sector reads, and the kernel is bound
// by scattered sector reads from HBM).
struct alignas(32) LPt { double x, y, z; u16 head, perm; uint32_t pad; };

// What the walkers of one frame share (shared memory on the device): the edge pool's bump pointer, sticky error
// bits, and one "arrived" bit per inner node of the top of the recursion tree (heap numbering: node (d, k) is bit
// (1 << d) + k) -- the first child to arrive at a node sets the bit and retires, the second one runs the merge.
struct LMeta { uint32_t bump, err, ready; };

// Frame storage in global memory.
struct LFrame {
    LPt* pt;        // [n] sorted by (x, y)
    LRec* rec;      // [2 * cap_e]
    LMeta* m;
    int n;
    int cap_e;
};

#if defined(__CUDA_ARCH__)
GT_INLINE void lf_fence() { __threadfence_block(); }
GT_INLINE uint32_t lf_fetch_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
#else
GT_INLINE void lf_fence() {}
GT_INLINE uint32_t lf_fetch_or(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o | v; return o; }
#endif
#if defined(__CUDA_ARCH__)
GT_INLINE LPair lf_pair(const LFrame& f, int h) {
    const uint4 v = *reinterpret_cast<const uint4*>(f.rec + (h & ~1));
    LPair p; LRec r0, r1;
    r0.dst = (u16)(v.x & 0xffffu); r0.nxt = (u16)(v.x >> 16); r0.prv = (u16)(v.y & 0xffffu); r0.pad = 0;
    r1.dst = (u16)(v.z & 0xffffu); r1.nxt = (u16)(v.z >> 16); r1.prv = (u16)(v.w & 0xffffu); r1.pad = 0;
    if (h & 1) { p.a = r1; p.b = r0; } else { p.a = r0; p.b = r1; }
    return p;
}
GT_INLINE Pt lf_pt(const LFrame& f, int v) { const double2 d = *reinterpret_cast<const double2*>(f.pt + v); Pt p; p.x = d.x; p.y = d.y; return p; }
// coordinates and ring head of vertex v with one 256-bit load
GT_INLINE Pt lf_pth(const LFrame& f, int v, int& head) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(f.pt + v) : "memory");
    Pt p; p.x = __longlong_as_double((long long)a); p.y = __longlong_as_double((long long)b);
    head = (int)(d & 0xffffull);
    return p;
}
#else
GT_INLINE LPair lf_pair(const LFrame& f, int h) { LPair p; p.a = f.rec[h]; p.b = f.rec[h ^ 1]; return p; }
GT_INLINE Pt lf_pt(const LFrame& f, int v) { Pt p; p.x = f.pt[v].x; p.y = f.pt[v].y; return p; }
GT_INLINE Pt lf_pth(const LFrame& f, int v, int& head) { head = f.pt[v].head; return lf_pt(f, v); }
#endif
GT_INLINE int lf_head(const LFrame& f, int v) { return f.pt[v].head; }
GT_INLINE void lf_set_head(const LFrame& f, int v, int h) { f.pt[v].head = (u16)h; }
GT_INLINE void lf_set_nxt(const LFrame& f, int h, int v) { f.rec[h].nxt = (u16)v; }
GT_INLINE void lf_set_prv(const LFrame& f, int h, int v) { f.rec[h].prv = (u16)v; }
GT_INLINE void lf_set_rec(const LFrame& f, int h, int dst, int nxt, int prv) {
    LRec r; r.dst = (u16)dst; r.nxt = (u16)nxt; r.prv = (u16)prv; r.pad = 0; f.rec[h] = r;
}

// LP_DONE: the walker retired (its slot is free); LP_ROOT: the walker finished the root merge -- the frame is complete.
enum LPhase : int { LP_LEAF = 0, LP_TINIT, LP_TAN, LP_INS, LP_SCAN, LP_V, LP_C, LP_DONE, LP_ROOT, LP_NPHASE };

// Per-walker state between iterations. A walker owns the subtree rooted at depth `top` that contains its current
// node (d, k): it walks that subtree in post-order, and when the root is done it arrives at the parent (lp_finish_node).
struct LState {
    int phase;
    int d, k, ia, ib;                 // current node of the recursion tree
    uint32_t err;
    int top, free_head;               // depth of the owned subtree's root; private list of freed edges
    int first_edge, vR, vL, actR, actL, scan0, scan1;
    int adv;                          // the node is finished: the next LP_TINIT iteration first advances the tree cursor
    // tangents: upper (ux, uy) walks with (uHl, uHr), coordinates in pr1 / pl1; lower (li, ri) with (lHl, lHr), coordinates in
    // pli / pri. xSub: 0 = next test is the right hull's candidate, 1 = the left hull's, 2 = tangent found
    int ux, uy, uHl, uHr, uSub, lHl, lHr, lSub;
    // merge step
    int li, ri, e, headL, headR;              // headL / headR = ring entries ("first") of li / ri
    int hr1, r1, hr2, t1_prv, t1_nxt, hdR1;   // right candidate ri->r1, its clockwise neighbour, the twin's links at r1, first(r1)
    int hl1, l1, hl2, u1_prv, u1_nxt, hdL1;   // left candidate li->l1, its counter-clockwise neighbour, the twin's links at l1, first(l1)
    // the two ring insertions of the new edge h0 = li -> ri (endpoint 0 at li, endpoint 1 at ri)
    int h0, at0, nx0, at1, nx1;               // slots known from the topology of the merge (unused for the first edge: hull gap)
    int cur0, hd0, last0, cur1, hd1, last1;   // clockwise scans in flight
    int plNx0, plAt1;                         // where the half-edges went: succ(li, ri) and pred(ri, li) of the next step
    // cached coordinates: pli / pri = li / ri, pr1 / pl1 = the two candidates (tangents: upper end points)
    Pt pli, pri, pr1, pl1;
};

// ---------------------------------------------------------------- persistent form
constexpr int LS_WORDS = 37;
GT_INLINE uint32_t ls_lo(double v) { return (uint32_t)((uint64_t)dbl_bits(v) & 0xffffffffull); }
GT_INLINE uint32_t ls_hi(double v) { return (uint32_t)((uint64_t)dbl_bits(v) >> 32); }
GT_INLINE double ls_dbl(uint32_t lo, uint32_t hi) {
    const uint64_t b = ((uint64_t)hi << 32) | lo;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double v; memcpy(&v, &b, 8); return v;
#endif
}
#define LS_P2(a, b) ((uint32_t)((a) & 0xffff) | ((uint32_t)((b) & 0xffff) << 16))
template <class Put>
GT_INLINE void ls_pack(const LState& s, Put&& put) {
    put(0, (uint32_t)s.phase | ((uint32_t)s.d << 4) | ((uint32_t)s.k << 8) | ((s.err & 15u) << 20) | ((uint32_t)s.first_edge << 24) | ((uint32_t)s.vR << 25) |
           ((uint32_t)s.vL << 26) | ((uint32_t)s.actR << 27) | ((uint32_t)s.actL << 28) | ((uint32_t)s.scan0 << 29) | ((uint32_t)s.scan1 << 30) | ((uint32_t)s.adv << 31));
    put(1, (uint32_t)s.ia | ((uint32_t)s.ib << 12) | ((uint32_t)s.uSub << 24) | ((uint32_t)s.lSub << 26));
    put(2, LS_P2(s.top, s.free_head)); put(3, LS_P2(s.ux, s.uy)); put(4, LS_P2(s.uHl, s.uHr)); put(5, LS_P2(s.li, s.ri));
    put(6, LS_P2(s.lHl, s.lHr)); put(7, LS_P2(s.e, s.headL)); put(8, LS_P2(s.headR, s.hr1)); put(9, LS_P2(s.r1, s.hr2));
    put(10, LS_P2(s.t1_prv, s.t1_nxt)); put(11, LS_P2(s.hdR1, s.hl1)); put(12, LS_P2(s.l1, s.hl2)); put(13, LS_P2(s.u1_prv, s.u1_nxt));
    put(14, LS_P2(s.hdL1, s.h0)); put(15, LS_P2(s.at0, s.nx0)); put(16, LS_P2(s.at1, s.nx1)); put(17, LS_P2(s.cur0, s.hd0));
    put(18, LS_P2(s.last0, s.cur1)); put(19, LS_P2(s.hd1, s.last1)); put(20, LS_P2(s.plNx0, s.plAt1));
    put(21, ls_lo(s.pli.x)); put(22, ls_hi(s.pli.x)); put(23, ls_lo(s.pli.y)); put(24, ls_hi(s.pli.y));
    put(25, ls_lo(s.pri.x)); put(26, ls_hi(s.pri.x)); put(27, ls_lo(s.pri.y)); put(28, ls_hi(s.pri.y));
    put(29, ls_lo(s.pr1.x)); put(30, ls_hi(s.pr1.x)); put(31, ls_lo(s.pr1.y)); put(32, ls_hi(s.pr1.y));
    put(33, ls_lo(s.pl1.x)); put(34, ls_hi(s.pl1.x)); put(35, ls_lo(s.pl1.y)); put(36, ls_hi(s.pl1.y));
}
template <class Get>
GT_INLINE void ls_unpack(LState& s, Get&& get) {
    uint32_t w = get(0);
    s.phase = w & 15; s.d = (w >> 4) & 15; s.k = (w >> 8) & 0xfff; s.err = (w >> 20) & 15; s.first_edge = (w >> 24) & 1; s.vR = (w >> 25) & 1;
    s.vL = (w >> 26) & 1; s.actR = (w >> 27) & 1; s.actL = (w >> 28) & 1; s.scan0 = (w >> 29) & 1; s.scan1 = (w >> 30) & 1; s.adv = (w >> 31) & 1;
    w = get(1); s.ia = w & 0xfff; s.ib = (w >> 12) & 0xfff; s.uSub = (w >> 24) & 3; s.lSub = (w >> 26) & 3;
#define LS_U2(i, a, b) w = get(i); a = (int)(w & 0xffff); b = (int)(w >> 16)
    LS_U2(2, s.top, s.free_head); LS_U2(3, s.ux, s.uy); LS_U2(4, s.uHl, s.uHr); LS_U2(5, s.li, s.ri);
    LS_U2(6, s.lHl, s.lHr); LS_U2(7, s.e, s.headL); LS_U2(8, s.headR, s.hr1); LS_U2(9, s.r1, s.hr2);
    LS_U2(10, s.t1_prv, s.t1_nxt); LS_U2(11, s.hdR1, s.hl1); LS_U2(12, s.l1, s.hl2); LS_U2(13, s.u1_prv, s.u1_nxt);
    LS_U2(14, s.hdL1, s.h0); LS_U2(15, s.at0, s.nx0); LS_U2(16, s.at1, s.nx1); LS_U2(17, s.cur0, s.hd0);
    LS_U2(18, s.last0, s.cur1); LS_U2(19, s.hd1, s.last1); LS_U2(20, s.plNx0, s.plAt1);
#undef LS_U2
    s.pli.x = ls_dbl(get(21), get(22)); s.pli.y = ls_dbl(get(23), get(24));
    s.pri.x = ls_dbl(get(25), get(26)); s.pri.y = ls_dbl(get(27), get(28));
    s.pr1.x = ls_dbl(get(29), get(30)); s.pr1.y = ls_dbl(get(31), get(32));
    s.pl1.x = ls_dbl(get(33), get(34)); s.pl1.y = ls_dbl(get(35), get(36));
}

// Words a phase may modify (a word is rewritten from the unpacked fields, so a phase that changes one field of
// a word keeps the other one). tests/emu applies these masks after every iteration: a missing bit shows up
// as a triangulation mismatch.
#define LS_W(i) (1ull << (i))
#define LS_PT(i) (15ull << (i))
constexpr unsigned long long LS_M_NODE = LS_W(0) | LS_W(1);
constexpr unsigned long long LS_M_QUEUE = LS_W(0) | LS_W(2) | LS_W(5) | LS_W(7) | LS_W(8) | LS_W(14) | LS_W(15) | LS_W(16) | LS_PT(21) | LS_PT(25);   // lp_decide
constexpr unsigned long long LS_M_FINISH = LS_M_NODE | LS_W(7) | LS_W(8) | LS_W(11);                                                                  // lp_finish_connect
// `mask` is written back after every iteration of the phase; `tail` only when the iteration changed the phase (it then
// also queued a new edge or finished one: more fields move).
template <int PH> struct LsDirty { static constexpr unsigned long long mask = ~0ull, tail = 0; };
template <> struct LsDirty<LP_LEAF> { static constexpr unsigned long long mask = LS_W(2) | LS_M_NODE, tail = 0; };
template <> struct LsDirty<LP_TINIT> { static constexpr unsigned long long mask = LS_W(0) | LS_W(1) | LS_W(2) | LS_W(9) | LS_W(3) | LS_W(4) | LS_W(5) | LS_W(6) | LS_W(7) | LS_W(8) | LS_PT(21) | LS_PT(25) | LS_PT(29) | LS_PT(33), tail = 0; };
template <> struct LsDirty<LP_TAN> { static constexpr unsigned long long mask = LS_W(0) | LS_W(1) | LS_W(9) | LS_W(3) | LS_W(4) | LS_W(5) | LS_W(6) | LS_W(7) | LS_W(8) | LS_PT(21) | LS_PT(25) | LS_PT(29) | LS_PT(33), tail = LS_M_QUEUE; };
template <> struct LsDirty<LP_INS> { static constexpr unsigned long long mask = LS_W(0) | LS_W(7) | LS_W(8) | LS_W(17) | LS_W(18) | LS_W(19) | LS_W(20), tail = LS_M_FINISH; };
template <> struct LsDirty<LP_SCAN> { static constexpr unsigned long long mask = LsDirty<LP_INS>::mask, tail = LS_M_FINISH; };
template <> struct LsDirty<LP_V> { static constexpr unsigned long long mask = LS_W(0) | LS_W(9) | LS_W(10) | LS_W(11) | LS_W(12) | LS_W(13) | LS_W(14) | LS_PT(29) | LS_PT(33), tail = LS_M_QUEUE; };
template <> struct LsDirty<LP_C> { static constexpr unsigned long long mask = LS_W(0) | LS_W(2) | LS_W(7) | LS_W(8) | LS_W(9) | LS_W(10) | LS_W(11) | LS_W(12) | LS_W(13) | LS_W(14) | LS_PT(29) | LS_PT(33), tail = LS_M_QUEUE; };
template <unsigned long long MASK, class Put>
GT_INLINE void ls_pack_mask(const LState& s, Put&& put) {
    ls_pack(s, [&](int i, uint32_t v) { if ((MASK >> i) & 1ull) put(i, v); });
}
// write-back after one iteration of phase PH
template <int PH, class Put>
GT_INLINE void ls_writeback(const LState& s, Put&& put) {
    ls_pack_mask<LsDirty<PH>::mask>(s, put);
    if (s.phase != PH) ls_pack_mask<LsDirty<PH>::tail & ~LsDirty<PH>::mask>(s, put);
}

// ---------------------------------------------------------------- helpers
GT_INLINE void lp_free(const LFrame& f, LState& s, int e) { lf_set_nxt(f, 2 * e, s.free_head); s.free_head = e; }

// The current node (s.d, s.k) is complete: advance the post-order cursor over the reference's recursion tree inside the
// walker's subtree, or -- at the subtree's root -- arrive at the parent: the first of the two siblings to arrive sets the
// parent's bit and retires, the second one finds it set and owns the parent from then on (its merge reads both children's
// rings: the fences order them behind the bit). The root of the tree ends in LP_ROOT. A frame with an error set skips the
// merges above it (their inputs may be incomplete) but still arrives, so that the frame always completes.
// (One out-of-line copy, scalars in and a packed word out: the phases that end a node share it, and the walker kernel's
// code has to stay small -- it already misses the instruction cache.)
// result: phase | d << 4 | k << 8 | top << 20 | ia << 24 | ib << 36
GT_NOINLINE uint64_t lp_finish_core(int n, LMeta* m, int d, int k, int top) {
    int ia = 0, ib = 0, phase;
    for (;;) {
        if (d == top) {
            if (d == 0) { phase = LP_ROOT; break; }
            const uint32_t bit = 1u << ((1 << (d - 1)) + (k >> 1));
            lf_fence();
            const uint32_t old = lf_fetch_or(&m->ready, bit);
            if (!(old & bit)) { phase = LP_DONE; break; }
            lf_fence();
            d -= 1; k >>= 1; top = d;
            if (*(volatile uint32_t*)&m->err) continue;
        } else if (k & 1) {                                  // finished a right child: its parent is next, and it is a merge
            d -= 1; k >>= 1;
        } else {                                             // finished a left child: leftmost leaf of the sibling subtree
            k += 1;
            tree_node(n, d, k, &ia, &ib);
            while (ib - ia >= 3) { ib = (ia + ib) / 2; d += 1; k <<= 1; }      // left children down to the leaf (:535-538)
            phase = LP_LEAF;
            break;
        }
        tree_node(n, d, k, &ia, &ib);
        phase = LP_TINIT;
        break;
    }
    return (uint64_t)phase | (uint64_t)d << 4 | (uint64_t)k << 8 | (uint64_t)top << 20 | (uint64_t)ia << 24 | (uint64_t)ib << 36;
}
GT_INLINE void lp_finish_node(const LFrame& f, LState& s) {
    const uint64_t r = lp_finish_core(f.n, f.m, s.d, s.k, s.top);
    s.phase = (int)(r & 15u); s.d = (int)((r >> 4) & 15u); s.k = (int)((r >> 8) & 0xfffu); s.top = (int)((r >> 20) & 15u);
    s.ia = (int)((r >> 24) & 0xfffu); s.ib = (int)((r >> 36) & 0xfffu);
}

// Start a walker on the subtree rooted at node (top_d, top_k); the node must exist and hold at least two points.
GT_INLINE void lp_begin(const LFrame& f, LState& s, int top_d, int top_k) {
    s.err = 0; s.free_head = GT_NIL; s.adv = 0; s.top = top_d;
    s.first_edge = s.vR = s.vL = s.actR = s.actL = s.scan0 = s.scan1 = 0; s.uSub = s.lSub = 0;
    int d = top_d, k = top_k, ia, ib;
    tree_node(f.n, d, k, &ia, &ib);
    while (ib - ia >= 3) { ib = (ia + ib) / 2; d += 1; k <<= 1; }
    s.d = d; s.k = k; s.ia = ia; s.ib = ib;
    s.phase = LP_LEAF;
}
// Give up the walker's subtree after a hard error (pool exhausted, ring inconsistency): record it for the frame and arrive
// at the parent as if the subtree were complete.
GT_INLINE void lp_abort(const LFrame& f, LState& s) {
    atomic_or_u32(&f.m->err, s.err);
    s.k >>= (s.d - s.top); s.d = s.top;
    lp_finish_node(f, s);
}
// Deepest level at which every node of an n-point tree exists and holds at least two points, capped at `want`
// (a frame is split into 2^level subtrees that start in parallel).
GT_INLINE int lp_split_level(int n, int want) {
    int kk = 0;
    while (kk < want && (n >> (kk + 2)) >= 1) ++kk;          // n >= 2^(kk+1) after the increment
    return kk;
}
// One edge from the frame's shared pool (-1: exhausted).
GT_INLINE int lp_pool_take(const LFrame& f) {
    const uint32_t e = atomic_add_u32(&f.m->bump, 1u);
    return e < (uint32_t)f.cap_e ? (int)e : -1;
}

// A new cross edge li -> ri: take an edge from the pool and queue its two ring insertions.
//   spec    = records of the free list's head as loaded at the start of the iteration
//   freed   = this iteration pushed an edge on the free list after that load; then the head is that edge and
//             prev_head (the head before the push) is its successor -- no further load needed
GT_INLINE void lp_new_edge(const LFrame& f, LState& s, const LPair& spec, bool freed, int prev_head) {
    int ne;
    if (s.free_head != (int)GT_NIL) { ne = s.free_head; s.free_head = freed ? prev_head : (int)spec.a.nxt; }
    else { ne = lp_pool_take(f); if (ne < 0) { s.err |= GT_ERR_POOL; return; } }
    s.h0 = 2 * ne; s.scan0 = s.scan1 = 0;
    s.phase = LP_INS;
}

// The choice between the two candidates (reference :582-593; ties prefer the right one) and the new cross edge.
template <bool AP>
GT_INLINE void lp_decide(const LFrame& f, LState& s, const LPair& spec, bool freed, int prev_head) {
    bool go_left = true;
    if (s.vR) {
        go_left = false;
        if (s.vL && s.r1 != s.l1) {
            int sg = lp_incircle<AP>(s.pli.x, s.pli.y, s.pri.x, s.pri.y, s.pr1.x, s.pr1.y, s.pl1.x, s.pl1.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            go_left = sg > 0;
        }
    }
    if (go_left) {
        // new edge l1 -> ri: at l1 after the edge back to the old li, at ri before the old cross edge
        s.at0 = s.hl1 ^ 1; s.nx0 = s.u1_nxt; s.at1 = s.hr1; s.nx1 = s.e ^ 1;
        s.li = s.l1; s.pli = s.pl1; s.headL = s.hdL1;
    } else {
        // new edge li -> r1: at li after the old cross edge, at r1 before the edge back to the old ri
        s.at0 = s.e; s.nx0 = s.hl1; s.at1 = s.t1_prv; s.nx1 = s.hr1 ^ 1;
        s.ri = s.r1; s.pri = s.pr1; s.headR = s.hdR1;
    }
    lp_new_edge(f, s, spec, freed, prev_head);
}

// Both half-edges of the new edge are in their rings: next merge step, or the node is finished.
GT_INLINE void lp_finish_connect(const LFrame& f, LState& s) {
    s.e = s.h0;                                            // li -> ri
    s.hl1 = s.plNx0;                                       // succ(li, ri)
    s.hr1 = s.plAt1;                                       // pred(ri, li)
    // the merge is complete at the upper tangent. The tree cursor is advanced by the next LP_TINIT iteration, not
    // here: this is the tail of the insert step, the longest step of a round, and nearly every 32-slot chunk of it
    // holds a slot that finishes a merge
    if (s.li == s.ux && s.ri == s.uy) { s.adv = 1; s.phase = LP_TINIT; } else s.phase = LP_V;
}

// Put half-edge h (origin P, destination IN) between `at` and `nx` in P's ring (reference insertNodeAfter :202-208).
GT_INLINE void lp_place(const LFrame& f, LState& s, int which, int P, int IN, int h, int at, int nx, bool newhead, int hd) {
    lf_set_nxt(f, at, h); lf_set_prv(f, nx, h); lf_set_rec(f, h, IN, nx, at);
    if (newhead) lf_set_head(f, P, h);
    const int newhd = newhead ? h : hd;
    if (which == 0) { s.plNx0 = nx; s.headL = newhd; } else { s.plAt1 = at; s.headR = newhd; }
}

// ---------------------------------------------------------------- one iteration of phase PH
template <int PH, bool AP = true>
GT_INLINE void lpf_step_t(const LFrame& f, LState& s) {
    const LPair zero = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    // the free list's head, for the phases that may need a new edge at the end of the iteration
    LPair spec = zero;
    if ((PH == LP_LEAF || PH == LP_TAN || PH == LP_V || PH == LP_C) && s.free_head != (int)GT_NIL) spec = lf_pair(f, 2 * s.free_head);

    if (PH == LP_LEAF) {
        // reference :516-531 for sorted points a < b (< c); every orientation it asks is the sign of (a, b, c)
        const int a = s.ia, b = s.ia + 1, c = s.ib;
        const Pt pa = lf_pt(f, a), pb = lf_pt(f, b), pc = lf_pt(f, c);
        int sg = 0;
        if (c != b) { sg = lp_orient2d<AP>(pa.x, pa.y, pb.x, pb.y, pc.x, pc.y); if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; } }
        auto take = [&](LPair sp) -> int {
            if (s.free_head != (int)GT_NIL) { const int e = s.free_head; s.free_head = sp.a.nxt; return e; }
            const int e = lp_pool_take(f);
            if (e < 0) s.err |= GT_ERR_POOL;
            return e;
        };
        const int e0 = take(spec);
        if (e0 < 0) return;
        const int h0 = 2 * e0;
        if (c == b) {                                     // two points: one edge, two single-entry rings
            lf_set_rec(f, h0, b, h0, h0); lf_set_rec(f, h0 ^ 1, a, h0 ^ 1, h0 ^ 1);
            lf_set_head(f, a, h0); lf_set_head(f, b, h0 ^ 1);
        } else {
        // further edges: the free list's links are read directly here (leaves only)
        const int e1 = take(s.free_head != (int)GT_NIL ? lf_pair(f, 2 * s.free_head) : zero);
        if (e1 < 0) return;
        const int h1 = 2 * e1;
        const bool s1 = sg > 0, s2 = sg < 0;              // ccw(a,b,c), ccw(a,c,b)
        int h2 = -1;
        if (s1 || s2) {
            const int e2 = take(s.free_head != (int)GT_NIL ? lf_pair(f, 2 * s.free_head) : zero);
            if (e2 < 0) return;
            h2 = 2 * e2;
        }
        // ring of b: {b->a, b->c}; "first" moves to b->c iff c is right of the ray b->a (= ccw(a,b,c))
        lf_set_rec(f, h0 ^ 1, a, h1, h1); lf_set_rec(f, h1, c, h0 ^ 1, h0 ^ 1);
        lf_set_head(f, b, s1 ? h1 : (h0 ^ 1));
        if (h2 < 0) {                                     // collinear: a and c keep single-entry rings
            lf_set_rec(f, h0, b, h0, h0); lf_set_head(f, a, h0);
            lf_set_rec(f, h1 ^ 1, b, h1 ^ 1, h1 ^ 1); lf_set_head(f, c, h1 ^ 1);
        } else {
            // ring of a: {a->b, a->c}; first moves to a->c iff c is right of a->b (= ccw(a,c,b))
            lf_set_rec(f, h0, b, h2, h2); lf_set_rec(f, h2, c, h0, h0);
            lf_set_head(f, a, s2 ? h2 : h0);
            // ring of c: {c->b, c->a}; first moves to c->a iff a is right of c->b (= ccw(a,b,c))
            lf_set_rec(f, h1 ^ 1, b, h2 ^ 1, h2 ^ 1); lf_set_rec(f, h2 ^ 1, a, h1 ^ 1, h1 ^ 1);
            lf_set_head(f, c, s1 ? (h2 ^ 1) : (h1 ^ 1));
        }
        }
        lp_finish_node(f, s);
        return;
    }

    if (PH == LP_TINIT) {
        if (s.adv) { s.adv = 0; lp_finish_node(f, s); if (s.phase != LP_TINIT) return; }      // (see lp_finish_connect)
        // both tangents start at the rightmost vertex of the left half and the leftmost of the right half
        const int x = (s.ia + s.ib) / 2, y = x + 1;
        int hdx = 0, hdy = 0;
        const Pt qx = lf_pth(f, x, hdx), qy = lf_pth(f, y, hdy);
        const LPair px = lf_pair(f, hdx), py = lf_pair(f, hdy);
        s.ux = x; s.uy = y; s.uHl = hdx; s.uHr = py.a.prv; s.pr1 = qx; s.pl1 = qy; s.uSub = 0;      // upper: x -> first(x), y -> pred(y, first(y))   (:378-381)
        s.li = x; s.ri = y; s.lHl = px.a.prv; s.lHr = hdy; s.pli = qx; s.pri = qy; s.lSub = 0;      // lower: x -> pred(x, first(x)), y -> first(y)   (:342-345)
        s.headL = hdx; s.headR = hdy;
        s.hr2 = 0;                                         // (tangent phase: iteration guard, see LP_TAN)
        s.phase = LP_TAN;
        return;
    }

    if (PH == LP_TAN) {
        // chain 0: upper tangent (uct :373-406), leftOf(p, x, y) = ccw(p, x, y); chain 1: lower (lct :337-370), rightOf(p, x, y) = ccw(p, y, x)
        // a hull walk visits every vertex at most once per side: more iterations than that means inconsistent predicate
        // answers (only after a range error) -- stop instead of circling forever
        if (++s.hr2 > 2 * f.n + 16) { s.err |= GT_ERR_RING; return; }
        const bool a0 = s.uSub < 2, a1 = s.lSub < 2;
        LPair p0 = zero, p1 = zero; Pt c0, c1; c0.x = c0.y = 0.0; c1 = c0; int hd1c = 0;
        if (a0) p0 = lf_pair(f, s.uSub == 0 ? s.uHr : s.uHl);
        if (a1) p1 = lf_pair(f, s.lSub == 0 ? s.lHr : s.lHl);
        if (a0) c0 = lf_pt(f, p0.a.dst);
        if (a1) c1 = lf_pth(f, p1.a.dst, hd1c);
        if (a0) {
            int sg = lp_orient2d<AP>(c0.x, c0.y, s.pr1.x, s.pr1.y, s.pl1.x, s.pl1.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            if (s.uSub == 0) { if (sg > 0) { s.uHr = p0.b.prv; s.uy = p0.a.dst; s.pl1 = c0; } else s.uSub = 1; }        // pred(rfast, y)
            else if (sg > 0) { s.uHl = p0.b.nxt; s.ux = p0.a.dst; s.pr1 = c0; s.uSub = 0; }                           // succ(lfast, x)
            else s.uSub = 2;
        }
        if (a1) {
            int sg = lp_orient2d<AP>(c1.x, c1.y, s.pri.x, s.pri.y, s.pli.x, s.pli.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            if (s.lSub == 0) { if (sg > 0) { s.lHr = p1.b.nxt; s.ri = p1.a.dst; s.pri = c1; s.headR = hd1c; } else s.lSub = 1; }   // succ(rfast, y)
            else if (sg > 0) { s.lHl = p1.b.prv; s.li = p1.a.dst; s.pli = c1; s.headL = hd1c; s.lSub = 0; }                   // pred(lfast, x)
            else s.lSub = 2;
        }
        if (s.uSub == 2 && s.lSub == 2) {
            // first cross edge = lower tangent: both ends take it in their hull gap (before "first")
            s.first_edge = 1;
            lp_new_edge(f, s, spec, false, 0);
        }
        return;
    }

    if (PH == LP_INS || PH == LP_SCAN) {
        // (the two phases were also tried as one function with a run-time switch -- 5 KB less code, one chunk per round
        // less: 308 ms instead of 277 ms per 100k frames, the scans ride along at the insertion's three load levels)
        constexpr bool ins = PH == LP_INS;
        // endpoint 0: parent li, in ri, half-edge h0; endpoint 1: parent ri, in li, half-edge h0^1.
        // rightOf(in, parent, q) = ccw(in, q, parent)  (insertNode :216-224)
        const bool a0 = ins || s.scan0, a1 = ins || s.scan1;
        LPair p0 = zero, p1 = zero; Pt c0, c1; c0.x = c0.y = 0.0; c1 = c0;
        if (a0) p0 = lf_pair(f, ins ? s.headL : s.cur0);
        if (a1) p1 = lf_pair(f, ins ? s.headR : s.cur1);
        if (a0) c0 = lf_pt(f, p0.a.dst);
        if (a1) c1 = lf_pt(f, p1.a.dst);
        // LP_INS also fetches the ring's LAST entry (prv of first): when the new neighbour is right of "first" the
        // reference's clockwise scan starts there (:217-219) and usually stops there, so its first test is taken in
        // the same iteration and the scan only goes on when that test holds too
        LPair q0 = zero, q1 = zero; Pt d0 = c0, d1 = c1;
        const bool m0 = ins && (int)p0.a.prv != s.headL, m1 = ins && (int)p1.a.prv != s.headR;
        if (m0) q0 = lf_pair(f, p0.a.prv);
        if (m1) q1 = lf_pair(f, p1.a.prv);
        if (m0) d0 = lf_pt(f, q0.a.dst);
        if (m1) d1 = lf_pt(f, q1.a.dst);
        // Each endpoint: decide (at, nx, newhead) -- or that the scan goes on -- and place once. One copy of the placement
        // per endpoint and both orientation filters evaluated for every lane keep the step short and its lanes together
        // (it is the longest step of a round: the round lasts as long as this).
        if (a0) {
            int sg = lp_orient2d<AP>(s.pri.x, s.pri.y, c0.x, c0.y, s.pli.x, s.pli.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            const bool res = sg > 0;
            bool place = true, newhead = false; int at, nx, hd;
            if (ins) {
                hd = s.headL; const int last = p0.a.prv;
                int sl = m0 ? lp_orient2d<AP>(s.pri.x, s.pri.y, d0.x, d0.y, s.pli.x, s.pli.y) : 0;   // right of "last" too?
                if (sl == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sl = 0; }
                at = last; nx = hd;                                                  // after "last" (= before "first")
                if (!res) { if (!s.first_edge) { at = s.at0; nx = s.nx0; } }          // not right of "first": the slot known from the merge
                else if (last == hd) { at = hd; newhead = true; }                     // single-entry ring: right of all
                else if (!(sl > 0)) { }                                               // not right of "last": goes after it
                else if ((int)q0.a.prv == hd) newhead = true;                         // right of all: new "first"
                else { place = false; s.cur0 = q0.a.prv; s.hd0 = hd; s.last0 = last; s.scan0 = 1; }
            } else {
                hd = s.hd0; at = s.cur0; nx = p0.a.nxt;
                if (res) {
                    const int cur = p0.a.prv;
                    if (cur == hd) { at = s.last0; nx = hd; newhead = true; }
                    else { place = false; s.cur0 = cur; }
                }
                if (place) s.scan0 = 0;
            }
            if (place) lp_place(f, s, 0, s.li, s.ri, s.h0, at, nx, newhead, hd);
        }
        if (a1) {
            int sg = lp_orient2d<AP>(s.pli.x, s.pli.y, c1.x, c1.y, s.pri.x, s.pri.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            const bool res = sg > 0;
            const int h1 = s.h0 ^ 1;
            bool place = true, newhead = false; int at, nx, hd;
            if (ins) {
                hd = s.headR; const int last = p1.a.prv;
                int sl = m1 ? lp_orient2d<AP>(s.pli.x, s.pli.y, d1.x, d1.y, s.pri.x, s.pri.y) : 0;
                if (sl == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sl = 0; }
                at = last; nx = hd;
                if (!res) { if (!s.first_edge) { at = s.at1; nx = s.nx1; } }
                else if (last == hd) { at = hd; newhead = true; }
                else if (!(sl > 0)) { }
                else if ((int)q1.a.prv == hd) newhead = true;
                else { place = false; s.cur1 = q1.a.prv; s.hd1 = hd; s.last1 = last; s.scan1 = 1; }
            } else {
                hd = s.hd1; at = s.cur1; nx = p1.a.nxt;
                if (res) {
                    const int cur = p1.a.prv;
                    if (cur == hd) { at = s.last1; nx = hd; newhead = true; }
                    else { place = false; s.cur1 = cur; }
                }
                if (place) s.scan1 = 0;
            }
            if (place) lp_place(f, s, 1, s.ri, s.li, h1, at, nx, newhead, hd);
        }
        if (s.scan0 || s.scan1) s.phase = LP_SCAN;
        else { s.first_edge = 0; lp_finish_connect(f, s); }
        return;
    }

    if (PH == LP_V) {
        // chain 0: leftOf(r1, li, ri) with r1 = pred(ri, li) (:555-556); chain 1: rightOf(l1, ri, li) with l1 = succ(li, ri) (:568-569)
        const LPair p0 = lf_pair(f, s.hr1), p1 = lf_pair(f, s.hl1);
        int hdr = 0, hdl = 0;
        const Pt c0 = lf_pth(f, p0.a.dst, hdr), c1 = lf_pth(f, p1.a.dst, hdl);
        s.hdR1 = hdr; s.hdL1 = hdl;
        int sg0 = lp_orient2d<AP>(c0.x, c0.y, s.pli.x, s.pli.y, s.pri.x, s.pri.y);
        int sg1 = lp_orient2d<AP>(c1.x, c1.y, s.pli.x, s.pli.y, s.pri.x, s.pri.y);
        if (sg0 == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg0 = 0; }
        if (sg1 == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg1 = 0; }
        s.r1 = p0.a.dst; s.pr1 = c0; s.hr2 = p0.a.prv; s.t1_prv = p0.b.prv; s.t1_nxt = p0.b.nxt;
        s.l1 = p1.a.dst; s.pl1 = c1; s.hl2 = p1.a.nxt; s.u1_prv = p1.b.prv; s.u1_nxt = p1.b.nxt;
        s.vR = sg0 > 0; s.vL = sg1 > 0; s.actR = s.vR; s.actL = s.vL;
        if (s.actR || s.actL) s.phase = LP_C; else lp_decide<AP>(f, s, spec, false, 0);
        return;
    }

    if (PH == LP_C) {
        // chain 0: while incircle(r1, li, ri, r2): cutVerts(ri, r1) (:559-567); chain 1: while incircle(li, ri, l1, l2): cutVerts(li, l1) (:572-580)
        const bool a0 = s.actR, a1 = s.actL;
        LPair p0 = zero, p1 = zero; Pt c0, c1; c0.x = c0.y = 0.0; c1 = c0; int hn0 = 0, hn1 = 0;
        if (a0) p0 = lf_pair(f, s.hr2);
        if (a1) p1 = lf_pair(f, s.hl2);
        if (a0) c0 = lf_pth(f, p0.a.dst, hn0);
        if (a1) c1 = lf_pth(f, p1.a.dst, hn1);
        bool freed = false; int prev_head = 0;
        if (a0) {
            const int r2 = p0.a.dst;
            int sg = 0;
            if (!(r2 == s.li || r2 == s.ri || r2 == s.r1))                          // a repeated point: exactly 0
                sg = lp_incircle<AP>(s.pr1.x, s.pr1.y, s.pli.x, s.pli.y, s.pri.x, s.pri.y, c0.x, c0.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            if (sg > 0) {
                // ri's ring loses hr1 (its next is the cross edge), r1's ring loses the twin
                const int h = s.hr1, tw = h ^ 1;
                lf_set_nxt(f, s.hr2, s.e ^ 1); lf_set_prv(f, s.e ^ 1, s.hr2);
                if (s.headR == h) { s.headR = s.e ^ 1; lf_set_head(f, s.ri, s.e ^ 1); }
                lf_set_nxt(f, s.t1_prv, s.t1_nxt); lf_set_prv(f, s.t1_nxt, s.t1_prv);
                if (s.hdR1 == tw) lf_set_head(f, s.r1, (s.t1_nxt == tw) ? GT_NIL : s.t1_nxt);
                prev_head = s.free_head; freed = true; lp_free(f, s, h >> 1);
                s.hr1 = s.hr2; s.r1 = r2; s.pr1 = c0; s.hr2 = p0.a.prv; s.t1_prv = p0.b.prv; s.t1_nxt = p0.b.nxt; s.hdR1 = hn0;
            } else s.actR = 0;
        }
        if (a1) {
            const int l2 = p1.a.dst;
            int sg = 0;
            if (!(l2 == s.li || l2 == s.ri || l2 == s.l1))
                sg = lp_incircle<AP>(s.pli.x, s.pli.y, s.pri.x, s.pri.y, s.pl1.x, s.pl1.y, c1.x, c1.y);
            if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
            if (sg > 0) {
                const int h = s.hl1, tw = h ^ 1;
                lf_set_nxt(f, s.e, s.hl2); lf_set_prv(f, s.hl2, s.e);
                if (s.headL == h) { s.headL = s.hl2; lf_set_head(f, s.li, s.hl2); }
                lf_set_nxt(f, s.u1_prv, s.u1_nxt); lf_set_prv(f, s.u1_nxt, s.u1_prv);
                if (s.hdL1 == tw) lf_set_head(f, s.l1, (s.u1_nxt == tw) ? GT_NIL : s.u1_nxt);
                prev_head = s.free_head; freed = true; lp_free(f, s, h >> 1);
                s.hl1 = s.hl2; s.l1 = l2; s.pl1 = c1; s.hl2 = p1.a.nxt; s.u1_prv = p1.b.prv; s.u1_nxt = p1.b.nxt; s.hdL1 = hn1;
            } else s.actL = 0;
        }
        if (!(s.actR || s.actL)) lp_decide<AP>(f, s, spec, freed, prev_head);
        return;
    }
}

// Generic entry (host emulation): dispatch on the frame's phase. Returns false when the tree is finished.
GT_INLINE bool lpf_step(const LFrame& f, LState& s) {
    switch (s.phase) {
    case LP_LEAF: lpf_step_t<LP_LEAF>(f, s); break;
    case LP_TINIT: lpf_step_t<LP_TINIT>(f, s); break;
    case LP_TAN: lpf_step_t<LP_TAN>(f, s); break;
    case LP_INS: lpf_step_t<LP_INS>(f, s); break;
    case LP_SCAN: lpf_step_t<LP_SCAN>(f, s); break;
    case LP_V: lpf_step_t<LP_V>(f, s); break;
    case LP_C: lpf_step_t<LP_C>(f, s); break;
    default: return false;
    }
    return s.phase != LP_DONE;
}

// Triangles owned by sorted vertex i, in the reference's emission order (convertTrisFreeAdj :606-635);
// same rule as dt_core.cuh::vertex_triangles on the packed records.
template <class Visit>
GT_INLINE int lf_vertex_triangles(const LFrame& f, int i, Visit&& visit) {
    const int hd = lf_head(f, i);
    if (hd == (int)GT_NIL || f.rec[hd].nxt == hd) return 0;
    int cnt = 0, h = hd;
    do {
        const int hn = f.rec[h].nxt;
        const int n1 = f.rec[h].dst, n2 = f.rec[hn].dst;
        if (n1 > i && n2 > i) {
            if (hn == hd) {                                 // wrap-around pair: stop at the hull gap (:619-621)
                const Pt p1 = lf_pt(f, n1), pi = lf_pt(f, i), p2 = lf_pt(f, n2);
                if (!(orient2d_sign(p1.x, p1.y, p2.x, p2.y, pi.x, pi.y) > 0)) break;     // !rightOf(n1, i, n2) = !ccw(n1, n2, i)
            }
            visit(n1, n2);
            ++cnt;
        }
        h = hn;
    } while (h != hd);
    return cnt;
}

// The same enumeration for the area pass, arranged for memory-level parallelism: the ring is one chain of dependent
// 8-byte record loads (the record of the next entry is fetched before the current triangle is evaluated), every
// neighbour's point record is loaded once and handed from one triangle to the next. pi = the vertex's own record
// (coordinates + ring head). visit(n1, n2, p1, p2) in emission order; returns the count.
template <class Visit>
GT_INLINE int lf_vertex_triangles_pts(const LFrame& f, int i, const LPt& pi, Visit&& visit) {
    const int hd = pi.head;
    if (hd == (int)GT_NIL) return 0;
    LRec cur = f.rec[hd];
    if (cur.nxt == hd) return 0;
    LPt none; none.x = none.y = none.z = 0.0; none.head = none.perm = 0; none.pad = 0;
    LPt p1 = (int)cur.dst > i ? f.pt[cur.dst] : none;
    int cnt = 0, h = hd;
    do {
        const int hn = cur.nxt;
        const LRec nx = f.rec[hn];
        const int n1 = cur.dst, n2 = nx.dst;
        const LPt p2 = n2 > i ? f.pt[n2] : none;
        if (n1 > i && n2 > i) {
            if (hn == hd && !(orient2d_sign(p1.x, p1.y, p2.x, p2.y, pi.x, pi.y) > 0)) break;      // the hull gap (:619-621): !rightOf(n1, i, n2)
            visit(n1, n2, p1, p2);
            ++cnt;
        }
        h = hn; cur = nx; p1 = p2;
    } while (h != hd);
    return cnt;
}

}  // namespace gt
