// lpf_core.cuh -- the Delaunay divide-and-conquer as a per-frame state machine ("lane per frame").
//
// dt_core.cuh::merge is a nest of data-dependent loops: when the lanes of a warp walk different merges
// they diverge and the warp runs them one after another. Here the same algorithm -- the same predicates
// on the same operands in the same order, hence the same rings as the reference (src/delaunay_tri.c:
// 511-601, tangents :337-406, insertNode :210-245, deleteNode :247-268) -- is cut into ITERATIONS.
// One iteration of one frame is
//     [A: ring entry of a vertex] -> [B: one edge = its two half-edge records] -> [C: one point]
//     -> ONE predicate (orient2d or incircle) -> a short transition that edits the rings and names
//     the phase of the next iteration.
// Each phase is a template instance lpf_step_t<PH>: straight-line code. The GPU scheduler
// (gta_lpf.cuh) sorts the frames in flight by phase every round and lets a warp run ONE phase for 32
// different frames, so there is no divergence between lanes and their (uncoalesced) loads are in flight
// together. Per frame the persistent state is 36 words (ls_pack / ls_unpack); the frame itself lives in
// global memory:
//   point  16 B {x, y}   record 8 B {dst, nxt, prv, -} per half-edge, twins adjacent (edge = 16 B, one load)
//   head    2 B per vertex
// Edge pool, free list and the cursor over the recursion tree are private to the frame: no atomics.
//
// The two tangents are computed upper first, then lower (the reference does lower first, :542-543; both are
// read-only walks over the untouched halves, so the order cannot matter): the lower tangent's end points are
// then already in the registers that the first merge step needs.
//
// Host-compilable: tests/emu drives lpf_step() sequentially and compares with the oracle.
#pragma once
#include "dt_core.cuh"

namespace gt {

struct LRec { u16 dst, nxt, prv, pad; };                 // one half-edge
struct LPair { LRec a, b; };                             // a = the half-edge asked for, b = its twin

// Frame storage in global memory.
struct LFrame {
    Pt* pt;         // [n] sorted (x, y)
    LRec* rec;      // [2 * cap_e]
    u16* head;      // [n]
    int n;
    int cap_e;
};

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
#else
GT_INLINE LPair lf_pair(const LFrame& f, int h) { LPair p; p.a = f.rec[h]; p.b = f.rec[h ^ 1]; return p; }
GT_INLINE Pt lf_pt(const LFrame& f, int v) { return f.pt[v]; }
#endif
GT_INLINE void lf_set_nxt(const LFrame& f, int h, int v) { f.rec[h].nxt = (u16)v; }
GT_INLINE void lf_set_prv(const LFrame& f, int h, int v) { f.rec[h].prv = (u16)v; }
GT_INLINE void lf_set_rec(const LFrame& f, int h, int dst, int nxt, int prv) {
    LRec r; r.dst = (u16)dst; r.nxt = (u16)nxt; r.prv = (u16)prv; r.pad = 0; f.rec[h] = r;
}

enum LPhase : int {
    LP_LEAF0 = 0, LP_LEAF1, LP_LEAF2,
    LP_TINIT0, LP_TINIT1, LP_TAN_R, LP_TAN_L,
    LP_INS_T0, LP_INS_SCAN,
    LP_VR, LP_CR, LP_VL, LP_CL, LP_F,
    LP_DONE,          // whole tree finished
    LP_NPHASE
};

// Per-frame state between iterations.
struct LState {
    int phase;
    int d, k, ia, ib;                 // current node of the recursion tree
    uint32_t err;
    int bump, free_head;              // private edge pool
    int up, first_edge, vR, vL, iGap, i1Gap, iWhich;
    // tangents: current end points tx / ty (their coordinates are pli / pri), walk positions hl / hr
    int tx, ty, hl, hr, ux, uy;
    // merge step
    int li, ri, e, headL, headR;
    int hr1, r1, hr2, t1_prv, t1_nxt;    // right candidate ri->r1, its clockwise neighbour, the twin's links at r1
    int hl1, l1, hl2, u1_prv, u1_nxt;    // left candidate li->l1, its counter-clockwise neighbour, the twin's links at l1
    // ring insertion in flight (endpoint 0 = the left vertex li, endpoint 1 = the right vertex ri)
    int iP, iIN, iH, iAt, iNx, iCur, iHd, iLast, at0, nx0;
    int i1H, i1At, i1Nx;
    // cached coordinates: pli / pri = li / ri (tangents: tx / ty), pr1 / pl1 = the two candidates (leaves: points a, b)
    Pt pli, pri, pr1, pl1;
};

// ---------------------------------------------------------------- persistent form (36 words)
constexpr int LS_WORDS = 36;
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
    put(0, (uint32_t)s.phase | ((uint32_t)s.d << 4) | ((uint32_t)s.k << 8) | ((s.err & 15u) << 20) | ((uint32_t)s.up << 24) | ((uint32_t)s.first_edge << 25) |
           ((uint32_t)s.vR << 26) | ((uint32_t)s.vL << 27) | ((uint32_t)s.iGap << 28) | ((uint32_t)s.i1Gap << 29) | ((uint32_t)s.iWhich << 30));
    put(1, LS_P2(s.ia, s.ib)); put(2, LS_P2(s.bump, s.free_head));
    put(3, LS_P2(s.tx, s.ty)); put(4, LS_P2(s.hl, s.hr)); put(5, LS_P2(s.ux, s.uy));
    put(6, LS_P2(s.li, s.ri)); put(7, LS_P2(s.e, s.headL)); put(8, LS_P2(s.headR, s.hr1)); put(9, LS_P2(s.r1, s.hr2));
    put(10, LS_P2(s.t1_prv, s.t1_nxt)); put(11, LS_P2(s.hl1, s.l1)); put(12, LS_P2(s.hl2, s.u1_prv)); put(13, LS_P2(s.u1_nxt, s.iP));
    put(14, LS_P2(s.iIN, s.iH)); put(15, LS_P2(s.iAt, s.iNx)); put(16, LS_P2(s.iCur, s.iHd)); put(17, LS_P2(s.iLast, s.at0));
    put(18, LS_P2(s.nx0, s.i1H)); put(19, LS_P2(s.i1At, s.i1Nx));
    put(20, ls_lo(s.pli.x)); put(21, ls_hi(s.pli.x)); put(22, ls_lo(s.pli.y)); put(23, ls_hi(s.pli.y));
    put(24, ls_lo(s.pri.x)); put(25, ls_hi(s.pri.x)); put(26, ls_lo(s.pri.y)); put(27, ls_hi(s.pri.y));
    put(28, ls_lo(s.pr1.x)); put(29, ls_hi(s.pr1.x)); put(30, ls_lo(s.pr1.y)); put(31, ls_hi(s.pr1.y));
    put(32, ls_lo(s.pl1.x)); put(33, ls_hi(s.pl1.x)); put(34, ls_lo(s.pl1.y)); put(35, ls_hi(s.pl1.y));
}
template <class Get>
GT_INLINE void ls_unpack(LState& s, Get&& get) {
    uint32_t w = get(0);
    s.phase = w & 15; s.d = (w >> 4) & 15; s.k = (w >> 8) & 0xfff; s.err = (w >> 20) & 15; s.up = (w >> 24) & 1; s.first_edge = (w >> 25) & 1;
    s.vR = (w >> 26) & 1; s.vL = (w >> 27) & 1; s.iGap = (w >> 28) & 1; s.i1Gap = (w >> 29) & 1; s.iWhich = (w >> 30) & 1;
#define LS_U2(i, a, b) w = get(i); a = (int)(w & 0xffff); b = (int)(w >> 16)
    LS_U2(1, s.ia, s.ib); LS_U2(2, s.bump, s.free_head);
    LS_U2(3, s.tx, s.ty); LS_U2(4, s.hl, s.hr); LS_U2(5, s.ux, s.uy);
    LS_U2(6, s.li, s.ri); LS_U2(7, s.e, s.headL); LS_U2(8, s.headR, s.hr1); LS_U2(9, s.r1, s.hr2);
    LS_U2(10, s.t1_prv, s.t1_nxt); LS_U2(11, s.hl1, s.l1); LS_U2(12, s.hl2, s.u1_prv); LS_U2(13, s.u1_nxt, s.iP);
    LS_U2(14, s.iIN, s.iH); LS_U2(15, s.iAt, s.iNx); LS_U2(16, s.iCur, s.iHd); LS_U2(17, s.iLast, s.at0);
    LS_U2(18, s.nx0, s.i1H); LS_U2(19, s.i1At, s.i1Nx);
#undef LS_U2
    s.pli.x = ls_dbl(get(20), get(21)); s.pli.y = ls_dbl(get(22), get(23));
    s.pri.x = ls_dbl(get(24), get(25)); s.pri.y = ls_dbl(get(26), get(27));
    s.pr1.x = ls_dbl(get(28), get(29)); s.pr1.y = ls_dbl(get(30), get(31));
    s.pl1.x = ls_dbl(get(32), get(33)); s.pl1.y = ls_dbl(get(34), get(35));
}

// Words a phase may modify (a word is rewritten from the unpacked fields, so a phase that changes one field of
// a word keeps the other one). tests/emu applies these masks after every iteration: a missing bit shows up
// as a triangulation mismatch.
#define LS_W(i) (1ull << (i))
#define LS_PT(i) (15ull << (i))
constexpr unsigned long long LS_M_NODE = LS_W(0) | LS_W(1) | LS_W(3);
template <int PH> struct LsDirty { static constexpr unsigned long long mask = ~0ull; };
template <> struct LsDirty<LP_LEAF0> { static constexpr unsigned long long mask = LS_W(0) | LS_PT(28); };
template <> struct LsDirty<LP_LEAF1> { static constexpr unsigned long long mask = LS_W(0) | LS_PT(32); };
template <> struct LsDirty<LP_LEAF2> { static constexpr unsigned long long mask = LS_W(2) | LS_M_NODE; };
template <> struct LsDirty<LP_TINIT0> { static constexpr unsigned long long mask = LS_W(0) | LS_W(4) | LS_PT(20); };
template <> struct LsDirty<LP_TINIT1> { static constexpr unsigned long long mask = LS_W(0) | LS_W(4) | LS_PT(24); };
template <> struct LsDirty<LP_TAN_R> { static constexpr unsigned long long mask = LS_W(0) | LS_W(3) | LS_W(4) | LS_PT(24); };
template <> struct LsDirty<LP_TAN_L> { static constexpr unsigned long long mask = LS_W(0) | LS_W(3) | LS_W(4) | LS_W(5) | LS_W(6) | LS_PT(20); };
template <> struct LsDirty<LP_INS_T0> { static constexpr unsigned long long mask = LS_M_NODE | LS_W(7) | LS_W(8) | LS_W(11) | LS_W(13) | LS_W(14) | LS_W(15) | LS_W(16) | LS_W(17) | LS_W(18); };
template <> struct LsDirty<LP_INS_SCAN> { static constexpr unsigned long long mask = LsDirty<LP_INS_T0>::mask; };
template <> struct LsDirty<LP_VR> { static constexpr unsigned long long mask = LS_W(0) | LS_W(9) | LS_W(10) | LS_PT(28); };
template <> struct LsDirty<LP_CR> { static constexpr unsigned long long mask = LS_W(0) | LS_W(2) | LS_W(8) | LS_W(9) | LS_W(10) | LS_PT(28); };
template <> struct LsDirty<LP_VL> { static constexpr unsigned long long mask = LS_W(0) | LS_W(11) | LS_W(12) | LS_W(13) | LS_PT(32); };
template <> struct LsDirty<LP_CL> { static constexpr unsigned long long mask = LS_W(0) | LS_W(2) | LS_W(7) | LS_W(11) | LS_W(12) | LS_W(13) | LS_PT(32); };
template <> struct LsDirty<LP_F> { static constexpr unsigned long long mask = LS_W(0) | LS_W(2) | LS_W(6) | LS_W(13) | LS_W(14) | LS_W(15) | LS_W(18) | LS_W(19) | LS_PT(20) | LS_PT(24); };
template <unsigned long long MASK, class Put>
GT_INLINE void ls_pack_mask(const LState& s, Put&& put) {
    ls_pack(s, [&](int i, uint32_t v) { if ((MASK >> i) & 1ull) put(i, v); });
}

// ---------------------------------------------------------------- helpers
GT_INLINE int lp_alloc(const LFrame& f, LState& s, const LPair& freerec) {
    if (s.free_head != (int)GT_NIL) { int e = s.free_head; s.free_head = freerec.a.nxt; return e; }
    if (s.bump >= f.cap_e) { s.err |= GT_ERR_POOL; return -1; }
    return s.bump++;
}
GT_INLINE void lp_free(const LFrame& f, LState& s, int e) { lf_set_nxt(f, 2 * e, s.free_head); s.free_head = e; }

// queue the two ring insertions of the new edge h0 = li -> ri (endpoint 0 at li, endpoint 1 at ri)
GT_INLINE void lp_queue_inserts(LState& s, int h0, int at0, int nx0, int gap0, int at1, int nx1, int gap1) {
    s.iP = s.li; s.iIN = s.ri; s.iH = h0; s.iAt = at0; s.iNx = nx0; s.iGap = gap0; s.iWhich = 0;
    s.i1H = h0 ^ 1; s.i1At = at1; s.i1Nx = nx1; s.i1Gap = gap1;
    s.phase = LP_INS_T0;
}

// Advance the post-order cursor over the reference's recursion tree to the next node and enter its first phase.
GT_INLINE void lp_next_node(const LFrame& f, LState& s) {
    int d = s.d, k = s.k;
    if (d < 0) { d = 0; k = 0; }                           // first call: descend from the root
    else if (d == 0) { s.phase = LP_DONE; return; }
    else if (k & 1) {                                      // finished a right child: its parent is next, and it is a merge
        d -= 1; k >>= 1;
        int ia, ib; tree_node(f.n, d, k, &ia, &ib);
        s.d = d; s.k = k; s.ia = ia; s.ib = ib;
        s.phase = LP_TINIT0; s.up = 1; s.tx = (ia + ib) / 2; s.ty = s.tx + 1;
        return;
    } else k += 1;                                         // finished a left child: leftmost leaf of the sibling subtree
    int ia, ib;
    for (;;) {
        tree_node(f.n, d, k, &ia, &ib);
        if (ib - ia < 3) break;
        d += 1; k <<= 1;
    }
    s.d = d; s.k = k; s.ia = ia; s.ib = ib;
    s.phase = (ib - ia < 1) ? LP_DONE : LP_LEAF0;
}

GT_INLINE void lp_begin(const LFrame& f, LState& s) {
    s.d = -1; s.k = 0; s.err = 0; s.bump = 0; s.free_head = GT_NIL; s.first_edge = 0; s.up = 0; s.vR = s.vL = 0;
    s.iGap = s.i1Gap = s.iWhich = 0;
    lp_next_node(f, s);
}

// ---------------------------------------------------------------- one iteration of phase PH
template <int PH>
GT_INLINE void lpf_step_t(const LFrame& f, LState& s) {
    // ---- what to fetch
    int fa = -1, fb = -1, fc = -1; bool bA = false, cB = false;
    const int freeh = (s.free_head != (int)GT_NIL) ? 2 * s.free_head : -1;
    if (PH == LP_LEAF0) fc = s.ia;
    else if (PH == LP_LEAF1) fc = s.ia + 1;
    else if (PH == LP_LEAF2) { fc = s.ib; fb = freeh; }
    else if (PH == LP_TINIT0) { fa = s.tx; bA = true; fc = s.tx; }
    else if (PH == LP_TINIT1) { fa = s.ty; bA = true; fc = s.ty; }
    else if (PH == LP_TAN_R) { fb = s.hr; cB = true; }
    else if (PH == LP_TAN_L) { fb = s.hl; cB = true; }
    else if (PH == LP_INS_T0) { fa = s.iP; bA = true; cB = true; }
    else if (PH == LP_INS_SCAN) { fb = s.iCur; cB = true; }
    else if (PH == LP_VR) { fb = s.hr1; cB = true; }
    else if (PH == LP_VL) { fb = s.hl1; cB = true; }
    else if (PH == LP_CR) { fa = s.r1; fb = s.hr2; cB = true; }
    else if (PH == LP_CL) { fa = s.l1; fb = s.hl2; cB = true; }
    else if (PH == LP_F) fb = freeh;
    int hdA = 0;
    if (fa >= 0) hdA = f.head[fa];
    const int hb = bA ? hdA : fb;
    LPair pr; pr.a.dst = pr.a.nxt = pr.a.prv = pr.a.pad = 0; pr.b = pr.a;
    if (hb >= 0) pr = lf_pair(f, hb);
    const int vc = cB ? (int)pr.a.dst : fc;
    Pt pc; pc.x = pc.y = 0.0;
    if (vc >= 0) pc = lf_pt(f, vc);

    // ---- the iteration's predicate
    int sg = 0;
    if (PH == LP_LEAF2) { if (s.ib - s.ia == 2) sg = orient2d_sign(s.pr1.x, s.pr1.y, s.pl1.x, s.pl1.y, pc.x, pc.y); }
    else if (PH == LP_TAN_R || PH == LP_TAN_L) {
        // lower: rightOf(p, x, y) = ccw(p, y, x); upper: leftOf(p, x, y) = ccw(p, x, y)   (x = tx at pli, y = ty at pri)
        const Pt A = s.up ? s.pli : s.pri, B = s.up ? s.pri : s.pli;
        sg = orient2d_sign(pc.x, pc.y, A.x, A.y, B.x, B.y);
    } else if (PH == LP_INS_T0 || PH == LP_INS_SCAN) {
        // rightOf(in, parent, q) = ccw(in, q, parent); endpoint 0: parent li, in ri; endpoint 1: the other way round
        const Pt pP = s.iWhich ? s.pri : s.pli, pIN = s.iWhich ? s.pli : s.pri;
        sg = orient2d_sign(pIN.x, pIN.y, pc.x, pc.y, pP.x, pP.y);
    } else if (PH == LP_VR || PH == LP_VL) {
        sg = orient2d_sign(pc.x, pc.y, s.pli.x, s.pli.y, s.pri.x, s.pri.y);       // leftOf(r1, li, ri) / rightOf(l1, ri, li)
    } else if (PH == LP_CR) {
        if (!(vc == s.li || vc == s.ri || vc == s.r1))                              // a repeated point: exactly 0
            sg = incircle_sign(s.pr1.x, s.pr1.y, s.pli.x, s.pli.y, s.pri.x, s.pri.y, pc.x, pc.y);
    } else if (PH == LP_CL) {
        if (!(vc == s.li || vc == s.ri || vc == s.l1))
            sg = incircle_sign(s.pli.x, s.pli.y, s.pri.x, s.pri.y, s.pl1.x, s.pl1.y, pc.x, pc.y);
    } else if (PH == LP_F) {
        if (s.vR && s.vL && !s.first_edge && s.r1 != s.l1)
            sg = incircle_sign(s.pli.x, s.pli.y, s.pri.x, s.pri.y, s.pr1.x, s.pr1.y, s.pl1.x, s.pl1.y);
    }
    if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
    const bool res = sg > 0;

    // ---- transition
    if (PH == LP_LEAF0) { s.pr1 = pc; s.phase = LP_LEAF1; }
    else if (PH == LP_LEAF1) { s.pl1 = pc; s.phase = LP_LEAF2; }
    else if (PH == LP_LEAF2) {
        // reference :516-531 for sorted points a < b (< c); every orientation it asks is the sign of (a, b, c)
        const int a = s.ia, b = s.ia + 1, c = s.ib;
        const int e0 = lp_alloc(f, s, pr);
        if (e0 < 0) { s.phase = LP_DONE; return; }
        const int h0 = 2 * e0;
        if (c == b) {                                     // two points: one edge, two single-entry rings
            lf_set_rec(f, h0, b, h0, h0); lf_set_rec(f, h0 ^ 1, a, h0 ^ 1, h0 ^ 1);
            f.head[a] = (u16)h0; f.head[b] = (u16)(h0 ^ 1);
            lp_next_node(f, s); return;
        }
        // further edges: the free list's links are read directly here (leaves only)
        LPair g2 = pr; if (s.free_head != (int)GT_NIL) g2 = lf_pair(f, 2 * s.free_head);
        const int e1 = lp_alloc(f, s, g2);
        if (e1 < 0) { s.phase = LP_DONE; return; }
        const int h1 = 2 * e1;
        const bool s1 = sg > 0, s2 = sg < 0;              // ccw(a,b,c), ccw(a,c,b)
        int h2 = -1;
        if (s1 || s2) {
            LPair g3 = pr; if (s.free_head != (int)GT_NIL) g3 = lf_pair(f, 2 * s.free_head);
            const int e2 = lp_alloc(f, s, g3);
            if (e2 < 0) { s.phase = LP_DONE; return; }
            h2 = 2 * e2;
        }
        // ring of b: {b->a, b->c}; "first" moves to b->c iff c is right of the ray b->a (= ccw(a,b,c))
        lf_set_rec(f, h0 ^ 1, a, h1, h1); lf_set_rec(f, h1, c, h0 ^ 1, h0 ^ 1);
        f.head[b] = (u16)(s1 ? h1 : (h0 ^ 1));
        if (h2 < 0) {                                     // collinear: a and c keep single-entry rings
            lf_set_rec(f, h0, b, h0, h0); f.head[a] = (u16)h0;
            lf_set_rec(f, h1 ^ 1, b, h1 ^ 1, h1 ^ 1); f.head[c] = (u16)(h1 ^ 1);
        } else {
            // ring of a: {a->b, a->c}; first moves to a->c iff c is right of a->b (= ccw(a,c,b))
            lf_set_rec(f, h0, b, h2, h2); lf_set_rec(f, h2, c, h0, h0);
            f.head[a] = (u16)(s2 ? h2 : h0);
            // ring of c: {c->b, c->a}; first moves to c->a iff a is right of c->b (= ccw(a,b,c))
            lf_set_rec(f, h1 ^ 1, b, h2 ^ 1, h2 ^ 1); lf_set_rec(f, h2 ^ 1, a, h1 ^ 1, h1 ^ 1);
            f.head[c] = (u16)(s1 ? (h2 ^ 1) : (h1 ^ 1));
        }
        lp_next_node(f, s);
    }
    // -------------------------------------------------------------- tangents (reference uct :373-406, lct :337-370)
    else if (PH == LP_TINIT0) { s.hl = s.up ? hdA : (int)pr.a.prv; s.pli = pc; s.phase = LP_TINIT1; }     // upper: x -> first(x); lower: x -> pred(x, first(x))
    else if (PH == LP_TINIT1) { s.hr = s.up ? (int)pr.a.prv : hdA; s.pri = pc; s.phase = LP_TAN_R; }      // upper: y -> pred(y, first(y)); lower: y -> first(y)
    else if (PH == LP_TAN_R) {
        if (res) { s.hr = s.up ? pr.b.prv : pr.b.nxt; s.ty = vc; s.pri = pc; }      // pred(rfast, y) / succ(rfast, y)
        else s.phase = LP_TAN_L;
    } else if (PH == LP_TAN_L) {
        if (res) { s.hl = s.up ? pr.b.nxt : pr.b.prv; s.tx = vc; s.pli = pc; s.phase = LP_TAN_R; }
        else if (s.up) {
            s.ux = s.tx; s.uy = s.ty;
            s.up = 0; s.tx = (s.ia + s.ib) / 2; s.ty = s.tx + 1; s.phase = LP_TINIT0;
        } else {
            // first cross edge = lower tangent: both ends take it in their hull gap (before "first")
            s.li = s.tx; s.ri = s.ty; s.first_edge = 1; s.vR = 0; s.vL = 0;
            s.phase = LP_F;                               // LP_F allocates the edge and queues the inserts
        }
    }
    // -------------------------------------------------------------- ring insertion (reference insertNode :210-245)
    else if (PH == LP_INS_T0 || PH == LP_INS_SCAN) {
        int at = -1, nx = -1, hd; bool newhead = false, placed = true;
        if (PH == LP_INS_T0) {
            hd = hdA; const int last = pr.a.prv;
            if (!res) { if (s.iGap) { at = last; nx = hd; } else { at = s.iAt; nx = s.iNx; } }
            else if (last == hd) { at = hd; nx = hd; newhead = true; }             // single-entry ring: right of all
            else { s.iCur = last; s.iHd = hd; s.iLast = last; s.phase = LP_INS_SCAN; placed = false; }
        } else {
            hd = s.iHd;
            if (res) {
                const int cur = pr.a.prv;
                if (cur == hd) { at = s.iLast; nx = hd; newhead = true; }
                else { s.iCur = cur; placed = false; }
            } else { at = s.iCur; nx = pr.a.nxt; }
        }
        if (placed) {
            lf_set_nxt(f, at, s.iH); lf_set_prv(f, nx, s.iH); lf_set_rec(f, s.iH, s.iIN, nx, at);
            if (newhead) f.head[s.iP] = (u16)s.iH;
            const int newhd = newhead ? s.iH : hd;
            if (s.iWhich == 0) {
                s.at0 = at; s.nx0 = nx; s.headL = newhd;
                s.iP = s.ri; s.iIN = s.li; s.iH = s.i1H; s.iAt = s.i1At; s.iNx = s.i1Nx; s.iGap = s.i1Gap; s.iWhich = 1;
                s.phase = LP_INS_T0;
            } else {
                s.headR = newhd;
                s.e = s.iH ^ 1;                            // li -> ri
                s.hl1 = s.nx0;                             // succ(li, ri)
                s.hr1 = at;                                // pred(ri, li)
                if (s.li == s.ux && s.ri == s.uy) lp_next_node(f, s); else s.phase = LP_VR;
            }
        }
    }
    // -------------------------------------------------------------- merge step (reference :552-594)
    else if (PH == LP_VR) {
        s.r1 = vc; s.pr1 = pc; s.hr2 = pr.a.prv; s.t1_prv = pr.b.prv; s.t1_nxt = pr.b.nxt;
        s.vR = res; s.phase = res ? LP_CR : LP_VL;
    } else if (PH == LP_CR) {
        if (res) {
            // cutVerts(ri, r1): ri's ring loses hr1 (its next is the cross edge), r1's ring loses the twin
            const int h = s.hr1, tw = h ^ 1;
            lf_set_nxt(f, s.hr2, s.e ^ 1); lf_set_prv(f, s.e ^ 1, s.hr2);
            if (s.headR == h) { s.headR = s.e ^ 1; f.head[s.ri] = (u16)(s.e ^ 1); }
            lf_set_nxt(f, s.t1_prv, s.t1_nxt); lf_set_prv(f, s.t1_nxt, s.t1_prv);
            if (hdA == tw) f.head[s.r1] = (u16)((s.t1_nxt == tw) ? GT_NIL : s.t1_nxt);
            lp_free(f, s, h >> 1);
            s.hr1 = s.hr2; s.r1 = vc; s.pr1 = pc; s.hr2 = pr.a.prv; s.t1_prv = pr.b.prv; s.t1_nxt = pr.b.nxt;
        } else s.phase = LP_VL;
    } else if (PH == LP_VL) {
        s.l1 = vc; s.pl1 = pc; s.hl2 = pr.a.nxt; s.u1_prv = pr.b.prv; s.u1_nxt = pr.b.nxt;
        s.vL = res; s.phase = res ? LP_CL : LP_F;
    } else if (PH == LP_CL) {
        if (res) {
            const int h = s.hl1, tw = h ^ 1;
            lf_set_nxt(f, s.e, s.hl2); lf_set_prv(f, s.hl2, s.e);
            if (s.headL == h) { s.headL = s.hl2; f.head[s.li] = (u16)s.hl2; }
            lf_set_nxt(f, s.u1_prv, s.u1_nxt); lf_set_prv(f, s.u1_nxt, s.u1_prv);
            if (hdA == tw) f.head[s.l1] = (u16)((s.u1_nxt == tw) ? GT_NIL : s.u1_nxt);
            lp_free(f, s, h >> 1);
            s.hl1 = s.hl2; s.l1 = vc; s.pl1 = pc; s.hl2 = pr.a.nxt; s.u1_prv = pr.b.prv; s.u1_nxt = pr.b.nxt;
        } else s.phase = LP_F;
    } else if (PH == LP_F) {
        const int ne = lp_alloc(f, s, pr);
        if (ne < 0) { s.phase = LP_DONE; return; }
        const int h0 = 2 * ne;
        if (s.first_edge) {
            s.first_edge = 0;
            lp_queue_inserts(s, h0, 0, 0, 1, 0, 0, 1);
            return;
        }
        const bool go_left = !s.vR ? true : (!s.vL ? false : res);        // ties prefer the right candidate (:588-593)
        if (go_left) {
            // new edge l1 -> ri: at l1 after the edge back to the old li, at ri before the old cross edge
            const int at0 = s.hl1 ^ 1, nx0 = s.u1_nxt, at1 = s.hr1, nx1 = s.e ^ 1;
            s.li = s.l1; s.pli = s.pl1;
            lp_queue_inserts(s, h0, at0, nx0, 0, at1, nx1, 0);
        } else {
            // new edge li -> r1: at li after the old cross edge, at r1 before the edge back to the old ri
            const int at0 = s.e, nx0 = s.hl1, at1 = s.t1_prv, nx1 = s.hr1 ^ 1;
            s.ri = s.r1; s.pri = s.pr1;
            lp_queue_inserts(s, h0, at0, nx0, 0, at1, nx1, 0);
        }
    }
}

// Generic entry (host emulation): dispatch on the frame's phase. Returns false when the tree is finished.
GT_INLINE bool lpf_step(const LFrame& f, LState& s) {
    switch (s.phase) {
    case LP_LEAF0: lpf_step_t<LP_LEAF0>(f, s); break;
    case LP_LEAF1: lpf_step_t<LP_LEAF1>(f, s); break;
    case LP_LEAF2: lpf_step_t<LP_LEAF2>(f, s); break;
    case LP_TINIT0: lpf_step_t<LP_TINIT0>(f, s); break;
    case LP_TINIT1: lpf_step_t<LP_TINIT1>(f, s); break;
    case LP_TAN_R: lpf_step_t<LP_TAN_R>(f, s); break;
    case LP_TAN_L: lpf_step_t<LP_TAN_L>(f, s); break;
    case LP_INS_T0: lpf_step_t<LP_INS_T0>(f, s); break;
    case LP_INS_SCAN: lpf_step_t<LP_INS_SCAN>(f, s); break;
    case LP_VR: lpf_step_t<LP_VR>(f, s); break;
    case LP_CR: lpf_step_t<LP_CR>(f, s); break;
    case LP_VL: lpf_step_t<LP_VL>(f, s); break;
    case LP_CL: lpf_step_t<LP_CL>(f, s); break;
    case LP_F: lpf_step_t<LP_F>(f, s); break;
    default: return false;
    }
    return s.phase != LP_DONE;
}

// Triangles owned by sorted vertex i, in the reference's emission order (convertTrisFreeAdj :606-635);
// same rule as dt_core.cuh::vertex_triangles on the packed records.
template <class Visit>
GT_INLINE int lf_vertex_triangles(const LFrame& f, int i, Visit&& visit) {
    const int hd = f.head[i];
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

}  // namespace gt
