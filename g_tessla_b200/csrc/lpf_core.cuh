// lpf_core.cuh -- the Delaunay divide-and-conquer as a per-lane state machine ("lane per frame").
//
// dt_core.cuh::merge is a nest of data-dependent loops: when the lanes of a warp walk different merges
// they diverge and the warp runs them one after another. Here the same algorithm -- the same predicates
// on the same operands in the same order, hence the same rings as the reference (src/delaunay_tri.c:
// 511-601, tangents :337-406, insertNode :210-245, deleteNode :247-268) -- is unrolled into a flat
// loop of uniform ITERATIONS. One iteration of one lane is:
//     [address A: ring entry of a vertex] -> [address B: one edge = its two half-edge records]
//     -> [address C: one point] -> ONE predicate (orient2d or incircle) -> a short transition.
// Every lane of a warp executes that same sequence whatever phase of whatever merge of whatever frame it
// is in, so 32 lanes = 32 independent frames advance in lock step, their (uncoalesced) loads in flight
// together. The per-lane state is a couple of dozen registers; the frame lives in global memory:
//   point   32 B  {x, y, z, -}          (sorted order)
//   record   8 B  {dst, nxt, prv, -}    per half-edge, twins adjacent (edge = 16 B, one load)
//   head     2 B  per vertex
// Edge pool, free list and the cursor of the recursion tree are private to the lane: no atomics, no barriers.
//
// Host-compilable: tests/emu drives lpf_step() sequentially and compares with the oracle.
#pragma once
#include "dt_core.cuh"

namespace gt {

struct LRec { u16 dst, nxt, prv, pad; };                 // one half-edge
struct LPair { LRec a, b; };                             // a = the half-edge asked for, b = its twin
struct LPt { double x, y, z, w; };

// Frame storage in global memory (one lane, one frame).
struct LFrame {
    LPt* pt;        // [n]
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
// device workspace keeps 16-byte (x, y) points (z apart): f.pt is really a Pt* there
GT_INLINE Pt lf_pt(const LFrame& f, int v) { const double2 d = *(reinterpret_cast<const double2*>(f.pt) + v); Pt p; p.x = d.x; p.y = d.y; return p; }
#else
GT_INLINE LPair lf_pair(const LFrame& f, int h) { LPair p; p.a = f.rec[h]; p.b = f.rec[h ^ 1]; return p; }
GT_INLINE Pt lf_pt(const LFrame& f, int v) { Pt p; p.x = f.pt[v].x; p.y = f.pt[v].y; return p; }
#endif
GT_INLINE void lf_set_nxt(const LFrame& f, int h, int v) { f.rec[h].nxt = (u16)v; }
GT_INLINE void lf_set_prv(const LFrame& f, int h, int v) { f.rec[h].prv = (u16)v; }
GT_INLINE void lf_set_rec(const LFrame& f, int h, int dst, int nxt, int prv) {
    LRec r; r.dst = (u16)dst; r.nxt = (u16)nxt; r.prv = (u16)prv; r.pad = 0; f.rec[h] = r;
}

enum LPhase : int {
    LP_NODE = 0,      // (unused: the node cursor advances inside the transitions, lp_next_node)
    LP_LEAF0, LP_LEAF1, LP_LEAF2,
    LP_TINIT0, LP_TINIT1, LP_TAN_R, LP_TAN_L,
    LP_INS_T0, LP_INS_SCAN,
    LP_VR, LP_CR, LP_VL, LP_CL, LP_F,
    LP_DONE           // whole tree finished
};

// Per-lane state. Everything a phase needs besides its three loads is kept here (registers on the device).
struct LState {
    int phase;
    int d, k, ia, ib;                 // current node
    uint32_t err;
    int bump, free_head;              // private edge pool
    // tangents
    int tx, ty, hl, hr; Pt ptx, pty; int up;
    int t_hd0, t_pv0, t_hd1, t_pv1; Pt pm0, pm1;     // ring entries at the start vertices, saved for the upper tangent
    int lx, ly; Pt plx, ply;
    int ux, uy;
    // merge step
    int li, ri, e; Pt pli, pri; int headL, headR;
    int hr1, r1, hr2; Pt pr1; int t1_prv, t1_nxt;    // right candidate: ri->r1, its predecessor, twin's links at r1
    int hl1, l1, hl2; Pt pl1; int u1_prv, u1_nxt;    // left candidate: li->l1, its successor, twin's links at l1
    int vR, vL;
    // insertion in flight
    int iP, iIN, iH, iAt, iNx, iGap, iWhich, iCur, iHd, iLast; Pt pP, pIN;
    int at0, nx0;                                      // where endpoint 0's half-edge went
    int i1P, i1IN, i1H, i1At, i1Nx, i1Gap; Pt p1P, p1IN;   // endpoint 1, queued
    int go_left, first_edge;
    // leaf scratch
    Pt qa, qb;
    // fetch / predicate descriptor of the next iteration (lp_prepare)
    int fa, fb, fc, bA, cB, kind, slot, rp0, rp1, rp2; Pt oX, oA, oB, oY;
};

GT_INLINE int lp_sign_ccw(Pt x, Pt a, Pt b) { return orient2d_sign(x.x, x.y, a.x, a.y, b.x, b.y); }

GT_INLINE void lp_start(LState& s) {
    s.phase = LP_NODE; s.d = -1; s.k = 0; s.err = 0; s.bump = 0; s.free_head = GT_NIL;
}

GT_INLINE int lp_alloc(const LFrame& f, LState& s, const LPair& freerec) {
    if (s.free_head != (int)GT_NIL) { int e = s.free_head; s.free_head = freerec.a.nxt; return e; }
    if (s.bump >= f.cap_e) { s.err |= GT_ERR_POOL; return -1; }
    return s.bump++;
}
GT_INLINE void lp_free(const LFrame& f, LState& s, int e) { lf_set_nxt(f, 2 * e, s.free_head); s.free_head = e; }

// queue the two ring insertions of a new edge h0 = a->b (endpoint 0 at a, endpoint 1 at b)
GT_INLINE void lp_queue_inserts(LState& s, int h0, int a, Pt pa, int b, Pt pb, int at0, int nx0, int gap0, int at1, int nx1, int gap1) {
    s.iP = a; s.pP = pa; s.iIN = b; s.pIN = pb; s.iH = h0; s.iAt = at0; s.iNx = nx0; s.iGap = gap0; s.iWhich = 0;
    s.i1P = b; s.p1P = pb; s.i1IN = a; s.p1IN = pa; s.i1H = h0 ^ 1; s.i1At = at1; s.i1Nx = nx1; s.i1Gap = gap1;
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
        s.phase = LP_TINIT0; s.up = 0; s.tx = (ia + ib) / 2; s.ty = s.tx + 1;
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

// Fill the fetch / predicate descriptor of the phase the lane is now in. Called at the end of a transition, so
// that the top of the next iteration is one straight-line sequence for every lane whatever its phase
// (written as per-phase branches up there, the compiler specialises the loads per phase and the warp ends up
// executing 13 copies of the load-predicate sequence one after another: measured 37k cycles per iteration).
GT_INLINE void lp_prepare(LState& s) {
    int fa = -1, fb = -1, fc = -1, bA = 0, cB = 0, kind = 0, slot = 0;
    Pt oX, oA, oB, oY; oX.x = oX.y = 0.0; oA = oB = oY = oX;
    int r0 = -2, r1 = -2, r2 = -2;                        // never equal to a vertex or to the 'no point' marker -1
    const int freeh = (s.free_head != (int)GT_NIL) ? 2 * s.free_head : -1;
    switch (s.phase) {
    case LP_LEAF0: fc = s.ia; break;
    case LP_LEAF1: fc = s.ia + 1; break;
    case LP_LEAF2: fc = s.ib; fb = freeh; if (s.ib - s.ia == 2) { kind = 3; oX = s.qa; oA = s.qb; slot = 2; } break;
    case LP_TINIT0: fa = s.tx; bA = 1; fc = s.tx; break;
    case LP_TINIT1: fa = s.ty; bA = 1; fc = s.ty; break;
    case LP_TAN_R: case LP_TAN_L:
        fb = (s.phase == LP_TAN_R) ? s.hr : s.hl; cB = 1; kind = 1; slot = 0;
        oA = s.up ? s.ptx : s.pty; oB = s.up ? s.pty : s.ptx; break;
    case LP_INS_T0: fa = s.iP; bA = 1; cB = 1; kind = 1; slot = 1; oX = s.pIN; oB = s.pP; break;
    case LP_INS_SCAN: fb = s.iCur; cB = 1; kind = 1; slot = 1; oX = s.pIN; oB = s.pP; break;
    case LP_VR: fb = s.hr1; cB = 1; kind = 1; slot = 0; oA = s.pli; oB = s.pri; break;
    case LP_VL: fb = s.hl1; cB = 1; kind = 1; slot = 0; oA = s.pli; oB = s.pri; break;
    case LP_CR: fa = s.r1; fb = s.hr2; cB = 1; kind = 2; slot = 3; oX = s.pr1; oA = s.pli; oB = s.pri; r0 = s.li; r1 = s.ri; r2 = s.r1; break;
    case LP_CL: fa = s.l1; fb = s.hl2; cB = 1; kind = 2; slot = 3; oX = s.pli; oA = s.pri; oB = s.pl1; r0 = s.li; r1 = s.ri; r2 = s.l1; break;
    case LP_F: fb = freeh; if (s.vR && s.vL) { kind = 2; slot = 4; oX = s.pli; oA = s.pri; oB = s.pr1; oY = s.pl1; } break;
    default: break;
    }
    s.fa = fa; s.fb = fb; s.fc = fc; s.bA = bA; s.cB = cB; s.kind = kind; s.slot = slot;
    s.oX = oX; s.oA = oA; s.oB = oB; s.oY = oY; s.rp0 = r0; s.rp1 = r1; s.rp2 = r2;
}

GT_INLINE void lp_begin(const LFrame& f, LState& s) { lp_start(s); lp_next_node(f, s); lp_prepare(s); }

// One iteration of one lane. Returns false when the lane's tree is finished.
GT_INLINE bool lpf_step(const LFrame& f, LState& s) {
    if (s.phase == LP_DONE) return false;

    // ------------------------------------------------------------------ uniform part: three dependent loads, one predicate
    int hdA = 0;
    if (s.fa >= 0) hdA = f.head[s.fa];
    const int hb = s.bA ? hdA : s.fb;
    LPair pr; pr.a.dst = pr.a.nxt = pr.a.prv = pr.a.pad = 0; pr.b = pr.a;
    if (hb >= 0) pr = lf_pair(f, hb);
    const int vc = s.cB ? (int)pr.a.dst : s.fc;
    Pt pc; pc.x = pc.y = 0.0;
    if (vc >= 0) pc = lf_pt(f, vc);
    // kind 1 / 3: orient2d(X, A, B)    kind 2: incircle(X, A, B, Y); the loaded point takes operand `slot`
    const Pt X = s.slot == 0 ? pc : s.oX, A = s.slot == 1 ? pc : s.oA, B = s.slot == 2 ? pc : s.oB, Y = s.slot == 3 ? pc : s.oY;
    const bool rep = (vc == s.rp0) | (vc == s.rp1) | (vc == s.rp2);     // incircle with a repeated point is exactly 0
    int sg = 0;
    if (s.kind == 2) { if (!rep) sg = incircle_sign(X.x, X.y, A.x, A.y, B.x, B.y, Y.x, Y.y); }
    else if (s.kind != 0) sg = orient2d_sign(X.x, X.y, A.x, A.y, B.x, B.y);
    if (sg == GT_RANGE_ERR) { s.err |= GT_ERR_RANGE; sg = 0; }
    const bool res = sg > 0;

    // ------------------------------------------------------------------ transitions (short, stores only)
    switch (s.phase) {
    case LP_LEAF0: s.qa = pc; s.phase = LP_LEAF1; break;
    case LP_LEAF1: s.qb = pc; s.phase = LP_LEAF2; break;
    case LP_LEAF2: {
        // reference :516-531 for sorted points a < b (< c); every orientation it asks is the sign of (a, b, c)
        const int a = s.ia, b = s.ia + 1, c = s.ib;
        const int e0 = lp_alloc(f, s, pr);
        if (e0 < 0) { s.phase = LP_DONE; break; }
        const int h0 = 2 * e0;
        if (c == b) {                                     // two points: one edge, two single-entry rings
            lf_set_rec(f, h0, b, h0, h0); lf_set_rec(f, h0 ^ 1, a, h0 ^ 1, h0 ^ 1);
            f.head[a] = (u16)h0; f.head[b] = (u16)(h0 ^ 1);
            lp_next_node(f, s); break;
        }
        // further edges: the free list's links are read directly here (leaves only)
        LPair g2 = pr; if (s.free_head != (int)GT_NIL) g2 = lf_pair(f, 2 * s.free_head);
        const int e1 = lp_alloc(f, s, g2);
        if (e1 < 0) { s.phase = LP_DONE; break; }
        const int h1 = 2 * e1;
        const bool s1 = sg > 0, s2 = sg < 0;              // ccw(a,b,c), ccw(a,c,b)
        int h2 = -1;
        if (s1 || s2) {
            LPair g3 = pr; if (s.free_head != (int)GT_NIL) g3 = lf_pair(f, 2 * s.free_head);
            const int e2 = lp_alloc(f, s, g3);
            if (e2 < 0) { s.phase = LP_DONE; break; }
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
        lp_next_node(f, s); break;
    }
    // -------------------------------------------------------------- tangents (reference lct :337-370, uct :373-406)
    case LP_TINIT0: s.t_hd0 = hdA; s.t_pv0 = pr.a.prv; s.pm0 = pc; s.phase = LP_TINIT1; break;
    case LP_TINIT1:
        s.t_hd1 = hdA; s.t_pv1 = pr.a.prv; s.pm1 = pc;
        s.ptx = s.pm0; s.pty = s.pm1; s.up = 0;
        s.hl = s.t_pv0;            // lower: x -> pred(x, first(x))
        s.hr = s.t_hd1;            //        y -> first(y)
        s.phase = LP_TAN_R; break;
    case LP_TAN_R:
        if (res) { s.hr = s.up ? pr.b.prv : pr.b.nxt; s.ty = vc; s.pty = pc; }      // pred(rfast, y) / succ(rfast, y)
        else s.phase = LP_TAN_L;
        break;
    case LP_TAN_L:
        if (res) { s.hl = s.up ? pr.b.nxt : pr.b.prv; s.tx = vc; s.ptx = pc; s.phase = LP_TAN_R; }
        else if (!s.up) {
            s.lx = s.tx; s.ly = s.ty; s.plx = s.ptx; s.ply = s.pty;
            s.up = 1; s.tx = (s.ia + s.ib) / 2; s.ty = s.tx + 1; s.ptx = s.pm0; s.pty = s.pm1;
            s.hl = s.t_hd0;        // upper: x -> first(x)
            s.hr = s.t_pv1;        //        y -> pred(y, first(y))
            s.phase = LP_TAN_R;
        } else {
            s.ux = s.tx; s.uy = s.ty;
            // first cross edge = lower tangent: both ends take it in their hull gap (before "first")
            s.li = s.lx; s.ri = s.ly; s.pli = s.plx; s.pri = s.ply; s.first_edge = 1;
            s.phase = LP_F; s.vR = 0; s.vL = 0;           // LP_F allocates the edge and queues the inserts
        }
        break;
    // -------------------------------------------------------------- ring insertion (reference insertNode :210-245)
    case LP_INS_T0: case LP_INS_SCAN: {
        int at = -1, nx = -1; bool newhead = false, placed = true;
        int hd;
        if (s.phase == LP_INS_T0) {
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
        if (!placed) break;
        lf_set_nxt(f, at, s.iH); lf_set_prv(f, nx, s.iH); lf_set_rec(f, s.iH, s.iIN, nx, at);
        if (newhead) f.head[s.iP] = (u16)s.iH;
        const int newhd = newhead ? s.iH : hd;
        if (s.iWhich == 0) {
            s.at0 = at; s.nx0 = nx; s.headL = newhd;       // endpoint 0 is always the (new) left vertex li
            s.iP = s.i1P; s.pP = s.p1P; s.iIN = s.i1IN; s.pIN = s.p1IN; s.iH = s.i1H; s.iAt = s.i1At; s.iNx = s.i1Nx; s.iGap = s.i1Gap; s.iWhich = 1;
            s.phase = LP_INS_T0;
        } else {
            s.headR = newhd;
            s.e = s.iH ^ 1;                                // li -> ri
            s.hl1 = s.nx0;                                 // succ(li, ri)
            s.hr1 = at;                                    // pred(ri, li)
            if (s.li == s.ux && s.ri == s.uy) lp_next_node(f, s); else s.phase = LP_VR;
        }
        break;
    }
    // -------------------------------------------------------------- merge step (reference :552-594)
    case LP_VR:
        s.r1 = vc; s.pr1 = pc; s.hr2 = pr.a.prv; s.t1_prv = pr.b.prv; s.t1_nxt = pr.b.nxt;
        s.vR = res; s.phase = res ? LP_CR : LP_VL;
        break;
    case LP_CR:
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
        break;
    case LP_VL:
        s.l1 = vc; s.pl1 = pc; s.hl2 = pr.a.nxt; s.u1_prv = pr.b.prv; s.u1_nxt = pr.b.nxt;
        s.vL = res; s.phase = res ? LP_CL : LP_F;
        break;
    case LP_CL:
        if (res) {
            const int h = s.hl1, tw = h ^ 1;
            lf_set_nxt(f, s.e, s.hl2); lf_set_prv(f, s.hl2, s.e);
            if (s.headL == h) { s.headL = s.hl2; f.head[s.li] = (u16)s.hl2; }
            lf_set_nxt(f, s.u1_prv, s.u1_nxt); lf_set_prv(f, s.u1_nxt, s.u1_prv);
            if (hdA == tw) f.head[s.l1] = (u16)((s.u1_nxt == tw) ? GT_NIL : s.u1_nxt);
            lp_free(f, s, h >> 1);
            s.hl1 = s.hl2; s.l1 = vc; s.pl1 = pc; s.hl2 = pr.a.nxt; s.u1_prv = pr.b.prv; s.u1_nxt = pr.b.nxt;
        } else s.phase = LP_F;
        break;
    case LP_F: {
        const int ne = lp_alloc(f, s, pr);
        if (ne < 0) { s.phase = LP_DONE; break; }
        const int h0 = 2 * ne;
        if (s.first_edge) {
            s.first_edge = 0;
            lp_queue_inserts(s, h0, s.li, s.pli, s.ri, s.pri, 0, 0, 1, 0, 0, 1);
            break;
        }
        const bool go_left = !s.vR ? true : (!s.vL ? false : res);        // ties prefer the right candidate (:588-593)
        if (go_left) {
            // new edge l1 -> ri: at l1 after the edge back to the old li, at ri before the old cross edge
            const int at0 = s.hl1 ^ 1, nx0 = s.u1_nxt, at1 = s.hr1, nx1 = s.e ^ 1;
            s.li = s.l1; s.pli = s.pl1;
            lp_queue_inserts(s, h0, s.li, s.pli, s.ri, s.pri, at0, nx0, 0, at1, nx1, 0);
        } else {
            // new edge li -> r1: at li after the old cross edge, at r1 before the edge back to the old ri
            const int at0 = s.e, nx0 = s.hl1, at1 = s.t1_prv, nx1 = s.hr1 ^ 1;
            s.ri = s.r1; s.pri = s.pr1;
            lp_queue_inserts(s, h0, s.li, s.pli, s.ri, s.pri, at0, nx0, 0, at1, nx1, 0);
        }
        break;
    }
    default: break;
    }
    if (s.phase == LP_DONE) return false;
    lp_prepare(s);
    return true;
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
