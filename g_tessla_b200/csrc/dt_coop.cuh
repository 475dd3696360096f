// dt_coop.cuh -- warp-cooperative version of the Delaunay merge step (one warp = one merge).
//
// The upper levels of the reference's recursion tree (src/delaunay_tri.c:535-599) hold few, long
// merges, and a merge walk is inherently sequential: every step of the "rising bubble" asks a handful
// of predicates whose outcomes select the next step. One thread per merge (dt_core.cuh::merge) runs
// them as one dependent FP64 chain after another, and lanes of a warp that walk DIFFERENT merges
// diverge and serialise. Here a whole warp owns one merge and evaluates, in ONE round per step, every
// predicate the reference could ask for in that step:
//   lane 0-1    candidate validity                          (:556, :569)
//   lane 2-5    the first two tests of each incircle deletion chain (:559, :572)
//   lane 6-9    the final choice for (0|1, 0|1) deletions   (:588)
//   lane 10-25  "is the new neighbour right of first / of last" of the two ring insertions that
//               follow (:216-224), for both outcomes and both possible surviving candidates
// and then picks, with one ballot, the outcomes the sequential code would have used. A second round
// runs only when a chain deletes more than one edge or a deletion moved the "first" of an endpoint.
// The decisions, their order and the resulting rings are exactly those of dt_core.cuh::merge (same
// predicates on the same operands, evaluated speculatively and in parallel).
//
// Device only (warp shuffles); validated on the GPU against the oracle by tests/test_gpu_parity.py.
#pragma once
#include "dt_core.cuh"

namespace gt {

#if defined(__CUDACC__)

// Exact fallbacks, out of line and rare.
static __device__ __noinline__ bool ccw_slow(uint32_t* err, Pt px, Pt pa, Pt pb) {
    int s = orient2d_exact(px.x, px.y, pa.x, pa.y, pb.x, pb.y);
    if (s == GT_RANGE_ERR) { atomic_or_u32(err, GT_ERR_RANGE); return false; }
    return s > 0;
}
static __device__ __noinline__ bool inc_slow(uint32_t* err, Pt px, Pt pa, Pt pb, Pt py) {
    int s = incircle_exact(px.x, px.y, pa.x, pa.y, pb.x, pb.y, py.x, py.y);
    if (s == GT_RANGE_ERR) { atomic_or_u32(err, GT_ERR_RANGE); return false; }
    return s > 0;
}

// One predicate per lane, one code path for the warp: every lane evaluates the FP64 filters of BOTH
// ccw(x, a, b) (reference predicates.c:1628-1651) and incircle(x, a, b, y) (:3238-3267) on its own
// operands and keeps the one its role asks for (kind 1 / 2, 0 = idle). Lanes with different roles
// never diverge; only a lane whose filter is inconclusive branches to the exact stage.
template <class F>
__device__ __forceinline__ bool lane_pred(const F& f, int kind, int x, int a, int b, int y) {
    const Pt px = f.pt[x], pa = f.pt[a], pb = f.pt[b], py = f.pt[y];
    const double detleft = (px.x - pb.x) * (pa.y - pb.y), detright = (px.y - pb.y) * (pa.x - pb.x);
    const double odet = detleft - detright;
    const double obound = GT_CCW_ERRBOUND_A * (fabs(detleft) + fabs(detright));
    // same-sign products are the only inconclusive case (:1632-1646); then |det| must reach the bound
    const bool osure = !((detleft > 0.0 && detright > 0.0) || (detleft < 0.0 && detright < 0.0)) || fabs(odet) >= obound;
    const double adx = px.x - py.x, bdx = pa.x - py.x, cdx = pb.x - py.x;
    const double ady = px.y - py.y, bdy = pa.y - py.y, cdy = pb.y - py.y;
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, alift = adx * adx + ady * ady;
    const double cdxady = cdx * ady, adxcdy = adx * cdy, blift = bdx * bdx + bdy * bdy;
    const double adxbdy = adx * bdy, bdxady = bdx * ady, clift = cdx * cdx + cdy * cdy;
    const double idet = alift * (bdxcdy - cdxbdy) + blift * (cdxady - adxcdy) + clift * (adxbdy - bdxady);
    const double perm = (fabs(bdxcdy) + fabs(cdxbdy)) * alift + (fabs(cdxady) + fabs(adxcdy)) * blift
                      + (fabs(adxbdy) + fabs(bdxady)) * clift;
    // a repeated point makes the determinant exactly 0 (e.g. the chain wrapped around a short ring)
    const bool rep = (x == y) | (a == y) | (b == y) | (x == a) | (x == b) | (a == b);
    const bool isure = rep | (fabs(idet) > GT_ICC_ERRBOUND_A * perm);
    bool res = (kind == 1) ? (odet > 0.0) : ((kind == 2) & !rep & (idet > 0.0));
    if (kind == 1 && !osure) res = ccw_slow(f.err, px, pa, pb);
    if (kind == 2 && !isure) res = inc_slow(f.err, px, pa, pb, py);
    return res;
}

// insertNode (reference :210-245) with its first two orientation tests precomputed:
// t0 = rightOf(in, parent, first), t1 = rightOf(in, parent, last). `after` = the slot known from the
// topology of the merge, used when t0 is false (see ring_insert_hint).
template <class F>
__device__ __forceinline__ void coop_insert(const F& f, int parent, int in, int h, int after, bool t0, bool t1) {
    f.dst[h] = (typename F::idx)in;
    const int hd = f.head[parent];
    if (!t0) { link_after(f, after, h); return; }
    int cur = f.prv[hd];
    if (cur != hd && t1) {
        cur = f.prv[cur];
        while (cur != hd && right_of(f, in, parent, f.dst[cur])) cur = f.prv[cur];
    }
    if (cur == hd) { f.head[parent] = (typename F::idx)h; link_after(f, f.prv[cur], h); }
    else link_after(f, cur, h);
}

// Orientation-only variant for the tangent walks.
template <class F>
__device__ __forceinline__ bool lane_ccw(const F& f, bool active, int x, int a, int b) {
    const Pt px = f.pt[x], pa = f.pt[a], pb = f.pt[b];
    const double detleft = (px.x - pb.x) * (pa.y - pb.y), detright = (px.y - pb.y) * (pa.x - pb.x);
    const double odet = detleft - detright;
    const bool osure = !((detleft > 0.0 && detright > 0.0) || (detleft < 0.0 && detright < 0.0)) ||
                       fabs(odet) >= GT_CCW_ERRBOUND_A * (fabs(detleft) + fabs(detright));
    bool res = odet > 0.0;
    if (active && !osure) res = ccw_slow(f.err, px, pa, pb);
    return active & res;
}

// Merge of the finished halves [.., mid] and [mid+1, ..] by one full warp. All 32 lanes call this
// with the same arguments; `ln` is used by lane 0 only.
//
// Candidate registry: after the gather, lane j holds candidate j as (vertex cv, half-edge ch):
//   0 R0  1 R1  2 R2   (ri's neighbours clockwise from li:        ch = prv^(j+1)(e^1))
//   3 L0  4 L1  5 L2   (li's neighbours counter-clockwise from ri: ch = nxt^(j-2)(e))
//   6 li  7 ri
// and any lane fetches an operand with one shuffle.
template <class F>
__device__ __forceinline__ void coop_merge(const F& f, Lane& ln, int mid, unsigned long long* dbg = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#ifdef GT_PHASE_TIMING
    long long c0_ = clock64();
#endif

    // ---- tangents (reference lct :337-370, uct :373-406): read-only walks. Lanes 0-1 walk the lower one,
    // lanes 2-3 the upper one; the even lane of a pair tests the right hull's candidate, the odd lane the left
    // hull's (the reference asks the second only if the first fails: same outcome, one round per move).
    int lx, ly, ux, uy;
    {
        const bool up = (lane >> 1) & 1, second = lane & 1;
        int tx = mid, ty = mid + 1;
        int hl = up ? f.head[tx] : f.prv[f.head[tx]];
        int hr = up ? f.prv[f.head[ty]] : f.head[ty];
        bool done = lane >= 4;
        for (;;) {
            const int cand = f.dst[second ? hl : hr];
            // lower: rightOf(p, x, y) = ccw(p, y, x); upper: leftOf(p, x, y) = ccw(p, x, y)
            const bool res = lane_ccw(f, !done, cand, up ? tx : ty, up ? ty : tx);
            const unsigned bal = __ballot_sync(FULL, res);
            const unsigned pr = (bal >> (lane & 2)) & 3u;          // this pair's (right test, left test)
            if (!done) {
                if (pr & 1u) { const int c = f.dst[hr]; const int t = hr ^ 1; hr = up ? f.prv[t] : f.nxt[t]; ty = c; }
                else if (pr & 2u) { const int c = f.dst[hl]; const int t = hl ^ 1; hl = up ? f.nxt[t] : f.prv[t]; tx = c; }
                else done = true;
            }
            if (__ballot_sync(FULL, !done) == 0u) break;
        }
        lx = __shfl_sync(FULL, tx, 0); ly = __shfl_sync(FULL, ty, 0);
        ux = __shfl_sync(FULL, tx, 2); uy = __shfl_sync(FULL, ty, 2);
    }

    // ---- first cross edge (lower tangent): both ends take it in their hull gap
    int e = 0;
    if (lane == 0) { const int ne = edge_alloc(f, ln); e = ne < 0 ? -1 : 2 * ne; }
    e = __shfl_sync(FULL, e, 0);
    if (e < 0) return;
    if (lane == 0) ring_insert_hint(f, lx, ly, e, f.prv[f.head[lx]]);
    else if (lane == 1) ring_insert_hint(f, ly, lx, e ^ 1, f.prv[f.head[ly]]);
    __syncwarp();
#ifdef GT_PHASE_TIMING
    long long c1_ = clock64(); int steps_ = 0, rounds_ = 0;
#endif

    // ---- per-lane role constants (shuffle source lanes into the candidate registry)
    //   lane 0 / 1    validity: ccw(R0 | L0, li, ri)                                   (:556, :569)
    //   lane 2, 3     right chain: inc(R_k, li, ri, R_k+1), k = 0, 1                    (:559)
    //   lane 4, 5     left chain:  inc(li, ri, L_k, L_k+1) = inc(L_k, li, ri, L_k+1)   (:572)
    //   lane 6..9     choice:      inc(li, ri, R_a, L_b)   = inc(R_a, li, ri, L_b), a, b in {0, 1}   (:588)
    //   lane 10..25   insertion tests rightOf(IN, P, first(P) | last(P)) = ccw(IN, q, P)  (:216-224):
    //                 ep = (lane-10)>>2: 0 (P = L_c, IN = ri)  1 (P = ri, IN = L_c)  2 (P = li, IN = R_c)  3 (P = R_c, IN = li),
    //                 c = ((lane-10)>>1)&1, odd lanes test "last"
    int kind = 0, selX = 6, selY = 6, selP = 6;
    if (lane == 0) { kind = 1; selX = 0; }
    else if (lane == 1) { kind = 1; selX = 3; }
    else if (lane < 4) { kind = 2; selX = lane - 2; selY = lane - 1; }
    else if (lane < 6) { kind = 2; selX = lane - 1; selY = lane; }
    else if (lane < 10) { kind = 2; selX = (lane - 6) >> 1; selY = 3 + ((lane - 6) & 1); }
    else if (lane < 26) {
        kind = 1;
        const int ep = (lane - 10) >> 2, c = ((lane - 10) >> 1) & 1;
        selP = ep == 0 ? 3 + c : (ep == 1 ? 7 : (ep == 2 ? 6 : c));
        selX = ep == 0 ? 7 : (ep == 1 ? 3 + c : (ep == 2 ? c : 6));
    }
    const bool ins_lane = lane >= 10 && lane < 26;
    const bool last_test = lane & 1;
    const int hops = lane < 6 ? (lane < 3 ? lane : lane - 3) + 1 : 0;
    const typename F::idx* const link = lane < 3 ? f.prv : f.nxt;

    int li = lx, ri = ly;
    int guard = 8 * f.n + 64;
    while (li != ux || ri != uy) {
        if (--guard < 0) { if (lane == 0) atomic_or_u32(f.err, GT_ERR_RING); return; }
        bool vR = false, vL = false, first_round = true;
        int nR, nL, ch, cv;
        unsigned bal;
#ifdef GT_PHASE_TIMING
        ++steps_;
#endif
        for (;;) {
#ifdef GT_PHASE_TIMING
            ++rounds_;
#endif
            // ---- gather: lanes 0-5 walk to their candidate
            ch = lane < 3 ? (e ^ 1) : e;
            if (hops > 0) ch = link[ch];
            if (hops > 1) ch = link[ch];
            if (hops > 2) ch = link[ch];
            cv = lane < 6 ? (int)f.dst[ch] : (lane == 6 ? li : ri);
            // ---- operands of this lane's test
            const int X = __shfl_sync(FULL, cv, selX);
            const int Y = __shfl_sync(FULL, cv, selY);
            const int P = __shfl_sync(FULL, cv, selP);
            int A = li, B = ri;
            if (ins_lane) {
                const int hd = f.head[P];
                A = f.dst[last_test ? f.prv[hd] : hd];
                B = P;
            }
            const int k = (first_round || lane >= 2) ? kind : 0;
            bal = __ballot_sync(FULL, lane_pred(f, k, X, A, B, Y));
            if (first_round) { vR = bal & 1u; vL = (bal >> 1) & 1u; }
            nR = vR ? ((bal >> 2) & 1u ? ((bal >> 3) & 1u ? 2 : 1) : 0) : 0;
            nL = vL ? ((bal >> 4) & 1u ? ((bal >> 5) & 1u ? 2 : 1) : 0) : 0;
            if ((nR | nL) == 0) break;                                      // the common case: nothing to delete
            // ---- apply the deletions (cutVerts :277-282). Lanes 0,1 own hr0,hr1 and lanes 3,4 own hl0,hl1: each
            // unlinks the twin of its edge; lane 0 / lane 3 also splice the ring of ri / li.
            const int survR = __shfl_sync(FULL, ch, nR), survL = __shfl_sync(FULL, ch, 3 + nL);
            const int h1R = __shfl_sync(FULL, ch, 1), h1L = __shfl_sync(FULL, ch, 4);
            const int h0L = __shfl_sync(FULL, ch, 3);
            if (lane < 2 ? lane < nR : (lane >= 3 && lane < 5 && lane - 3 < nL)) ring_unlink(f, cv, ch ^ 1);
            if (lane == 0 && nR > 0) {
                f.nxt[survR] = (typename F::idx)(e ^ 1); f.prv[e ^ 1] = (typename F::idx)survR;
                const int hd = f.head[ri];
                if (hd == ch || (nR > 1 && hd == h1R)) f.head[ri] = (typename F::idx)(e ^ 1);
            } else if (lane == 3 && nL > 0) {
                f.nxt[e] = (typename F::idx)survL; f.prv[survL] = (typename F::idx)e;
                const int hd = f.head[li];
                if (hd == ch || (nL > 1 && hd == h1L)) f.head[li] = (typename F::idx)survL;
            }
            __syncwarp();
            if (lane == 0) {
                if (nR > 0) edge_free(f, ln, ch >> 1);
                if (nR > 1) edge_free(f, ln, h1R >> 1);
                if (nL > 0) edge_free(f, ln, h0L >> 1);
                if (nL > 1) edge_free(f, ln, h1L >> 1);
            }
            first_round = false;
            if (nR < 2 && nL < 2) break;
            __syncwarp();
            // a chain used up its two speculated tests: look again from the survivors
            // (the validity flags are kept: the reference tests them once per step)
        }
        // survivors (relative to the last gather)
        const int r1 = __shfl_sync(FULL, cv, nR), hR = __shfl_sync(FULL, ch, nR);
        const int l1 = __shfl_sync(FULL, cv, 3 + nL), hL = __shfl_sync(FULL, ch, 3 + nL);
        // ---- choice (:582-593, ties prefer the right candidate); F was speculated for (nR, nL) in {0,1}^2
        const bool F = (bal >> (6 + 2 * nR + nL)) & 1u;
        const bool go_left = !vR ? true : (!vL ? false : F);
        // ---- insertion tests: speculated ones are valid unless this round's deletions touched the ring of the
        // endpoint that stays (ri when going left, li when going right); then redo the four tests.
        unsigned tb;     // bit 0/1: t0/t1 of the new vertex's ring, bit 2/3: of the staying endpoint's ring
        if (go_left ? (nR == 0) : (nL == 0)) {
            const int base = go_left ? 10 + 2 * nL : 18 + 2 * nR;
            tb = ((bal >> base) & 3u) | (((bal >> (base + 4)) & 3u) << 2);
            if (!go_left) tb = ((tb & 3u) << 2) | (tb >> 2);                 // (ep2, ep3) -> (new vertex = ep3, staying = ep2)
        } else {
            __syncwarp();
#ifdef GT_PHASE_TIMING
            ++rounds_;
#endif
            const int nv = go_left ? l1 : r1, st = go_left ? ri : li;     // new vertex, staying endpoint
            const int P = (lane & 2) ? st : nv, X = (lane & 2) ? nv : st;
            int A = li;
            if (lane < 4) { const int hd = f.head[P]; A = f.dst[last_test ? f.prv[hd] : hd]; }
            tb = __ballot_sync(FULL, lane_ccw(f, lane < 4, X, A, P)) & 15u;
        }
        int ne = 0;
        if (lane == 0) ne = edge_alloc(f, ln);
        ne = __shfl_sync(FULL, ne, 0);
        if (ne < 0) return;
        if (go_left) {
            // new edge l1 -> ri: at l1 it follows (CCW) the edge back to the old li, at ri it precedes the old cross edge
            if (lane == 0) coop_insert(f, l1, ri, 2 * ne, hL ^ 1, tb & 1u, (tb >> 1) & 1u);
            else if (lane == 1) coop_insert(f, ri, l1, 2 * ne + 1, f.prv[e ^ 1], (tb >> 2) & 1u, (tb >> 3) & 1u);
            li = l1;
        } else {
            // new edge li -> r1: at li it follows the old cross edge, at r1 it precedes the edge back to the old ri
            if (lane == 0) coop_insert(f, li, r1, 2 * ne, e, (tb >> 2) & 1u, (tb >> 3) & 1u);
            else if (lane == 1) coop_insert(f, r1, li, 2 * ne + 1, f.prv[hR ^ 1], tb & 1u, (tb >> 1) & 1u);
            ri = r1;
        }
        e = 2 * ne;
        __syncwarp();
    }
#ifdef GT_PHASE_TIMING
    if (lane == 0 && dbg) {
        long long c2_ = clock64();
        atomicAdd(&dbg[22], (unsigned long long)(c1_ - c0_)); atomicAdd(&dbg[23], (unsigned long long)(c2_ - c1_));
        atomicAdd(&dbg[24], (unsigned long long)rounds_); atomicAdd(&dbg[25], (unsigned long long)steps_); atomicAdd(&dbg[26], 1ull);
    }
#endif
}

#endif  // __CUDACC__

}  // namespace gt
