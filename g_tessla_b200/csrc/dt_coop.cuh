// dt_coop.cuh -- cooperative (16 lanes per merge) version of the Delaunay merge step.
//
// The upper levels of the reference's recursion tree (src/delaunay_tri.c:535-599) hold only a few,
// long merges, and a merge walk is inherently sequential: every step of the "rising bubble" tests a
// handful of predicates whose outcomes select the next step. One thread per merge (dt_core.cuh)
// therefore runs at the latency of one dependent FP64 chain after another. Here a group of 16 lanes
// owns one merge and evaluates, in one round, every predicate the reference COULD ask for in the
// current step -- candidate validity (:556, :569), the incircle deletion chains (:559, :572), the final
// choice (:588) and the "is the new neighbour right of first / of last" tests of the two ring
// insertions (:216-224) -- then picks the outcomes the sequential code would have used with ballots.
// The decisions, their order and the resulting rings are exactly those of dt_core.cuh::merge (the
// same predicates on the same operands; only evaluated speculatively and in parallel).
//
// Device only (warp shuffles); validated on the GPU against the oracle by tests/test_gpu_parity.py.
#pragma once
#include "dt_core.cuh"

namespace gt {

#if defined(__CUDACC__)

constexpr int GT_G = 16;     // lanes per cooperative merge (sub-warp group)

// insertNode (reference :210-245) with its first two orientation tests precomputed:
// t0 = rightOf(in, parent, first), t1 = rightOf(in, parent, last). `after` = the slot known from the
// topology of the merge, used when t0 is false (see ring_insert_hint).
__device__ __forceinline__ void coop_insert(const Frame& f, int parent, int in, int h, int after, bool t0, bool t1) {
    f.dst[h] = (u16)in;
    const int hd = f.head[parent];
    if (!t0) { link_after(f, after, h); return; }
    int cur = f.prv[hd];
    if (cur != hd && t1) {
        cur = f.prv[cur];
        while (cur != hd && right_of(f, in, parent, f.dst[cur])) cur = f.prv[cur];
    }
    if (cur == hd) { f.head[parent] = (u16)h; link_after(f, f.prv[cur], h); }
    else link_after(f, cur, h);
}

// Merge of the finished halves [.., mid] and [mid+1, ..] by the 16 lanes named in gmask.
// gl = lane index inside the group. All lanes of the group call this with the same arguments;
// `ln` is used by group lane 0 only.
__device__ __forceinline__ void coop_merge(const Frame& f, Lane& ln, int mid, unsigned gmask, int gl) {
    // ---- tangents (reference lct :337-370, uct :373-406): read-only walks, lane 0 and lane 1
    int tx = mid, ty = mid + 1;
    if (gl == 0) {
        int hr = f.head[ty], hl = f.prv[f.head[tx]];
        for (;;) {
            int rfast = f.dst[hr], lfast = f.dst[hl];
            if (right_of(f, rfast, tx, ty)) { hr = f.nxt[hr ^ 1]; ty = rfast; }
            else if (right_of(f, lfast, tx, ty)) { hl = f.prv[hl ^ 1]; tx = lfast; }
            else break;
        }
    } else if (gl == 1) {
        int hl = f.head[tx], hr = f.prv[f.head[ty]];
        for (;;) {
            int rfast = f.dst[hr], lfast = f.dst[hl];
            if (left_of(f, rfast, tx, ty)) { hr = f.prv[hr ^ 1]; ty = rfast; }
            else if (left_of(f, lfast, tx, ty)) { hl = f.nxt[hl ^ 1]; tx = lfast; }
            else break;
        }
    }
    __syncwarp(gmask);
    const int lx = __shfl_sync(gmask, tx, 0, GT_G), ly = __shfl_sync(gmask, ty, 0, GT_G);
    const int ux = __shfl_sync(gmask, tx, 1, GT_G), uy = __shfl_sync(gmask, ty, 1, GT_G);
    const int gshift = (threadIdx.x & 31) & ~(GT_G - 1);      // position of the group's lane 0 in the warp

    // ---- first cross edge (lower tangent): both ends take it in their hull gap
    int e = 0;
    if (gl == 0) { int ne = edge_alloc(f, ln); e = ne < 0 ? -1 : 2 * ne; }
    e = __shfl_sync(gmask, e, 0, GT_G);
    if (e < 0) return;
    if (gl == 0) ring_insert_hint(f, lx, ly, e, f.prv[f.head[lx]]);
    else if (gl == 1) ring_insert_hint(f, ly, lx, e ^ 1, f.prv[f.head[ly]]);
    __syncwarp(gmask);

    int li = lx, ri = ly;
    int guard = 8 * f.n + 64;
    while (li != ux || ri != uy) {
        if (--guard < 0) { if (gl == 0) atomic_or_u32(f.err, GT_ERR_RING); return; }
        bool vR = false, vL = false, first_round = true, haveF = false, F = false;
        int hr0, hr1, hr2, hr3, hl0, hl1, hl2, hl3, R0, R1, R2, R3, L0, L1, L2, L3;
        int nR, nL;
        // ---- phase 1: candidate validity + deletion chains (+ the final choice for <= 1 deletion a side).
        // Repeats only when a side deletes 3 edges in one step (its chain may go on).
        for (;;) {
            hr0 = f.prv[e ^ 1]; hr1 = f.prv[hr0]; hr2 = f.prv[hr1]; hr3 = f.prv[hr2];
            hl0 = f.nxt[e];     hl1 = f.nxt[hl0]; hl2 = f.nxt[hl1]; hl3 = f.nxt[hl2];
            R0 = f.dst[hr0]; R1 = f.dst[hr1]; R2 = f.dst[hr2]; R3 = f.dst[hr3];
            L0 = f.dst[hl0]; L1 = f.dst[hl1]; L2 = f.dst[hl2]; L3 = f.dst[hl3];
            bool res = false;
            if (gl == 0) { if (first_round) res = left_of(f, R0, li, ri); }                 // :556
            else if (gl == 1) { if (first_round) res = right_of(f, L0, ri, li); }           // :569
            else if (gl < 5) {                                                              // :559 chain
                const int k = gl - 2;
                res = in_circle(f, k == 0 ? R0 : (k == 1 ? R1 : R2), li, ri, k == 0 ? R1 : (k == 1 ? R2 : R3));
            } else if (gl < 8) {                                                            // :572 chain
                const int k = gl - 5;
                res = in_circle(f, li, ri, k == 0 ? L0 : (k == 1 ? L1 : L2), k == 0 ? L1 : (k == 1 ? L2 : L3));
            } else if (gl < 12) {                                                           // :588 for (a, b) in {0,1}^2
                const int a = (gl - 8) >> 1, b = (gl - 8) & 1;
                res = in_circle(f, li, ri, a ? R1 : R0, b ? L1 : L0);
            }
            const unsigned bal = (__ballot_sync(gmask, res) >> gshift) & 0xFFFFu;
            if (first_round) { vR = bal & 1u; vL = (bal >> 1) & 1u; }
            const unsigned cr = (bal >> 2) & 7u, cl = (bal >> 5) & 7u;
            nR = vR ? ((cr & 1u) ? ((cr & 2u) ? ((cr & 4u) ? 3 : 2) : 1) : 0) : 0;
            nL = vL ? ((cl & 1u) ? ((cl & 2u) ? ((cl & 4u) ? 3 : 2) : 1) : 0) : 0;
            haveF = (nR < 2 && nL < 2);
            F = (bal >> (8 + 2 * nR + nL)) & 1u;
            // ---- apply the deletions (cutVerts :277-282): twins in parallel, the two spliced rings by lanes 0 / 1
            if (gl >= 2 && gl < 5) {
                const int k = gl - 2;
                if (k < nR) { const int h = (k == 0 ? hr0 : (k == 1 ? hr1 : hr2)) ^ 1; ring_unlink(f, k == 0 ? R0 : (k == 1 ? R1 : R2), h); }
            } else if (gl >= 5 && gl < 8) {
                const int k = gl - 5;
                if (k < nL) { const int h = (k == 0 ? hl0 : (k == 1 ? hl1 : hl2)) ^ 1; ring_unlink(f, k == 0 ? L0 : (k == 1 ? L1 : L2), h); }
            } else if (gl == 0 && nR > 0) {
                const int surv = nR == 1 ? hr1 : (nR == 2 ? hr2 : hr3);
                f.nxt[surv] = (u16)(e ^ 1); f.prv[e ^ 1] = (u16)surv;
                const int hd = f.head[ri];
                if (hd == hr0 || (nR > 1 && hd == hr1) || (nR > 2 && hd == hr2)) f.head[ri] = (u16)(e ^ 1);
            } else if (gl == 1 && nL > 0) {
                const int surv = nL == 1 ? hl1 : (nL == 2 ? hl2 : hl3);
                f.nxt[e] = (u16)surv; f.prv[surv] = (u16)e;
                const int hd = f.head[li];
                if (hd == hl0 || (nL > 1 && hd == hl1) || (nL > 2 && hd == hl2)) f.head[li] = (u16)surv;
            }
            __syncwarp(gmask);
            if (gl == 0) {
                if (nR > 0) edge_free(f, ln, hr0 >> 1);
                if (nR > 1) edge_free(f, ln, hr1 >> 1);
                if (nR > 2) edge_free(f, ln, hr2 >> 1);
                if (nL > 0) edge_free(f, ln, hl0 >> 1);
                if (nL > 1) edge_free(f, ln, hl1 >> 1);
                if (nL > 2) edge_free(f, ln, hl2 >> 1);
            }
            __syncwarp(gmask);
            first_round = false;
            if (nR < 3 && nL < 3) break;
            // a chain used up its 3 speculated tests: look again from the survivors
            // (validity flags are kept: the reference tests them once per step)
        }
        // candidates that survived (relative to the last gather)
        const int hR = nR == 0 ? hr0 : (nR == 1 ? hr1 : (nR == 2 ? hr2 : hr3));
        const int r1 = nR == 0 ? R0 : (nR == 1 ? R1 : (nR == 2 ? R2 : R3));
        const int hL = nL == 0 ? hl0 : (nL == 1 ? hl1 : (nL == 2 ? hl2 : hl3));
        const int l1 = nL == 0 ? L0 : (nL == 1 ? L1 : (nL == 2 ? L2 : L3));
        // after a repeated round the speculated F belongs to that round's gather: still (r1, l1) if nR,nL < 2

        // ---- phase 2: insertion tests for both possible outcomes (+ the final choice if not speculated)
        //   ep0: parent l1, in ri   ep1: parent ri, in l1   (go left)
        //   ep2: parent li, in r1   ep3: parent r1, in li   (go right)
        bool res2 = false;
        if (gl < 8) {
            const int ep = gl >> 1;
            const int P = ep == 0 ? l1 : (ep == 1 ? ri : (ep == 2 ? li : r1));
            const int IN = ep == 0 ? ri : (ep == 1 ? l1 : (ep == 2 ? r1 : li));
            const int hd = f.head[P];
            const int hq = (gl & 1) ? f.prv[hd] : hd;
            res2 = right_of(f, IN, P, f.dst[hq]);
        } else if (gl == 8) {
            if (vR && vL && !haveF) res2 = in_circle(f, li, ri, r1, l1);
        }
        const unsigned bal2 = (__ballot_sync(gmask, res2) >> gshift) & 0xFFFFu;
        if (!haveF) F = (bal2 >> 8) & 1u;
        const bool go_left = !vR ? true : (!vL ? false : F);          // :582-593, ties prefer the right candidate
        int ne = 0;
        if (gl == 0) ne = edge_alloc(f, ln);
        ne = __shfl_sync(gmask, ne, 0, GT_G);
        if (ne < 0) return;
        if (go_left) {
            if (gl == 0) coop_insert(f, l1, ri, 2 * ne, hL ^ 1, bal2 & 1u, (bal2 >> 1) & 1u);
            else if (gl == 1) coop_insert(f, ri, l1, 2 * ne + 1, f.prv[e ^ 1], (bal2 >> 2) & 1u, (bal2 >> 3) & 1u);
            li = l1;
        } else {
            if (gl == 0) coop_insert(f, li, r1, 2 * ne, e, (bal2 >> 4) & 1u, (bal2 >> 5) & 1u);
            else if (gl == 1) coop_insert(f, r1, li, 2 * ne + 1, f.prv[hR ^ 1], (bal2 >> 6) & 1u, (bal2 >> 7) & 1u);
            ri = r1;
        }
        e = 2 * ne;
        __syncwarp(gmask);
    }
}

#endif  // __CUDACC__

}  // namespace gt
