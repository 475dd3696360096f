// gt_predicates.cuh -- sign-exact orient2d / incircle for sm_100a.
//
// Replaces, for the device, the two predicates the reference triangulator consumes:
//   orient2d  reference extern/predicates/predicates.c:1620-1654 (+ orient2dadapt :1536-1618)
//   incircle  reference extern/predicates/predicates.c:3226-3270 (+ incircleadapt :2652-3224)
// Every use in the reference is a strict comparison with 0.0 (src/delaunay_tri.c:138,186) and
// Shewchuk's routines return a value whose SIGN is that of the exact determinant, so the
// contract here is: return the exact sign (-1, 0, +1).  Any exact method then reproduces the
// reference's decisions, including its tie-breaking on collinear / cocircular input.
//
// Stages:
//   A  the reference's own FP64 filter (same error-bound constants as exactinit,
//      predicates.c:671-680), evaluated with every operation individually rounded -- this
//      translation unit is compiled with -fmad=false so nvcc cannot contract a*b+c.
//   X  exact evaluation in multi-word two's-complement integers.  The coordinates of one
//      call are scaled by a common power of two to integers of W 64-bit words (W = 1, 2, 4
//      chosen from the bit spread of the operands), the determinant is a degree-2 / degree-4
//      polynomial evaluated without rounding.  This replaces Shewchuk's adaptive expansion
//      stages B-D, which are a poor fit for a GPU (up to 2 x 1152-double scratch per thread,
//      predicates.c:2674); the integers take at most 0.8 KB of local memory and the stage is
//      out of line and deliberately register-lean (see "multi-word integers" below).
//      Operands whose bit spread exceeds 251 bits (e.g. 1e-40 next to 1e+20) return
//      GT_RANGE_ERR instead of a sign (by value: an out-parameter would force a local-memory
//      round trip on every call); the frame is then flagged GTA_B200_ST_PRECISION_RANGE.
//
// The file is also compilable as plain host C++ (tests/emu only) so that the kernel logic
// can be exercised without a GPU; the product library contains the device build only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GT_HD __host__ __device__
#define GT_INLINE __host__ __device__ __forceinline__
#define GT_NOINLINE static __host__ __device__ __noinline__
#else
#define GT_HD
#define GT_INLINE inline
#define GT_NOINLINE static __attribute__((noinline))
#include <math.h>
#include <string.h>
#endif

namespace gt {

// ---------------------------------------------------------------- constants (exactinit)
// epsilon = 2^-53; ccwerrboundA = (3 + 16 eps) eps; iccerrboundA = (10 + 96 eps) eps.
#define GT_EPS 1.1102230246251565e-16
#define GT_CCW_ERRBOUND_A ((3.0 + 16.0 * GT_EPS) * GT_EPS)
#define GT_ICC_ERRBOUND_A ((10.0 + 96.0 * GT_EPS) * GT_EPS)
#define GT_RANGE_ERR 2      // returned instead of a sign when the exact stage cannot represent the operands

GT_INLINE int64_t dbl_bits(double v) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    int64_t b; memcpy(&b, &v, 8); return b;
#endif
}
GT_INLINE int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
GT_INLINE int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return x ? __builtin_ctzll(x) : 64;
#endif
}
GT_INLINE uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// ---------------------------------------------------------------- multi-word integers
// Little-endian two's-complement integers of n 64-bit words. They live in local memory and the helpers are out of
// line with run-time loops ON PURPOSE: the exact stage runs for none of the predicates of cfg 1-3 and ~1 % of cfg 5's
// (tests/emu counters), but a register-resident, fully unrolled version costs every caller up to 140 registers at its
// call site (measured: the lane-per-frame incircle phase needed 250 registers, 108 without the call) -- and that
// halves the warps an SM can hold. Slow and small is the right trade for a path this rare.

// r = a + b  or  a - b  (n words; r may alias a or b)
GT_NOINLINE void xi_addsub(uint64_t* r, const uint64_t* a, const uint64_t* b, int n, bool sub) {
    uint64_t c = sub ? 1 : 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const uint64_t x = a[i], y = sub ? ~b[i] : b[i];
        const uint64_t s = x + y; const uint64_t c1 = s < x;
        const uint64_t t = s + c; const uint64_t c2 = t < s;
        r[i] = t; c = c1 | c2;
    }
}
// r[0..2n) = a * b, signed (r must not alias a or b). Unsigned schoolbook product, then the two's-complement
// correction: with ua = a + 2^(64n)[a<0], a*b = ua*ub - 2^(64n) ([a<0] ub + [b<0] ua)  (mod 2^(128n)).
GT_NOINLINE void xi_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, int n) {
#pragma unroll 1
    for (int i = 0; i < 2 * n; ++i) r[i] = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        uint64_t carry = 0; const uint64_t ai = a[i];
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const uint64_t lo = ai * b[j]; uint64_t hi = mulhi64(ai, b[j]);
            const uint64_t s = r[i + j] + lo; hi += (s < lo);
            const uint64_t t = s + carry; hi += (t < s);
            r[i + j] = t; carry = hi;
        }
        r[i + n] = carry;
    }
    if ((int64_t)a[n - 1] < 0) xi_addsub(r + n, r + n, b, n, true);
    if ((int64_t)b[n - 1] < 0) xi_addsub(r + n, r + n, a, n, true);
}
GT_NOINLINE int xi_sign(const uint64_t* a, int n) {
    if ((int64_t)a[n - 1] < 0) return -1;
    uint64_t o = 0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) o |= a[i];
    return o ? 1 : 0;
}

// ---------------------------------------------------------------- double -> scaled integer
// A finite double v != 0 is m * 2^e with integer m (53 bits). lowbit/highbit are the
// exponents of its lowest and highest set bits.
struct Spread { int lo, hi; };
GT_INLINE void spread_acc(double v, Spread& s) {
    int64_t b = dbl_bits(v);
    uint64_t man = (uint64_t)b & 0x000FFFFFFFFFFFFFull;
    int ex = (int)((b >> 52) & 0x7FF);
    if (ex == 0 && man == 0) return;                       // +-0 contributes nothing
    uint64_t m = ex ? (man | 0x0010000000000000ull) : man;
    int e = ex ? ex - 1075 : -1074;
    int lo = e + ctz64(m), hi = e + 63 - clz64(m);
    s.lo = lo < s.lo ? lo : s.lo;
    s.hi = hi > s.hi ? hi : s.hi;
}
// Device-side count of the predicates the FP filter left undecided (settled by stage A' or by the integers): the host
// reads it after a launch to learn whether the input is of the degenerate, few-bit kind that stage A' pays for
// (gta_b200.cu, launch_lpf).
#if defined(__CUDACC__)
static __device__ unsigned long long g_xcalls;
#endif
#if defined(__CUDA_ARCH__)
#define GT_XCOUNT() do { if ((threadIdx.x & 7) == 0) atomicAdd(&g_xcalls, 1ull); } while (0)        // a 1-in-8 sample: it is one address
#else
#define GT_XCOUNT() ((void)0)
#endif

// Common scale and word count of one call's operands: base = lowest set bit over all of them, n = 1, 2 or 4 words
// per coordinate from their bit spread (0: they do not fit, -1: all operands are zero).
GT_NOINLINE int xi_plan(const double* c, int count, int* base) {
    GT_XCOUNT();                                           // (here: this function is the first thing both integer stages call)
    Spread s; s.lo = 1 << 30; s.hi = -(1 << 30);
#pragma unroll 1
    for (int i = 0; i < count; ++i) spread_acc(c[i], s);
    if (s.hi < s.lo) return -1;
    *base = s.lo;
    const int span = s.hi - s.lo + 1;
    return span <= 59 ? 1 : (span <= 123 ? 2 : (span <= 251 ? 4 : 0));
}
// r = v / 2^base as an n-word integer (exact when base <= lowbit(v) and the value fits).
GT_NOINLINE void xi_set(uint64_t* r, int n, double v, int base) {
#pragma unroll 1
    for (int i = 0; i < n; ++i) r[i] = 0;
    const int64_t b = dbl_bits(v);
    const uint64_t man = (uint64_t)b & 0x000FFFFFFFFFFFFFull;
    const int ex = (int)((b >> 52) & 0x7FF);
    if (ex == 0 && man == 0) return;
    uint64_t m = ex ? (man | 0x0010000000000000ull) : man;
    int sh = (ex ? ex - 1075 : -1074) - base;              // may be negative by at most ctz(m)
    if (sh < 0) { m >>= -sh; sh = 0; }
    const int word = sh >> 6, bit = sh & 63;
    if (word < n) r[word] = m << bit;
    if (bit && word + 1 < n) r[word + 1] = m >> (64 - bit);
    if (b < 0) {                                           // negate: ~r + 1
        uint64_t c = 1;
#pragma unroll 1
        for (int i = 0; i < n; ++i) { const uint64_t t = ~r[i] + c; c = (t < c) ? 1 : 0; r[i] = t; }
    }
}

// Host-only counters of the exact stages taken (emulation build, -DGT_STATS): [0..2] orient2d with 1/2/4 words, [3..5] incircle.
#if defined(GT_STATS) && !defined(__CUDA_ARCH__)
extern long g_exact_calls[8];
#define GT_XSTAT(i) (++g_exact_calls[i])
#else
#define GT_XSTAT(i) ((void)0)
#endif

// ---------------------------------------------------------------- exact stages
// Exact sign of orient2d, det = (ax-cx)(by-cy) - (ay-cy)(bx-cx). Word count from the operands' bit spread: every
// coordinate needs spread+1 bits, differences one more, the degree-2 determinant 2*(spread+2)+1 <= 128*n - 1.
GT_NOINLINE int orient2d_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    double c[6] = {ax, ay, bx, by, cx, cy};
    int base = 0;
    const int n = xi_plan(c, 6, &base);
    if (n < 0) return 0;                                   // all coordinates are zero
    if (n == 0) return GT_RANGE_ERR;
    GT_XSTAT(n >> 1);
    uint64_t v[4][4], o[2][4], m1[8], m2[8];               // v = acx, acy, bcx, bcy
    xi_set(o[0], n, c[4], base); xi_set(o[1], n, c[5], base);
#pragma unroll 1
    for (int i = 0; i < 4; ++i) { xi_set(v[i], n, c[i], base); xi_addsub(v[i], v[i], o[i & 1], n, true); }
    xi_mul(m1, v[0], v[3], n); xi_mul(m2, v[1], v[2], n);
    xi_addsub(m1, m1, m2, 2 * n, true);
    return xi_sign(m1, 2 * n);
}
// Exact sign of incircle: degree 4, 4*(spread+2)+3 <= 256*n - 1.
//   det = alift (bdx cdy - cdx bdy) + blift (cdx ady - adx cdy) + clift (adx bdy - bdx ady),  alift = adx^2 + ady^2, ...
GT_NOINLINE int incircle_exact(double ax, double ay, double bx, double by, double cx, double cy,
                               double dx, double dy) {
    double c[8] = {ax, ay, bx, by, cx, cy, dx, dy};
    int base = 0;
    const int n = xi_plan(c, 8, &base);
    if (n < 0) return 0;
    if (n == 0) return GT_RANGE_ERR;
    GT_XSTAT(3 + (n >> 1));
    uint64_t v[6][4], o[2][4];                             // v = adx, ady, bdx, bdy, cdx, cdy
    uint64_t lift[8], m1[8], m2[8], term[16], det[16];
    xi_set(o[0], n, c[6], base); xi_set(o[1], n, c[7], base);
#pragma unroll 1
    for (int i = 0; i < 6; ++i) { xi_set(v[i], n, c[i], base); xi_addsub(v[i], v[i], o[i & 1], n, true); }
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
        const int p = (t + 1) % 3, q = (t + 2) % 3;
        xi_mul(lift, v[2 * t], v[2 * t], n); xi_mul(m1, v[2 * t + 1], v[2 * t + 1], n);
        xi_addsub(lift, lift, m1, 2 * n, false);
        xi_mul(m1, v[2 * p], v[2 * q + 1], n); xi_mul(m2, v[2 * q], v[2 * p + 1], n);
        xi_addsub(m1, m1, m2, 2 * n, true);
        xi_mul(t == 0 ? det : term, lift, m1, 2 * n);
        if (t) xi_addsub(det, det, term, 4 * n, false);
    }
    return xi_sign(det, 4 * n);
}

// ---------------------------------------------------------------- public predicates
// Stage A' (between the filter and the integers): was every operation of the filter exact for these operands? Then the
// filter's det IS the determinant and its sign -- including 0 -- is final. True for few-bit coordinates, i.e. exactly
// the inputs that are degenerate on purpose (lattices: cfg 5 has ~320 undecided predicates per frame, all of this
// kind); costs ~100 flops on the undecided path only, but ~15 registers at every inlined call site, so it is a template
// switch (AP): the lane-per-frame kernel is built both ways and the host picks (gta_b200.cu, launch_lpf). Error-free
// transformations as in predicates.c:171-239 (Two_Sum / Two_Diff tails, and the product's tail by one FMA instead of
// Split + Two_Product).
GT_INLINE double gt_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
GT_INLINE bool exact_diff(double a, double b, double x) {          // x == fl(a - b)
    const double bvirt = a - x, avirt = x + bvirt, bround = bvirt - b, around = a - avirt;
    return (around + bround) == 0.0;
}
GT_INLINE bool exact_sum(double a, double b, double x) {           // x == fl(a + b)
    const double bvirt = x - a, avirt = x - bvirt, bround = b - bvirt, around = a - avirt;
    return (around + bround) == 0.0;
}
GT_INLINE bool exact_prod(double a, double b, double p) {          // p == fl(a * b); tiny products are "not proven": their tail may underflow
    return (p == 0.0 || fabs(p) > 1e-200) && gt_fma(a, b, -p) == 0.0;
}

template <bool AP = true>
GT_INLINE int orient2d_sign(double ax, double ay, double bx, double by, double cx, double cy) {
    // stage A: reference predicates.c:1628-1651, written without branches (lanes of a warp take this path
    // together): the sign of det is certain when the two products do not have the same strict sign, or when
    // |det| reaches the error bound
    const double acx = ax - cx, bcy = by - cy, acy = ay - cy, bcx = bx - cx;
    const double detleft = acx * bcy;
    const double detright = acy * bcx;
    const double det = detleft - detright;
    const bool same_sign = (detleft > 0.0 && detright > 0.0) || (detleft < 0.0 && detright < 0.0);
    const double errbound = GT_CCW_ERRBOUND_A * (fabs(detleft) + fabs(detright));
    if (!same_sign || fabs(det) >= errbound) return det > 0.0 ? 1 : (det < 0.0 ? -1 : 0);
    const bool exact = AP & exact_diff(ax, cx, acx) & exact_diff(by, cy, bcy) & exact_diff(ay, cy, acy) & exact_diff(bx, cx, bcx)
                     & exact_prod(acx, bcy, detleft) & exact_prod(acy, bcx, detright) & exact_diff(detleft, detright, det);
    if (exact) { GT_XCOUNT(); return det > 0.0 ? 1 : (det < 0.0 ? -1 : 0); }
    return orient2d_exact(ax, ay, bx, by, cx, cy);
}

template <bool AP = true>
GT_INLINE int incircle_sign(double ax, double ay, double bx, double by, double cx, double cy,
                            double dx, double dy) {
    // stage A: reference predicates.c:3238-3267
    const double adx = ax - dx, bdx = bx - dx, cdx = cx - dx;
    const double ady = ay - dy, bdy = by - dy, cdy = cy - dy;
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, adx2 = adx * adx, ady2 = ady * ady, alift = adx2 + ady2;
    const double cdxady = cdx * ady, adxcdy = adx * cdy, bdx2 = bdx * bdx, bdy2 = bdy * bdy, blift = bdx2 + bdy2;
    const double adxbdy = adx * bdy, bdxady = bdx * ady, cdx2 = cdx * cdx, cdy2 = cdy * cdy, clift = cdx2 + cdy2;
    const double ma = bdxcdy - cdxbdy, mb = cdxady - adxcdy, mc = adxbdy - bdxady;
    const double ta = alift * ma, tb = blift * mb, tc = clift * mc, tab = ta + tb;
    const double det = tab + tc;
    const double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * alift + (fabs(cdxady) + fabs(adxcdy)) * blift
                           + (fabs(adxbdy) + fabs(bdxady)) * clift;
    const double errbound = GT_ICC_ERRBOUND_A * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    bool exact = AP & exact_diff(ax, dx, adx) & exact_diff(bx, dx, bdx) & exact_diff(cx, dx, cdx)
               & exact_diff(ay, dy, ady) & exact_diff(by, dy, bdy) & exact_diff(cy, dy, cdy);
    exact = exact & exact_prod(bdx, cdy, bdxcdy) & exact_prod(cdx, bdy, cdxbdy) & exact_prod(cdx, ady, cdxady)
          & exact_prod(adx, cdy, adxcdy) & exact_prod(adx, bdy, adxbdy) & exact_prod(bdx, ady, bdxady);
    exact = exact & exact_prod(adx, adx, adx2) & exact_prod(ady, ady, ady2) & exact_prod(bdx, bdx, bdx2)
          & exact_prod(bdy, bdy, bdy2) & exact_prod(cdx, cdx, cdx2) & exact_prod(cdy, cdy, cdy2);
    exact = exact & exact_sum(adx2, ady2, alift) & exact_sum(bdx2, bdy2, blift) & exact_sum(cdx2, cdy2, clift)
          & exact_diff(bdxcdy, cdxbdy, ma) & exact_diff(cdxady, adxcdy, mb) & exact_diff(adxbdy, bdxady, mc);
    exact = exact & exact_prod(alift, ma, ta) & exact_prod(blift, mb, tb) & exact_prod(clift, mc, tc)
          & exact_sum(ta, tb, tab) & exact_sum(tab, tc, det);
    if (exact) { GT_XCOUNT(); return det > 0.0 ? 1 : (det < 0.0 ? -1 : 0); }
    return incircle_exact(ax, ay, bx, by, cx, cy, dx, dy);
}

}  // namespace gt
