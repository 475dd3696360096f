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
//      polynomial evaluated without rounding.  Branch-light, no local-memory expansions;
//      this replaces Shewchuk's adaptive expansion stages B-D, which are a poor fit for a
//      GPU (up to 2 x 1152-double scratch per thread, predicates.c:2674).
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

// ---------------------------------------------------------------- fixed-width integers
template <int N> struct Big { uint64_t w[N]; };   // two's complement, little-endian words

template <int N> GT_INLINE bool big_neg(const Big<N>& a) { return (int64_t)a.w[N - 1] < 0; }
template <int N> GT_INLINE bool big_zero(const Big<N>& a) {
    uint64_t o = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) o |= a.w[i];
    return o == 0;
}
template <int N> GT_INLINE Big<N> big_add(const Big<N>& a, const Big<N>& b) {
    Big<N> r; uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint64_t s = a.w[i] + b.w[i]; uint64_t c1 = s < a.w[i];
        uint64_t t = s + c; uint64_t c2 = t < s;
        r.w[i] = t; c = c1 | c2;
    }
    return r;
}
template <int N> GT_INLINE Big<N> big_negate(const Big<N>& a) {
    Big<N> r; uint64_t c = 1;
#pragma unroll
    for (int i = 0; i < N; ++i) { uint64_t t = ~a.w[i] + c; c = (t < c) ? 1 : 0; r.w[i] = t; }
    return r;
}
template <int N> GT_INLINE Big<N> big_sub(const Big<N>& a, const Big<N>& b) { return big_add(a, big_negate(b)); }
template <int N> GT_INLINE Big<N> big_abs(const Big<N>& a) { return big_neg(a) ? big_negate(a) : a; }

// signed N x N -> 2N words (schoolbook on magnitudes)
template <int N> GT_INLINE Big<2 * N> big_mul(const Big<N>& a, const Big<N>& b) {
    const bool neg = big_neg(a) != big_neg(b);
    const Big<N> ua = big_abs(a), ub = big_abs(b);
    Big<2 * N> r;
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) r.w[i] = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            uint64_t lo = ua.w[i] * ub.w[j], hi = mulhi64(ua.w[i], ub.w[j]);
            uint64_t s = r.w[i + j] + lo; hi += (s < lo);
            uint64_t t = s + carry; hi += (t < s);
            r.w[i + j] = t; carry = hi;
        }
        r.w[i + N] = carry;
    }
    return neg ? big_negate(r) : r;
}
template <int N> GT_INLINE int big_sign(const Big<N>& a) { return big_neg(a) ? -1 : (big_zero(a) ? 0 : 1); }

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
// v / 2^base as a W-word integer (exact when base <= lowbit(v) and the value fits).
template <int W> GT_INLINE Big<W> to_big(double v, int base) {
    Big<W> r;
#pragma unroll
    for (int i = 0; i < W; ++i) r.w[i] = 0;
    int64_t b = dbl_bits(v);
    uint64_t man = (uint64_t)b & 0x000FFFFFFFFFFFFFull;
    int ex = (int)((b >> 52) & 0x7FF);
    if (ex == 0 && man == 0) return r;
    uint64_t m = ex ? (man | 0x0010000000000000ull) : man;
    int sh = (ex ? ex - 1075 : -1074) - base;              // may be negative by at most ctz(m)
    if (sh < 0) { m >>= -sh; sh = 0; }
    int word = sh >> 6, bit = sh & 63;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (i == word) r.w[i] = m << bit;
        if (i == word + 1 && bit) r.w[i] = m >> (64 - bit);
    }
    return (b < 0) ? big_negate(r) : r;
}

// ---------------------------------------------------------------- exact stages
template <int W> GT_INLINE int orient2d_big(const double* c, int base) {
    // c = {ax, ay, bx, by, cx, cy};  det = (ax-cx)(by-cy) - (ay-cy)(bx-cx)
    Big<W> ax = to_big<W>(c[0], base), ay = to_big<W>(c[1], base), bx = to_big<W>(c[2], base),
           by = to_big<W>(c[3], base), cx = to_big<W>(c[4], base), cy = to_big<W>(c[5], base);
    Big<W> acx = big_sub(ax, cx), acy = big_sub(ay, cy), bcx = big_sub(bx, cx), bcy = big_sub(by, cy);
    return big_sign(big_sub(big_mul(acx, bcy), big_mul(acy, bcx)));
}
template <int W> GT_INLINE int incircle_big(const double* c, int base) {
    // c = {ax, ay, bx, by, cx, cy, dx, dy}
    Big<W> dx = to_big<W>(c[6], base), dy = to_big<W>(c[7], base);
    Big<W> adx = big_sub(to_big<W>(c[0], base), dx), ady = big_sub(to_big<W>(c[1], base), dy);
    Big<W> bdx = big_sub(to_big<W>(c[2], base), dx), bdy = big_sub(to_big<W>(c[3], base), dy);
    Big<W> cdx = big_sub(to_big<W>(c[4], base), dx), cdy = big_sub(to_big<W>(c[5], base), dy);
    Big<4 * W> det = big_mul(big_add(big_mul(adx, adx), big_mul(ady, ady)),
                             big_sub(big_mul(bdx, cdy), big_mul(cdx, bdy)));
    det = big_add(det, big_mul(big_add(big_mul(bdx, bdx), big_mul(bdy, bdy)),
                               big_sub(big_mul(cdx, ady), big_mul(adx, cdy))));
    det = big_add(det, big_mul(big_add(big_mul(cdx, cdx), big_mul(cdy, cdy)),
                               big_sub(big_mul(adx, bdy), big_mul(bdx, ady))));
    return big_sign(det);
}

// Exact sign of orient2d. Word count from the operands' bit spread: every coordinate needs
// spread+1 bits, differences one more, the degree-2 determinant 2*(spread+2)+1 <= 128*W - 1.
GT_NOINLINE int orient2d_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    double c[6] = {ax, ay, bx, by, cx, cy};
    Spread s; s.lo = 1 << 30; s.hi = -(1 << 30);
#pragma unroll
    for (int i = 0; i < 6; ++i) spread_acc(c[i], s);
    if (s.hi < s.lo) return 0;                             // all coordinates are zero
    int span = s.hi - s.lo + 1;
    if (span <= 59) return orient2d_big<1>(c, s.lo);
    if (span <= 123) return orient2d_big<2>(c, s.lo);
    if (span <= 251) return orient2d_big<4>(c, s.lo);
    return GT_RANGE_ERR;
}
// Exact sign of incircle: degree 4, 4*(spread+2)+3 <= 256*W - 1.
GT_NOINLINE int incircle_exact(double ax, double ay, double bx, double by, double cx, double cy,
                               double dx, double dy) {
    double c[8] = {ax, ay, bx, by, cx, cy, dx, dy};
    Spread s; s.lo = 1 << 30; s.hi = -(1 << 30);
#pragma unroll
    for (int i = 0; i < 8; ++i) spread_acc(c[i], s);
    if (s.hi < s.lo) return 0;
    int span = s.hi - s.lo + 1;
    if (span <= 59) return incircle_big<1>(c, s.lo);
    if (span <= 123) return incircle_big<2>(c, s.lo);
    if (span <= 251) return incircle_big<4>(c, s.lo);
    return GT_RANGE_ERR;
}

// ---------------------------------------------------------------- public predicates
GT_INLINE int orient2d_sign(double ax, double ay, double bx, double by, double cx, double cy) {
    // stage A: reference predicates.c:1628-1651, written without branches (lanes of a warp take this path
    // together): the sign of det is certain when the two products do not have the same strict sign, or when
    // |det| reaches the error bound
    const double detleft = (ax - cx) * (by - cy);
    const double detright = (ay - cy) * (bx - cx);
    const double det = detleft - detright;
    const bool same_sign = (detleft > 0.0 && detright > 0.0) || (detleft < 0.0 && detright < 0.0);
    const double errbound = GT_CCW_ERRBOUND_A * (fabs(detleft) + fabs(detright));
    if (!same_sign || fabs(det) >= errbound) return det > 0.0 ? 1 : (det < 0.0 ? -1 : 0);
    return orient2d_exact(ax, ay, bx, by, cx, cy);
}

GT_INLINE int incircle_sign(double ax, double ay, double bx, double by, double cx, double cy,
                            double dx, double dy) {
    // stage A: reference predicates.c:3238-3267
    double adx = ax - dx, bdx = bx - dx, cdx = cx - dx;
    double ady = ay - dy, bdy = by - dy, cdy = cy - dy;
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, alift = adx * adx + ady * ady;
    double cdxady = cdx * ady, adxcdy = adx * cdy, blift = bdx * bdx + bdy * bdy;
    double adxbdy = adx * bdy, bdxady = bdx * ady, clift = cdx * cdx + cdy * cdy;
    double det = alift * (bdxcdy - cdxbdy) + blift * (cdxady - adxcdy) + clift * (adxbdy - bdxady);
    double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * alift + (fabs(cdxady) + fabs(adxcdy)) * blift
                     + (fabs(adxbdy) + fabs(bdxady)) * clift;
    double errbound = GT_ICC_ERRBOUND_A * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return incircle_exact(ax, ay, bx, by, cx, cy, dx, dy);
}

}  // namespace gt
