/* oracle/oracle_predicates.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the two predicates the reference triangulator consumes:
 *   orient2d  (reference extern/predicates/predicates.c:1620-1654, adapt stage :1536-1618)
 *   incircle  (reference extern/predicates/predicates.c:3226-3270, adapt stage :2652-3224)
 * Every use in the reference is a comparison with 0.0 (src/delaunay_tri.c:138,186), and
 * Shewchuk's routines guarantee the sign of the returned value is the sign of the exact
 * determinant.  This restatement therefore keeps the reference's stage-A floating-point
 * filter (same error-bound constants as exactinit, predicates.c:671-680) and replaces the
 * adaptive stages B-D by one straightforward exact evaluation with non-overlapping
 * expansions (the GROW-/SCALE-EXPANSION algorithms of Shewchuk's 1996 report, written
 * from the published description with fma() as the error-free product).  The returned
 * value has the exact sign; its magnitude past stage A is only an estimate.
 * It is deliberately a different exact method from the device code (multi-word integers),
 * so that the three implementations (reference, this port, CUDA) check one another. */
#include <math.h>
#include <string.h>
#include "oracle_port.h"

static const double EPS = 1.1102230246251565e-16;            /* 2^-53, predicates.c:644-670 */
#define CCW_ERRBOUND_A ((3.0 + 16.0 * EPS) * EPS)            /* predicates.c:673 */
#define ICC_ERRBOUND_A ((10.0 + 96.0 * EPS) * EPS)           /* predicates.c:678 */

static int g_count = 0;   /* counters are for single-threaded analysis runs only */
void port_pred_count_enable(int on) { g_count = on; }
static long long g_orient_exact = 0, g_incircle_exact = 0, g_orient_calls = 0, g_incircle_calls = 0;
void port_pred_counters(long long *out4, int reset) {
    out4[0] = g_orient_calls; out4[1] = g_orient_exact; out4[2] = g_incircle_calls; out4[3] = g_incircle_exact;
    if (reset) g_orient_calls = g_orient_exact = g_incircle_calls = g_incircle_exact = 0;
}

/* ---- error-free transformations ---- */
static inline void two_sum(double a, double b, double *x, double *y) {
    double s = a + b, bv = s - a, av = s - bv;
    *x = s; *y = (a - av) + (b - bv);
}
static inline void fast_two_sum(double a, double b, double *x, double *y) { /* |a| >= |b| */
    double s = a + b; *x = s; *y = b - (s - a);
}
static inline void two_prod(double a, double b, double *x, double *y) {
    double p = a * b; *x = p; *y = fma(a, b, -p);
}

/* Expansion = array of doubles, non-overlapping, increasing magnitude, zeros removed.
 * MAXE bounds the longest intermediate of the incircle evaluation (3 * 2*16*16). */
#define MAXE 1600

/* h = e + b (GROW-EXPANSION with zero elimination); returns length of h. h may alias e. */
static int grow(int n, const double *e, double b, double *h) {
    double q = b, hh; int m = 0;
    for (int i = 0; i < n; ++i) {
        double qn; two_sum(q, e[i], &qn, &hh); q = qn;
        if (hh != 0.0) h[m++] = hh;
    }
    if (q != 0.0 || m == 0) h[m++] = q;
    return m;
}
/* h = e + f, by growing one component at a time (EXPANSION-SUM). h must not alias f. */
static int esum(int ne, const double *e, int nf, const double *f, double *h) {
    int n = ne;
    if (h != e) memcpy(h, e, sizeof(double) * (size_t)ne);
    for (int i = 0; i < nf; ++i) n = grow(n, h, f[i], h);
    return n;
}
/* h = e * b (SCALE-EXPANSION with zero elimination). h must not alias e. */
static int escale(int n, const double *e, double b, double *h) {
    double q, hh, t1, t0, s; int m = 0;
    if (n == 0) { h[0] = 0.0; return 1; }
    two_prod(e[0], b, &q, &hh);
    if (hh != 0.0) h[m++] = hh;
    for (int i = 1; i < n; ++i) {
        two_prod(e[i], b, &t1, &t0);
        two_sum(q, t0, &s, &hh);
        if (hh != 0.0) h[m++] = hh;
        fast_two_sum(t1, s, &q, &hh);
        if (hh != 0.0) h[m++] = hh;
    }
    if (q != 0.0 || m == 0) h[m++] = q;
    return m;
}
/* h = e * f */
static int emul(int ne, const double *e, int nf, const double *f, double *h) {
    double tmp[MAXE]; int n = 1; h[0] = 0.0;
    for (int i = 0; i < ne; ++i) {
        int nt = escale(nf, f, e[i], tmp);
        n = esum(n, h, nt, tmp, h);
    }
    return n;
}
static void eneg(int n, double *e) { for (int i = 0; i < n; ++i) e[i] = -e[i]; }
static double esign_value(int n, const double *e) {     /* most significant component */
    for (int i = n - 1; i >= 0; --i) if (e[i] != 0.0) return e[i];
    return 0.0;
}
/* exact a - b as a 2-expansion */
static int ediff2(double a, double b, double *h) {
    double x, y; two_sum(a, -b, &x, &y);
    int m = 0; if (y != 0.0) h[m++] = y; if (x != 0.0 || m == 0) h[m++] = x; return m;
}

static double orient2d_exact(const double *pa, const double *pb, const double *pc) {
    double acx[2], acy[2], bcx[2], bcy[2], l[16], r[16], d[40];
    int nacx = ediff2(pa[0], pc[0], acx), nacy = ediff2(pa[1], pc[1], acy);
    int nbcx = ediff2(pb[0], pc[0], bcx), nbcy = ediff2(pb[1], pc[1], bcy);
    int nl = emul(nacx, acx, nbcy, bcy, l);
    int nr = emul(nacy, acy, nbcx, bcx, r);
    eneg(nr, r);
    int nd = esum(nl, l, nr, r, d);
    return esign_value(nd, d);
}

double port_orient2d(const double *pa, const double *pb, const double *pc) {
    /* stage A, reference predicates.c:1628-1651 */
    double detleft = (pa[0] - pc[0]) * (pb[1] - pc[1]);
    double detright = (pa[1] - pc[1]) * (pb[0] - pc[0]);
    double det = detleft - detright, detsum;
if (g_count) g_orient_calls++;
    if (detleft > 0.0) { if (detright <= 0.0) return det; detsum = detleft + detright; }
    else if (detleft < 0.0) { if (detright >= 0.0) return det; detsum = -detleft - detright; }
    else return det;
    double errbound = CCW_ERRBOUND_A * detsum;
    if (det >= errbound || -det >= errbound) return det;
if (g_count) g_orient_exact++;
    return orient2d_exact(pa, pb, pc);
}

static double incircle_exact(const double *pa, const double *pb, const double *pc, const double *pd) {
    /* det = alift*(bdx*cdy - cdx*bdy) + blift*(cdx*ady - adx*cdy) + clift*(adx*bdy - bdx*ady)
     * with every difference carried exactly as a 2-expansion. */
    double dx[3][2], dy[3][2]; int ndx[3], ndy[3];
    const double *p[3] = {pa, pb, pc};
    for (int i = 0; i < 3; ++i) { ndx[i] = ediff2(p[i][0], pd[0], dx[i]); ndy[i] = ediff2(p[i][1], pd[1], dy[i]); }
    static _Thread_local double total[MAXE], term[MAXE], lift[64], minor[64], t1[32], t2[32];
    int ntotal = 1; total[0] = 0.0;
    for (int i = 0; i < 3; ++i) {
        int j = (i + 1) % 3, k = (i + 2) % 3;
        int n1 = emul(ndx[j], dx[j], ndy[k], dy[k], t1);
        int n2 = emul(ndx[k], dx[k], ndy[j], dy[j], t2);
        eneg(n2, t2);
        int nminor = esum(n1, t1, n2, t2, minor);
        n1 = emul(ndx[i], dx[i], ndx[i], dx[i], t1);
        n2 = emul(ndy[i], dy[i], ndy[i], dy[i], t2);
        int nlift = esum(n1, t1, n2, t2, lift);
        int nterm = emul(nlift, lift, nminor, minor, term);
        ntotal = esum(ntotal, total, nterm, term, total);
    }
    return esign_value(ntotal, total);
}

double port_incircle(const double *pa, const double *pb, const double *pc, const double *pd) {
    /* stage A, reference predicates.c:3238-3267 */
    double adx = pa[0] - pd[0], bdx = pb[0] - pd[0], cdx = pc[0] - pd[0];
    double ady = pa[1] - pd[1], bdy = pb[1] - pd[1], cdy = pc[1] - pd[1];
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, alift = adx * adx + ady * ady;
    double cdxady = cdx * ady, adxcdy = adx * cdy, blift = bdx * bdx + bdy * bdy;
    double adxbdy = adx * bdy, bdxady = bdx * ady, clift = cdx * cdx + cdy * cdy;
    double det = alift * (bdxcdy - cdxbdy) + blift * (cdxady - adxcdy) + clift * (adxbdy - bdxady);
    double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * alift + (fabs(cdxady) + fabs(adxcdy)) * blift
                     + (fabs(adxbdy) + fabs(bdxady)) * clift;
    double errbound = ICC_ERRBOUND_A * permanent;
if (g_count) g_incircle_calls++;
    if (det > errbound || -det > errbound) return det;
if (g_count) g_incircle_exact++;
    return incircle_exact(pa, pb, pc, pd);
}

void port_orient2d_batch(const double *p, int n, double *out) {
    for (int i = 0; i < n; ++i) out[i] = port_orient2d(p + 6 * i, p + 6 * i + 2, p + 6 * i + 4);
}
void port_incircle_batch(const double *p, int n, double *out) {
    for (int i = 0; i < n; ++i) out[i] = port_incircle(p + 8 * i, p + 8 * i + 2, p + 8 * i + 4, p + 8 * i + 6);
}
