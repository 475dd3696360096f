/* oracle/oracle_port.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, index-based restatement of the reference hot path:
 *   delaunay_tessellate / delaunay_surface_area  (reference src/gta_tri.c:80-329, :332-413)
 *   dtriangulate and its helpers                 (reference src/delaunay_tri.c:115-123, :210-635)
 *   area_tri                                     (reference include/gta_tri.h:91-99)
 * It keeps the reference's decision sequence (same sort order, same split index
 * mid=(ia+ib)/2, same strict "> 0" predicate tests, same ring-insertion scan, same
 * triangle emission order) but none of its data structures: vertices and ring nodes are
 * integers into flat arrays, there is one allocation per frame instead of one per
 * half-edge.  Parity status: PINNED against the unmodified reference compiled into
 * oracle/_ref (tests/test_oracle.py, tests/golden/), because the reference ships no
 * golden vectors of its own.
 *
 * GROMACS (un-vendored dependency of the reference; `real`, rvec math) is restated with
 * real = double: cprod/norm as documented for gromacs/legacyheaders/vec.h. */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_port.h"

#define NIL (-1)
#define DT_EPS 1e-12                       /* reference delaunay_tri.c:30 */

typedef struct {
    int n;                                 /* number of vertices */
    const double *pts;                     /* caller's interleaved x,y */
    int *ord;                              /* sorted position -> input index */
    int *head;                             /* ring entry point per vertex ("first"), NIL if none */
    int *nvert, *nprev, *nnext;            /* ring nodes: neighbour vertex, CW link, CCW link */
    int nfree, nused, cap;
} dt_t;

static inline const double *P(const dt_t *t, int v) { return t->pts + 2 * (size_t)t->ord[v]; }

/* reference delaunay_tri.c:126-187: every predicate use is a strict sign test */
static inline int ccw(const dt_t *t, int a, int b, int c) { return port_orient2d(P(t, a), P(t, b), P(t, c)) > 0.0; }
static inline int right_of(const dt_t *t, int x, int ea, int eb) { return ccw(t, x, eb, ea); }
static inline int left_of(const dt_t *t, int x, int ea, int eb) { return ccw(t, x, ea, eb); }
static inline int in_circle(const dt_t *t, int a, int b, int c, int d) {
    return port_incircle(P(t, a), P(t, b), P(t, c), P(t, d)) > 0.0;
}

static int node_new(dt_t *t, int v) {
    int k;
    if (t->nfree != NIL) { k = t->nfree; t->nfree = t->nnext[k]; }
    else {
        if (t->nused == t->cap) { fprintf(stderr, "oracle_port: node pool exhausted\n"); abort(); }
        k = t->nused++;
    }
    t->nvert[k] = v;
    return k;
}
static void node_free(dt_t *t, int k) { t->nnext[k] = t->nfree; t->nfree = k; }

static inline void link_after(dt_t *t, int at, int k) {    /* reference :202-208 */
    int nx = t->nnext[at];
    t->nnext[at] = k; t->nprev[nx] = k; t->nprev[k] = at; t->nnext[k] = nx;
}

/* Put vertex `in` into the CCW ring of `parent` (reference insertNode, :210-245). */
static void ring_insert(dt_t *t, int parent, int in) {
    int k = node_new(t, in);
    int h = t->head[parent];
    if (h == NIL) { t->head[parent] = k; t->nprev[k] = k; t->nnext[k] = k; return; }
    if (right_of(t, in, parent, t->nvert[h])) {
        int cur = t->nprev[h];
        while (cur != h && right_of(t, in, parent, t->nvert[cur])) cur = t->nprev[cur];
        if (cur == h) {                      /* new hull successor becomes "first" */
            t->head[parent] = k;
            link_after(t, t->nprev[cur], k);
        } else {
            link_after(t, cur, k);
        }
    } else {
        int cur = t->nnext[h];
        while (cur != h && left_of(t, in, parent, t->nvert[cur])) cur = t->nnext[cur];
        if (t->nvert[cur] == in) { node_free(t, k); return; }    /* already a neighbour */
        link_after(t, t->nprev[cur], k);
    }
}

/* Remove `child` from the ring of `parent` (reference deleteNode, :247-268). */
static void ring_delete(dt_t *t, int parent, int child) {
    int h = t->head[parent];
    if (h == NIL) return;
    int k = h;
    do {
        if (t->nvert[k] == child) {
            t->nnext[t->nprev[k]] = t->nnext[k];
            t->nprev[t->nnext[k]] = t->nprev[k];
            if (k == h) t->head[parent] = (t->nnext[k] == k) ? NIL : t->nnext[k];
            node_free(t, k);
            return;
        }
        k = t->nnext[k];
    } while (k != h);
}

static void connect(dt_t *t, int a, int b) { if (a != NIL && b != NIL && a != b) { ring_insert(t, a, b); ring_insert(t, b, a); } }
static void cut(dt_t *t, int a, int b) { if (a != NIL && b != NIL && a != b) { ring_delete(t, a, b); ring_delete(t, b, a); } }

static inline int first_of(const dt_t *t, int v) {          /* reference :285-289 */
    return (v != NIL && t->head[v] != NIL) ? t->nvert[t->head[v]] : NIL;
}
static int pred_of(const dt_t *t, int vi, int vj) {           /* CW neighbour, reference :291-311 */
    if (vi == NIL || vj == NIL) return NIL;
    int h = t->head[vi], k = h;
    if (h == NIL) return NIL;
    do { if (t->nvert[k] == vj) return t->nvert[t->nprev[k]]; k = t->nprev[k]; } while (k != h);
    return NIL;
}
static int succ_of(const dt_t *t, int vi, int vj) {           /* CCW neighbour, reference :313-333 */
    if (vi == NIL || vj == NIL) return NIL;
    int h = t->head[vi], k = h;
    if (h == NIL) return NIL;
    do { if (t->nvert[k] == vj) return t->nvert[t->nnext[k]]; k = t->nnext[k]; } while (k != h);
    return NIL;
}

/* Lower / upper common tangents of two x-separated hulls (reference lct :337-370, uct :373-406). */
static void lower_tangent(const dt_t *t, int lrightmost, int rleftmost, int *tl, int *tr) {
    int x = lrightmost, y = rleftmost;
    int rfast = first_of(t, y), lfast = NIL;
    int fx = first_of(t, x);
    if (fx != NIL) lfast = pred_of(t, x, fx);
    for (;;) {
        if (rfast != NIL && right_of(t, rfast, x, y)) { int tmp = rfast; rfast = succ_of(t, rfast, y); y = tmp; }
        else if (lfast != NIL && right_of(t, lfast, x, y)) { int tmp = lfast; lfast = pred_of(t, lfast, x); x = tmp; }
        else { *tl = x; *tr = y; return; }
    }
}
static void upper_tangent(const dt_t *t, int lrightmost, int rleftmost, int *tl, int *tr) {
    int x = lrightmost, y = rleftmost;
    int lfast = first_of(t, x), rfast = NIL;
    int fy = first_of(t, y);
    if (fy != NIL) rfast = pred_of(t, y, fy);
    for (;;) {
        if (rfast != NIL && left_of(t, rfast, x, y)) { int tmp = rfast; rfast = pred_of(t, rfast, y); y = tmp; }
        else if (lfast != NIL && left_of(t, lfast, x, y)) { int tmp = lfast; lfast = succ_of(t, lfast, x); x = tmp; }
        else { *tl = x; *tr = y; return; }
    }
}

/* Divide and conquer over sorted positions ia..ib (reference ord_dtriangulate, :511-601). */
static void build(dt_t *t, int ia, int ib) {
    int cnt = ib - ia + 1;
    if (cnt == 2) { connect(t, ia, ib); return; }
    if (cnt == 3) {
        connect(t, ia, ia + 1);
        connect(t, ia + 1, ib);
        if (ccw(t, ia, ia + 1, ib) || ccw(t, ia, ib, ia + 1)) connect(t, ia, ib);
        return;
    }
    if (cnt < 4) return;
    int mid = (ia + ib) / 2;
    build(t, ia, mid);
    build(t, mid + 1, ib);

    int li, ri, ul, ur;
    lower_tangent(t, mid, mid + 1, &li, &ri);
    upper_tangent(t, mid, mid + 1, &ul, &ur);    /* both tangents before any edit, :542-543 */
    while (li != ul || ri != ur) {
        int a = 0, b = 0, l1, l2, r1, r2;
        connect(t, li, ri);
        r1 = pred_of(t, ri, li);
        if (left_of(t, r1, li, ri)) {
            r2 = pred_of(t, ri, r1);
            while (in_circle(t, r1, li, ri, r2)) { cut(t, ri, r1); r1 = r2; r2 = pred_of(t, ri, r1); }
        } else a = 1;
        l1 = succ_of(t, li, ri);
        if (right_of(t, l1, ri, li)) {
            l2 = succ_of(t, li, l1);
            while (in_circle(t, li, ri, l1, l2)) { cut(t, li, l1); l1 = l2; l2 = succ_of(t, li, l1); }
        } else b = 1;
        if (a) li = l1;
        else if (b) ri = r1;
        else if (!in_circle(t, li, ri, r1, l1)) ri = r1;     /* ties prefer the right candidate, :588-593 */
        else li = l1;
    }
    connect(t, ul, ur);
}

/* Triangle enumeration in the reference's emission order (convertTrisFreeAdj, :606-635):
 * for each sorted vertex, each CCW-consecutive neighbour pair whose rings are still alive. */
static int enumerate(dt_t *t, int *tri) {
    int ntri = 0;
    for (int i = 0; i < t->n; ++i) {
        int h = t->head[i];
        if (h != NIL && t->nnext[h] != h) {
            int k = h;
            do {
                int n1 = t->nvert[k], n2 = t->nvert[t->nnext[k]];
                if (t->head[n1] != NIL && t->head[n2] != NIL) {
                    if (t->nnext[k] == h && !right_of(t, n1, i, n2)) break;     /* hull gap */
                    tri[3 * ntri] = t->ord[i]; tri[3 * ntri + 1] = t->ord[n1]; tri[3 * ntri + 2] = t->ord[n2];
                    ++ntri;
                }
                k = t->nnext[k];
            } while (k != h);
        }
        t->head[i] = NIL;                     /* "freed" marks the vertex as done */
    }
    return ntri;
}

/* qsort_r comparator restating compareVerts (reference :115-123): x-major, |dx| < 1e-12 is an
 * x-tie decided by y, never returns 0. */
static int cmp_pts(const void *pa, const void *pb, void *ctx) {
    const double *pts = (const double *)ctx;
    const double *a = pts + 2 * (size_t)*(const int *)pa, *b = pts + 2 * (size_t)*(const int *)pb;
    double d = a[0] - b[0];
    if (d < DT_EPS && d > -DT_EPS) d = a[1] - b[1];
    return 1 - 2 * (signbit(d) ? 1 : 0);
}

/* Returns ntriangles (triangles in reference emission order, indices into the input),
 * -1 for fewer than 2 points (reference :416-421), -2 if the input violates the duplicate
 * contract (the reference's own duplicate removal, :440-452, corrupts memory). */
int port_triangulate(const double *points, int npoints, int *tri_out, int *nverts_out) {
    if (npoints < 2) return -1;
    dt_t t;
    t.n = npoints; t.pts = points;
    t.cap = 6 * npoints + 16;
    int *mem = malloc(sizeof(int) * (size_t)(2 * npoints + 3 * t.cap));
    t.ord = mem; t.head = mem + npoints; t.nvert = mem + 2 * npoints; t.nprev = t.nvert + t.cap; t.nnext = t.nprev + t.cap;
    t.nfree = NIL; t.nused = 0;
    for (int i = 0; i < npoints; ++i) { t.ord[i] = i; t.head[i] = NIL; }
    qsort_r(t.ord, (size_t)npoints, sizeof(int), cmp_pts, (void *)points);
    for (int i = 1; i < npoints; ++i) {
        double dx = P(&t, i)[0] - P(&t, i - 1)[0], dy = P(&t, i)[1] - P(&t, i - 1)[1];
        if (dx < DT_EPS && dx > -DT_EPS && dy < DT_EPS && dy > -DT_EPS) { free(mem); return -2; }
    }
    build(&t, 0, npoints - 1);
    int nt = enumerate(&t, tri_out);
    if (nverts_out) *nverts_out = npoints;
    free(mem);
    return nt;
}

/* area_tri (reference include/gta_tri.h:91-99) with GROMACS rvec_sub/cprod/norm, real = double */
static inline double tri_area(const double *a, const double *b, const double *c) {
    double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double cx = ab[1] * ac[2] - ab[2] * ac[1];
    double cy = ab[2] * ac[0] - ab[0] * ac[2];
    double cz = ab[0] * ac[1] - ab[1] * ac[0];
    return sqrt(cx * cx + cy * cy + cz * cz) / 2.0;
}

/* delaunay_surface_area (reference gta_tri.c:332-413): pack XY, triangulate, sum areas in
 * emission order; the 2-D area is the same formula with z := 0. */
static int frame_area(const double *x, int n, int want2d, double *a2d, double *a3d, int *tri, double *xy) {
    for (int i = 0; i < n; ++i) { xy[2 * i] = x[3 * i]; xy[2 * i + 1] = x[3 * i + 1]; }
    int nt = port_triangulate(xy, n, tri, NULL);
    if (nt < 0) return nt;
    double s3 = 0.0, s2 = 0.0;
    for (int i = 0; i < nt; ++i) {
        const double *a = x + 3 * tri[3 * i], *b = x + 3 * tri[3 * i + 1], *c = x + 3 * tri[3 * i + 2];
        s3 += tri_area(a, b, c);
        if (want2d) {
            double a0[3] = {a[0], a[1], 0.0}, b0[3] = {b[0], b[1], 0.0}, c0[3] = {c[0], c[1], 0.0};
            s2 += tri_area(a0, b0, c0);
        }
    }
    *a3d = s3;
    if (want2d) *a2d = s2;
    return nt;
}

/* -corr box-edge points (reference gta_tri.c:112-272). x holds natoms atoms and room for
 * 4 + 2nx + 2ny more; returns the new point count. */
static int add_corr_points(double *x, int natoms, double bx, double by, double espace) {
    int nx = (int)(bx / espace), ny = (int)(by / espace);
    double bot_left = FLT_MAX, top_right = FLT_MIN, top_left = FLT_MAX, bot_right = FLT_MIN;   /* :119-120 */
    int bl = 0, tr = 0, tl = 0, br = 0;
    double *ymin = malloc(sizeof(double) * (size_t)(2 * (nx + 1) + 2 * (ny + 1)));
    double *ymax = ymin + nx + 1, *xmin = ymax + nx + 1, *xmax = xmin + ny + 1;
    int *iymin = calloc((size_t)(2 * (nx + 1) + 2 * (ny + 1)), sizeof(int));
    int *iymax = iymin + nx + 1, *ixmin = iymax + nx + 1, *ixmax = ixmin + ny + 1;
    for (int i = 0; i <= nx; ++i) { ymin[i] = FLT_MAX; ymax[i] = FLT_MIN; }
    for (int i = 0; i <= ny; ++i) { xmin[i] = FLT_MAX; xmax[i] = FLT_MIN; }
    for (int j = 0; j < natoms; ++j) {
        double X = x[3 * j], Y = x[3 * j + 1];
        double d = X * X + Y * Y;
        if (d < bot_left) { bot_left = d; bl = j; }
        if (d > top_right) { top_right = d; tr = j; }
        double dY = by - Y;
        d = X * X + dY * dY;
        if (d < top_left) { top_left = d; tl = j; }
        if (d > bot_right) { bot_right = d; br = j; }
        int xi = (int)((X / bx) * nx);
        if (Y < ymin[xi]) { ymin[xi] = Y; iymin[xi] = j; }
        if (Y > ymax[xi]) { ymax[xi] = Y; iymax[xi] = j; }
        int yi = (int)((Y / by) * ny);
        if (X < xmin[yi]) { xmin[yi] = X; ixmin[yi] = j; }
        if (X > xmax[yi]) { xmax[yi] = X; ixmax[yi] = j; }
    }
    double avg_z = (x[3 * bl + 2] + x[3 * tr + 2] + x[3 * tl + 2] + x[3 * br + 2]) / 4.0;
    int n = natoms;
    double corner[4][2] = {{0, 0}, {bx, 0}, {bx, by}, {0, by}};
    for (int c = 0; c < 4; ++c) { x[3 * n] = corner[c][0]; x[3 * n + 1] = corner[c][1]; x[3 * n + 2] = avg_z; ++n; }
    for (int j = 0; j < nx; ++j) {
        double d1 = x[3 * iymin[j] + 1], d2 = by - x[3 * iymax[j] + 1], d = d1 + d2;
        double z = x[3 * iymin[j] + 2] - (d1 / d) * x[3 * iymin[j] + 2] + x[3 * iymax[j] + 2] - (d2 / d) * x[3 * iymax[j] + 2];
        double e = j * espace + espace / 2;
        x[3 * n] = e; x[3 * n + 1] = 0; x[3 * n + 2] = z; ++n;
        x[3 * n] = e; x[3 * n + 1] = by; x[3 * n + 2] = z; ++n;
    }
    for (int j = 0; j < ny; ++j) {
        double d1 = x[3 * ixmin[j]], d2 = bx - x[3 * ixmax[j]], d = d1 + d2;
        double z = x[3 * ixmin[j] + 2] - (d1 / d) * x[3 * ixmin[j] + 2] + x[3 * ixmax[j] + 2] - (d2 / d) * x[3 * ixmax[j] + 2];
        double e = j * espace + espace / 2;
        x[3 * n] = 0; x[3 * n + 1] = e; x[3 * n + 2] = z; ++n;
        x[3 * n] = bx; x[3 * n + 1] = e; x[3 * n + 2] = z; ++n;
    }
    free(ymin); free(iymin);
    return n;
}

int port_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* delaunay_tessellate (reference gta_tri.c:80-329) over contiguous arrays; same call shape as
 * ref_tessellate in oracle/ref_harness.c. Returns wall seconds. */
double port_tessellate(const double *xyz, const double *box_diag, int nframes, int natoms,
                       double espace, int nthreads, unsigned flags,
                       double *area, double *area2d, double *areabox,
                       double *corr_out, int corr_cap, int *ncorr_out) {
    struct timespec t0, t1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const int corr = (flags & 1u) != 0, want2d = (flags & 2u) != 0;
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nframes; ++f) {
        double bx = box_diag[3 * f], by = box_diag[3 * f + 1];
        int cap = natoms;
        if (corr) cap += 2 * ((int)(bx / espace) + 1) + 2 * ((int)(by / espace) + 1);
        double *x = malloc(sizeof(double) * (size_t)cap * 5);
        double *xy = x + 3 * (size_t)cap;
        int *tri = malloc(sizeof(int) * 6 * (size_t)cap);
        memcpy(x, xyz + (size_t)f * natoms * 3, sizeof(double) * 3 * (size_t)natoms);
        int n = natoms;
        if (areabox) areabox[f] = bx * by;
        if (corr) {
            n = add_corr_points(x, natoms, bx, by, espace);
            if (ncorr_out) ncorr_out[f] = n - natoms;
            if (corr_out) {
                int nc = n - natoms; if (nc > corr_cap) nc = corr_cap;
                memcpy(corr_out + (size_t)f * corr_cap * 3, x + 3 * (size_t)natoms, sizeof(double) * 3 * (size_t)nc);
            }
        }
        double a2 = 0.0, a3 = 0.0;
        frame_area(x, n, want2d, &a2, &a3, tri, xy);
        if (area) area[f] = a3;
        if (want2d && area2d) area2d[f] = a2;
        free(tri); free(x);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
