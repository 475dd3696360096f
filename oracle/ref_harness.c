/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Flat-array entry points around the UNMODIFIED reference objects
 * (src/gta_tri.c, src/delaunay_tri.c, extern/predicates/predicates.c compiled
 * where they lie under /root/reference by oracle/Makefile into oracle/_ref/).
 * The reference API wants `rvec **x` with one heap block per frame
 * (extern/gkut/src/gkut_io.c:37) and reallocs each frame with -corr
 * (src/gta_tri.c:216); this file adapts contiguous test arrays to that shape,
 * supplies the three link stubs the path needs (read_traj, ndx_filter_traj,
 * print_log) and optional predicate call counters (ld --wrap). */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "gta_tri.h"
#include "delaunay_tri.h"

/* ---- link stubs (the real ones need GROMACS trxio; not on the hot path) ---- */
void read_traj(const char *f, rvec ***x, matrix **box, int *nframes, int *natoms, output_env_t *oenv) {
    (void)f; (void)x; (void)box; (void)nframes; (void)natoms; (void)oenv;
    fprintf(stderr, "oracle: read_traj is a stub\n");
    abort();
}
void ndx_filter_traj(const char *f, rvec **pre_x, rvec ***new_x, int nframes, int *natoms) {
    (void)f; (void)pre_x; (void)new_x; (void)nframes; (void)natoms;
    fprintf(stderr, "oracle: ndx_filter_traj is a stub\n");
    abort();
}
static int g_quiet = 1;
void print_log(const char *fmt, ...) {
    if (g_quiet) return;
    va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
}
void ref_set_quiet(int q) { g_quiet = q; }

/* ---- predicate counters (active only in the --wrap build) ---- */
long long g_n_orient = 0, g_n_orient_zero = 0, g_n_incircle = 0, g_n_incircle_zero = 0;
#ifdef ORACLE_WRAP
double __real_orient2d(double *pa, double *pb, double *pc);
double __real_incircle(double *pa, double *pb, double *pc, double *pd);
static FILE *g_dump = NULL;
void ref_set_dump(const char *path) { if (g_dump) fclose(g_dump); g_dump = path ? fopen(path, "wb") : NULL; }
double __wrap_orient2d(double *pa, double *pb, double *pc) {
    double r = __real_orient2d(pa, pb, pc);
    g_n_orient++; if (r == 0.0) g_n_orient_zero++;
    if (g_dump) { double rec[10] = {2.0, pa[0], pa[1], pb[0], pb[1], pc[0], pc[1], 0, 0, r}; fwrite(rec, 8, 10, g_dump); }
    return r;
}
double __wrap_incircle(double *pa, double *pb, double *pc, double *pd) {
    double r = __real_incircle(pa, pb, pc, pd);
    g_n_incircle++; if (r == 0.0) g_n_incircle_zero++;
    if (g_dump) { double rec[10] = {3.0, pa[0], pa[1], pb[0], pb[1], pc[0], pc[1], pd[0], pd[1], r}; fwrite(rec, 8, 10, g_dump); }
    return r;
}
#endif
void ref_counters(long long *out4, int reset) {
    out4[0] = g_n_orient; out4[1] = g_n_orient_zero; out4[2] = g_n_incircle; out4[3] = g_n_incircle_zero;
    if (reset) g_n_orient = g_n_orient_zero = g_n_incircle = g_n_incircle_zero = 0;
}

int ref_sizeof_real(void) { return (int)sizeof(real); }
int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Raw 2-D triangulation through the reference's dtriangulate (delaunay_tri.c:413).
 * tri_out must hold 3*2*(n-1) ints. Returns ntriangles, or -1 if the reference
 * refused the input (it then leaves the struct untouched). */
int ref_triangulate(const double *points, int npoints, int *tri_out, int *nverts_out) {
    struct dTriangulation t;
    t.points = (double *)points; t.npoints = npoints;
    t.triangles = NULL; t.ntriangles = -1; t.nverts = -1;
    dtinit();
    dtriangulate(&t);
    if (t.ntriangles < 0) return -1;
    if (tri_out) memcpy(tri_out, t.triangles, sizeof(int) * 3 * (size_t)t.ntriangles);
    if (nverts_out) *nverts_out = t.nverts;
    free(t.triangles);
    return t.ntriangles;
}

/* Whole-trajectory path through the reference's delaunay_tessellate (gta_tri.c:80).
 * xyz: [nframes][natoms][3] reals, box_diag: [nframes][3] (bx,by,bz).
 * corr_out (nullable): [nframes][corr_cap][3] receives the points the reference
 * appended to each frame (gta_tri.c:216-272); ncorr_out[fr] their count.
 * Returns wall seconds spent inside delaunay_tessellate. */
double ref_tessellate(const real *xyz, const real *box_diag, int nframes, int natoms,
                      real espace, int nthreads, unsigned flags,
                      real *area, real *area2d, real *areabox,
                      real *corr_out, int corr_cap, int *ncorr_out) {
    rvec **x = malloc(sizeof(rvec *) * (size_t)nframes);
    matrix *box = calloc((size_t)nframes, sizeof(matrix));
    for (int f = 0; f < nframes; ++f) {
        x[f] = malloc(sizeof(rvec) * (size_t)natoms);
        memcpy(x[f], xyz + (size_t)f * natoms * 3, sizeof(rvec) * (size_t)natoms);
        box[f][0][0] = box_diag[3 * f]; box[f][1][1] = box_diag[3 * f + 1]; box[f][2][2] = box_diag[3 * f + 2];
    }
    struct tri_area a; a.area = a.area2D = a.area2Dbox = NULL; a.natoms = natoms; a.nframes = nframes;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    delaunay_tessellate(x, box, espace, nthreads, &a, (unsigned char)flags);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (area)    memcpy(area, a.area, sizeof(real) * (size_t)nframes);
    if (areabox) memcpy(areabox, a.area2Dbox, sizeof(real) * (size_t)nframes);
    if (area2d && (flags & GTA_2D)) memcpy(area2d, a.area2D, sizeof(real) * (size_t)nframes);
    if (corr_out && (flags & GTA_CORRECT)) {
        for (int f = 0; f < nframes; ++f) {
            int nx = (int)(box[f][0][0] / espace), ny = (int)(box[f][1][1] / espace);
            int nc = 4 + 2 * nx + 2 * ny;
            if (ncorr_out) ncorr_out[f] = nc;
            if (nc > corr_cap) nc = corr_cap;
            memcpy(corr_out + (size_t)f * corr_cap * 3, x[f] + natoms, sizeof(rvec) * (size_t)nc);
        }
    }
    free_tri_area(&a);
    for (int f = 0; f < nframes; ++f) free(x[f]);
    free(x); free(box);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* One frame through the reference's delaunay_surface_area (gta_tri.c:332). */
void ref_surface_area(const real *xyz, int n, unsigned flags, real *a2d, real *a3d) {
    matrix box; memset(box, 0, sizeof box);
    dtinit();
    delaunay_surface_area((const rvec *)xyz, box, n, (unsigned char)flags, a2d, a3d);
}

/* Direct access to the reference predicates (sign tests for the device port). */
double ref_orient2d(const double *p) { dtinit(); return orient2d((double *)p, (double *)p + 2, (double *)p + 4); }
double ref_incircle(const double *p) { dtinit(); return incircle((double *)p, (double *)p + 2, (double *)p + 4, (double *)p + 6); }
void ref_orient2d_batch(const double *p, int n, double *out) {
    dtinit();
    for (int i = 0; i < n; ++i) out[i] = orient2d((double *)p + 6 * i, (double *)p + 6 * i + 2, (double *)p + 6 * i + 4);
}
void ref_incircle_batch(const double *p, int n, double *out) {
    dtinit();
    for (int i = 0; i < n; ++i) out[i] = incircle((double *)p + 8 * i, (double *)p + 8 * i + 2, (double *)p + 8 * i + 4, (double *)p + 8 * i + 6);
}

/* The reference's own print_areas (src/gta_tri.c:448-473) on caller-supplied arrays: the byte-exact .dat table
 * the product's print_areas is compared with. area2d may be NULL (then the 4-column table). */
void ref_print_areas(const char *fname, const real *area, const real *area2d, const real *areabox, int nframes, int natoms) {
    struct tri_area a;
    a.area = (real *)area; a.area2D = (real *)area2d; a.area2Dbox = (real *)areabox; a.natoms = natoms; a.nframes = nframes;
    print_areas(fname, &a);
}
