/* delaunay_tri.h -- kept entry points of the reference's triangulator, served by libgtessla_b200.so.
 *
 * Same names, types and ownership rules as the reference's include/delaunay_tri.h:27-46, so code
 * written against g_tessla's triangulator links against this library unchanged:
 *   - the caller owns `points` (npoints interleaved x,y pairs, read-only to the library);
 *   - dtriangulate() malloc()s `triangles` (3 ints per triangle, indices into `points`, counter-clockwise,
 *     in the reference's emission order delaunay_tri.c:613-629) and the caller free()s it (gta_tri.c:412);
 *   - with fewer than 2 points it prints "TRIANGULATION ERROR ..." on stderr and leaves
 *     triangles / ntriangles / nverts untouched (delaunay_tri.c:416-421).
 * What differs: the work runs on a CUDA device (one frame of a batched sm_100a kernel for up to
 * gta_b200_max_points() points, the single-large-frame path above that); there is no CPU fallback.
 * Inputs the reference mishandles are refused instead of reproduced: two points closer than 1e-12 in
 * x and y (the reference corrupts its vertex array, delaunay_tri.c:446) print a TRIANGULATION ERROR
 * and leave the struct untouched; dt_last_status() returns the GTA_B200_ST_* bits of the last call
 * on this thread.
 */
#ifndef DELAUNAY_TRI_H
#define DELAUNAY_TRI_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef REAL
#define REAL double            /* reference extern/predicates/predicates.h:36 */
#endif
typedef REAL dtreal;

struct dTriangulation {
    dtreal *points;    /* in:  x0,y0,x1,y1,... (2 reals per point) */
    int npoints;       /* in */
    int *triangles;    /* out: malloc'd by dtriangulate, 3 point indices per triangle */
    int ntriangles;    /* out */
    int nverts;        /* out: number of distinct input points (== npoints on accepted input) */
};

/* Once before any dtriangulate call (reference: fills the predicate error bounds; here: selects the
 * CUDA device and fails loudly on stderr if there is none). Safe to call repeatedly. */
void dtinit(void);

/* Delaunay triangulation of tri->points. Thread-safe (each host thread drives its own stream). */
void dtriangulate(struct dTriangulation *tri);

/* GTA_B200_ST_* bits (include/gta_b200.h) of the last dtriangulate / delaunay_surface_area call made
 * by the calling thread; 0 = ok. The reference has no error channel besides stderr. */
int dt_last_status(void);

#ifdef __cplusplus
}
#endif
#endif /* DELAUNAY_TRI_H */
