/* gta_tri.h -- kept entry points of the reference's per-frame area pipeline (reference
 * include/gta_tri.h:24-99), served by libgtessla_b200.so.
 *
 * Call-compatible with the reference: same flag values, same struct tri_area, same function names,
 * argument order and ownership. The frame loop the reference runs under OpenMP (src/gta_tri.c:106,
 * :301) is one batched CUDA launch per chunk of frames here; `nthreads` keeps its place in the
 * signature and now sizes the host staging (number of streams); all visible GPUs are used unless the
 * environment variable GTA_B200_NGPUS says otherwise.
 */
#ifndef GTA_TRI_H
#define GTA_TRI_H

#ifdef GTA_USE_GROMACS
#include "vec.h"
#ifdef GRO_V5
#include "pargs.h"
#else
#include "statutil.h"
#endif
#else
#include "gta_compat.h"
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* flags (bit values of the reference, include/gta_tri.h:24-28) */
enum {
    GTA_CORRECT = 1,   /* -corr : add box-edge points so the tessellation covers the periodic cell */
    GTA_2D = 2,        /* -2d   : also report the area of the XY projection */
    GTA_PRINT = 4      /* -print: write triangles<k>.node / .ele per frame (showme format) */
};

/* Totals per frame; divide by natoms for area per particle. natoms and nframes are INPUTS of
 * delaunay_tessellate (the reference sets them in tessellate_area, src/gta_tri.c:55). */
struct tri_area {
    real *area;        /* [nframes] 3-D tessellated area (with -corr: of the corrected point set) */
    real *area2D;      /* [nframes] area of the XY projection; NULL unless GTA_2D */
    real *area2Dbox;   /* [nframes] box[0][0] * box[1][1] */
    int natoms, nframes;
};

/* Read a trajectory (and optionally keep only the atoms of the first group of an index file), then
 * delaunay_tessellate it. Formats read without GROMACS: .gro, .pdb (MODEL/ENDMDL, CRYST1),
 * .trr (single or double precision) and .xtc. oenv is unused and may be NULL. */
void tessellate_area(const char *traj_fname, const char *ndx_fname, output_env_t *oenv,
                     real espace, int nthreads, struct tri_area *areas, unsigned char flags);

/* x[fr] -> natoms rvecs, box[fr] = 3x3 matrix (only [0][0] and [1][1] are read). Allocates
 * areas->area / area2Dbox (/ area2D with GTA_2D) with calloc; release with free_tri_area. Unlike the
 * reference (src/gta_tri.c:216) the caller's frames are never reallocated or modified. */
void delaunay_tessellate(rvec **x, matrix *box, real espace, int nthreads,
                         struct tri_area *areas, unsigned char flags);

/* One frame without correction points: a3D and/or a2D (either may be NULL) receive the summed
 * triangle areas, as in the reference (src/gta_tri.c:332-413). */
void delaunay_surface_area(const rvec *x, matrix box, int natoms, unsigned char flags, real *a2D, real *a3D);

/* ASCII table, byte-compatible with the reference (src/gta_tri.c:448-473). */
void print_areas(const char *fname, const struct tri_area *areas);

void free_tri_area(struct tri_area *areas);

/* Per-frame GTA_B200_ST_* bits of the last delaunay_tessellate call on this thread ([nframes], owned by
 * the library, valid until the next call); NULL if none. Frames with a nonzero status have area = NaN. */
const int *gta_last_frame_status(void);
/* GTA_B200_* return code (include/gta_b200.h; 0 = ok) behind the last tessellate_area / delaunay_tessellate /
 * delaunay_surface_area / dtriangulate call on this thread: the kept entry points are void, as in the reference. After a
 * failure every area of that call is NaN (never 0) and the cause is in gta_b200_last_error(). */
int gta_last_call_rc(void);

/* stdout + gta.log tee of the reference's gkut_log (extern/gkut/src/gkut_log.c:21-50) */
void init_log(const char *logfile, int argc, char *argv[]);
void close_log(void);
void print_log(const char *fmt, ...);

/* half the norm of (b-a) x (c-a) (reference include/gta_tri.h:91-99) */
static inline real area_tri(const rvec a, const rvec b, const rvec c) {
    rvec ab, ac, cp;
    rvec_sub(b, a, ab);
    rvec_sub(c, a, ac);
    cprod(ab, ac, cp);
    return norm(cp) / 2.0;
}

#ifdef __cplusplus
}
#endif
#endif /* GTA_TRI_H */
