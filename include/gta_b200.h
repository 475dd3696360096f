/* gta_b200.h -- C-ABI of the B200-native g_tessla hot path (libgtessla_b200.so).
 *
 * Drop-in boundary for the reference's per-frame path
 *     delaunay_tessellate -> delaunay_surface_area -> dtriangulate -> orient2d / incircle
 *     (reference src/gta_tri.c:80, :332; src/delaunay_tri.c:413;
 *      extern/predicates/predicates.c:1620, :3226).
 * The kept reference entry points (same names and signatures) are declared in
 * include/delaunay_tri.h and include/gta_tri.h; this header adds the batch-native calls they
 * are built on, plus per-frame status codes (every reference function is `void` and reports
 * errors only on stderr, reference src/delaunay_tri.c:416-421).
 *
 * Plain pointers and sizes only. `real` is double (the FP64 contract).
 * There is NO CPU implementation behind these calls: without a CUDA device they fail with
 * GTA_B200_E_NO_DEVICE. */
#ifndef GTA_B200_H
#define GTA_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flags: identical to the reference's enum (include/gta_tri.h:24-28) */
#define GTA_B200_CORRECT 1u /* -corr : add box-edge correction points (src/gta_tri.c:112-272) */
#define GTA_B200_2D      2u /* -2d   : also sum the XY-projected area (src/gta_tri.c:368-386)  */

/* per-frame status bits (0 = ok) */
#define GTA_B200_ST_OK              0
#define GTA_B200_ST_TOO_FEW_POINTS  1   /* < 2 points: reference prints "TRIANGULATION ERROR" (delaunay_tri.c:416) */
#define GTA_B200_ST_DUPLICATE       2   /* two points within 1e-12 in x and y: reference corrupts memory (delaunay_tri.c:446) */
#define GTA_B200_ST_CAPACITY        4   /* frame needs more points than max_points */
#define GTA_B200_ST_PRECISION_RANGE 8   /* operands of an exact predicate span > 251 bits */
#define GTA_B200_ST_OUT_OF_BOX      16  /* -corr with an atom outside [0,bx]x[0,by]: reference indexes out of bounds (gta_tri.c:178-201) */
#define GTA_B200_ST_POOL            32  /* internal edge pool exhausted (never on planar input) */
#define GTA_B200_ST_NONFINITE       64  /* NaN / Inf coordinate or box */
#define GTA_B200_ST_NEAR_TIE_X      128 /* a run of near-equal x (gaps < 1e-12) wider than 1e-12: reference comparator not transitive (delaunay_tri.c:115-123) */
#define GTA_B200_ST_INTERNAL        256 /* ring inconsistency (bug guard) */

/* return codes */
#define GTA_B200_OK              0
#define GTA_B200_E_NO_DEVICE    (-1)
#define GTA_B200_E_CUDA         (-2)
#define GTA_B200_E_ARG          (-3)
#define GTA_B200_E_TOO_LARGE    (-4)   /* max_points does not fit the shared-memory kernel */

/* Engines behind the batched calls (every one returns the reference's triangulation; areas agree to the rounding of the
 * summation order, |rel| ~ 1e-16):
 *   star path (default when no triangle list is asked for): one fused launch, a CTA per frame, a thread per point builds
 *     that point's Delaunay star from a uniform grid with exact predicates and certifies it (csrc/star_core.cuh). Frames
 *     whose Delaunay triangulation is not unique (cocircular ties), duplicates and refused frames are handed, inside the
 *     same call, to the divide-and-conquer engine, which keeps the reference's tie-breaking (delaunay_tri.c:582-593).
 *   divide and conquer (always when d_tri / tri is given: its list is in the reference's emission order,
 *     delaunay_tri.c:613-629): the two kernels described below.
 * gta_b200_set_path: 0 = automatic (above), 1 = divide and conquer only, 2 = star path also for triangle lists (the list
 * then holds the same triangles in another order). Environment GTA_B200_PATH sets the default. Returns the previous mode. */
int gta_b200_set_path(int mode);
int gta_b200_get_path(void);

/* Two kernels serve the divide-and-conquer engine: launches of at least `min_frames` frames run the throughput path (pre-pass,
 * a persistent kernel whose phase-sorted scheduler steps (frame, subtree) walkers, area kernel: highest throughput), smaller ones the fused one-CTA-per-frame kernel (lowest latency). Results are identical
 * triangulations; areas agree to rounding of the summation tree. Environment GTA_B200_LPF overrides the default;
 * 0 = never, 1 = always, negative = restore the default. Returns the previous value. */
#define GTA_B200_DEFAULT_BATCH_THRESHOLD 2048
int gta_b200_set_batch_threshold(int min_frames);

/* Largest max_points the batched shared-memory kernel accepts on this device. */
int gta_b200_max_points(void);
int gta_b200_device_count(void);
const char *gta_b200_last_error(void);

/* Points a frame can reach with -corr: natoms + 4 + 2*floor(bx/espace) + 2*floor(by/espace)
 * maximised over the given boxes (host arithmetic; mirrors gta_tri.c:112-113,216). */
int gta_b200_max_points_for(const double *box, int box_stride, int nframes, int natoms,
                            double espace, unsigned flags);

/* Threads per frame, dynamic shared memory and resident frames per SM the kernel will use for frames of
 * natoms caller points and at most max_points points in total (for reports). */
int gta_b200_launch_shape(int max_points, int natoms, int *threads_per_frame, size_t *smem_bytes, int *frames_per_sm);

/* ---- device-resident batch: every pointer is DEVICE memory, work is enqueued on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream) and NOT synchronised.
 *   d_xyz      [nframes][natoms][in_dim] doubles; in_dim = 3 (x,y,z atoms, replaces the reference's
 *              rvec **x, gta_tri.c:80) or 2 (raw 2-D point sets, replaces dTriangulation.points,
 *              delaunay_tri.h:28; z is taken as 0)
 *   d_box      [nframes][box_stride] doubles, bx at [0], by at [4] if box_stride == 9 (the
 *              reference's `matrix`), else at [1]; may be NULL when GTA_B200_CORRECT is off
 *   max_points capacity per frame (>= natoms; with -corr see gta_b200_max_points_for)
 *   d_area, d_area2d (nullable), d_areabox (nullable)   [nframes] doubles
 *   d_status   [nframes] ints (nullable)
 *   d_tri      (nullable) [nframes][3*tri_cap] ints, triangles in the reference's emission order
 *              (delaunay_tri.c:613-629), indices into the frame's points incl. appended -corr points
 *   d_ntri     (nullable) [nframes] triangle counts
 *   d_corr     (nullable) [nframes][corr_cap][3] appended -corr points, d_ncorr (nullable) their count
 *   d_work     one unsigned int of device scratch (frame counter), zeroed by the call
 */
int gta_b200_tessellate_device(const double *d_xyz, int in_dim, const double *d_box, int box_stride,
                               int nframes, int natoms, int max_points, double espace, unsigned flags,
                               double *d_area, double *d_area2d, double *d_areabox, int *d_status,
                               int *d_tri, int tri_cap, int *d_ntri,
                               double *d_corr, int corr_cap, int *d_ncorr,
                               unsigned int *d_work, void *stream);

/* ---- host batch: every pointer is HOST memory (pageable or pinned). Frames are staged through
 * pinned buffers in chunks, copied H2D on `nstreams` streams, tessellated, results copied back.
 * Replaces the OpenMP frame loop of delaunay_tessellate (gta_tri.c:106, :301). Synchronous.
 * device < 0 selects the current device. Same array shapes as above. */
int gta_b200_tessellate_batch(const double *xyz, int in_dim, const double *box, int box_stride,
                              int nframes, int natoms, double espace, unsigned flags,
                              int device, int nstreams,
                              double *area, double *area2d, double *areabox, int *status,
                              int *tri, int tri_cap, int *ntri,
                              double *corr, int corr_cap, int *ncorr);

/* Same, sharding contiguous frame ranges over ngpus devices (0..ngpus-1), one host thread per
 * device, no collective: each device writes its slice of the result arrays. */
int gta_b200_tessellate_multi(const double *xyz, int in_dim, const double *box, int box_stride,
                              int nframes, int natoms, double espace, unsigned flags,
                              int ngpus, int nstreams,
                              double *area, double *area2d, double *areabox, int *status);

/* Frame-pointer variant for the reference's own data layout: x[fr] points to natoms*3 reals
 * (one heap block per frame, extern/gkut/src/gkut_io.c:37), box[fr] is a 3x3 matrix.
 * nthreads = host staging threads (<= 0: all cores), the meaning -nthreads keeps here. */
int gta_b200_tessellate_frames(const double *const *x, const double *box9, int nframes, int natoms,
                               double espace, unsigned flags, int ngpus, int nthreads,
                               double *area, double *area2d, double *areabox, int *status);

/* ---- one large frame (reference dtriangulate on a point set beyond gta_b200_max_points(), e.g. the
 * 10^6-point coarse-grained patch): HOST pointers, raw 2-D points [npoints][2], triangles [tri_cap][3] in
 * the reference's emission order, *ntri their count, *status the GTA_B200_ST_* bits. Runs on one GPU
 * (a single triangulation does not shard). */
int gta_b200_triangulate_large(const double *points, int npoints, int device,
                               int *tri, int tri_cap, int *ntri, int *status);
/* Same path for one large frame of atoms xyz [natoms][3] WITHOUT correction points: also sums the 3-D triangle areas
 * (*area3) and, with GTA_B200_2D, the projected ones (*area2). delaunay_surface_area / delaunay_tessellate use it for
 * frames beyond gta_b200_max_points(). GTA_B200_CORRECT is refused here (GTA_B200_E_TOO_LARGE): gta_b200_tessellate_large. */
int gta_b200_surface_area_large(const double *xyz, int natoms, unsigned flags, int device,
                                double *area3, double *area2, int *ntri, int *status);
/* One large frame the way delaunay_tessellate treats a frame (reference src/gta_tri.c:106-296): with GTA_B200_CORRECT the
 * box-edge points are generated on the device, appended (*ncorr of them) and the natoms + *ncorr points go through the
 * large-frame path; box9 = the frame's 3 x 3 box (box9[0], box9[4] are used), *areabox = their product. A frame the
 * batched calls would refuse (atoms outside the box with -corr, non-finite input) gets the same status bits and NaN areas. */
int gta_b200_tessellate_large(const double *xyz, int natoms, const double *box9, double espace, unsigned flags, int device,
                              double *area3, double *area2, double *areabox, int *ntri, int *ncorr, int *status);
/* device milliseconds (sort + all levels + enumeration, without the host copies) of the last call on this thread */
double gta_b200_large_last_ms(void);
/* engine of the last large-frame call on this thread: 1 = star engine (csrc/gta_large_star.cuh: areas by default,
 * triangle lists -- in cell order -- under gta_b200_set_path(2)), 0 = level-by-level divide and conquer */
int gta_b200_large_last_engine(void);

/* Milliseconds the last gta_b200_tessellate_batch call on this thread spent in its kernels
 * (CUDA events on the launching streams), and the number of kernel launches the last batch / device call made. */
double gta_b200_last_kernel_ms(void);
int gta_b200_last_launches(void);
/* Per-kernel device time of the throughput path (pre-pass, walker kernel, area kernel) on `device` (< 0: the current
 * one): _reset forgets the launches seen so far; _ms waits for the launches made since and adds up each kernel's
 * duration, measured with CUDA events on the launching streams (at most 64 launches are kept). Any pointer may be NULL. */
int gta_b200_stage_reset(int device);
int gta_b200_stage_ms(int device, double *prepass_ms, double *walk_ms, double *area_ms, int *launches);
/* The same with the star kernel in front: out4 = {star kernel, pre-pass, walker kernel, area kernel} in ms. */
int gta_b200_stage_ms4(int device, double *out4, int *launches);
/* Frames the most recent sampled star launch on `device` handed to the divide-and-conquer engine (waits for it); < 0: none yet. */
long long gta_b200_last_fallback_frames(int device);
/* Running totals of the star engine on `device` since the last reset (synchronises the device): frames certified by their
 * stars and frames handed to the divide-and-conquer engine. Either pointer may be NULL; reset != 0 zeroes the totals. */
int gta_b200_star_counters(int device, unsigned long long *certified, unsigned long long *fallback, int reset);

/* FP64 peak of the device in TFLOP/s, measured with independent DFMA chains (2 flops each): the denominator of the
 * FP64 roofline bench.py reports (MEASURED_PEAKS.json holds no FP64 figure). <= 0 on error. device < 0: current. */
double gta_b200_fp64_peak_tflops(int device);

/* Predicate self-test hooks (device evaluation of the exact-sign predicates on arrays):
 * pts [n][6] -> sign [n] for orient2d, pts [n][8] for incircle; host pointers. */
int gta_b200_orient2d_signs(const double *pts, int n, int *sign_out, int device);
int gta_b200_incircle_signs(const double *pts, int n, int *sign_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* GTA_B200_H */
