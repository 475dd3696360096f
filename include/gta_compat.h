/* gta_compat.h -- the few GROMACS names the kept g_tessla headers need, without GROMACS.
 *
 * The reference builds against GROMACS 4.5-5.0 (makefile:7-33) for `real`, `rvec`, `matrix`, the
 * XX/YY/ZZ indices, `gmx_bool`, `output_env_t` and the inline vector helpers used by area_tri
 * (include/gta_tri.h:91-99). GROMACS is not vendored by the reference and is not available here, so
 * this header restates those names with their documented GROMACS meaning. `real` is double: the FP64
 * contract of this library (a GROMACS build with -DGMX_DOUBLE). When real GROMACS headers are wanted
 * instead, define GTA_USE_GROMACS before including gta_tri.h (and build GROMACS in double precision).
 */
#ifndef GTA_COMPAT_H
#define GTA_COMPAT_H

#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef GTA_USE_GROMACS
typedef double real;
typedef real rvec[3];
typedef real matrix[3][3];
typedef int gmx_bool;
typedef struct gta_output_env *output_env_t;   /* opaque; the library never dereferences it */
#ifndef XX
#define XX 0
#define YY 1
#define ZZ 2
#endif
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

static inline void copy_rvec(const rvec src, rvec dst) { dst[XX] = src[XX]; dst[YY] = src[YY]; dst[ZZ] = src[ZZ]; }
static inline void rvec_sub(const rvec a, const rvec b, rvec d) { d[XX] = a[XX] - b[XX]; d[YY] = a[YY] - b[YY]; d[ZZ] = a[ZZ] - b[ZZ]; }
static inline void cprod(const rvec a, const rvec b, rvec c) {
    c[XX] = a[YY] * b[ZZ] - a[ZZ] * b[YY];
    c[YY] = a[ZZ] * b[XX] - a[XX] * b[ZZ];
    c[ZZ] = a[XX] * b[YY] - a[YY] * b[XX];
}
static inline real norm(const rvec a) { return sqrt(a[XX] * a[XX] + a[YY] * a[YY] + a[ZZ] * a[ZZ]); }
#endif /* GTA_USE_GROMACS */

#ifdef __cplusplus
}
#endif
#endif /* GTA_COMPAT_H */
