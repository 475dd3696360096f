/* oracle/shim/vec.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the GROMACS 5.0 legacy header "vec.h" so that the
 * reference's src/gta_tri.c compiles unmodified from /root/reference
 * (it includes it through include/gta_tri.h:15).  GROMACS is an external,
 * un-vendored dependency of the reference (README.md:31, makefile:20-28);
 * these are restatements of its documented semantics, not copies.
 * real = double (GMX_DOUBLE build) is the parity contract of this repo. */
#ifndef ORACLE_SHIM_VEC_H
#define ORACLE_SHIM_VEC_H
#include <math.h>
#include <stdio.h>

#ifdef ORACLE_REAL_FLOAT
typedef float real;
#else
typedef double real;
#endif
#define DIM 3
#define XX 0
#define YY 1
#define ZZ 2
typedef real rvec[DIM];
typedef real matrix[DIM][DIM];
typedef int gmx_bool;
typedef int atom_id;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

static inline void copy_rvec(const rvec a, rvec b) {
    b[XX] = a[XX]; b[YY] = a[YY]; b[ZZ] = a[ZZ];
}
static inline void rvec_sub(const rvec a, const rvec b, rvec c) {
    real x = a[XX] - b[XX], y = a[YY] - b[YY], z = a[ZZ] - b[ZZ];
    c[XX] = x; c[YY] = y; c[ZZ] = z;
}
static inline void cprod(const rvec a, const rvec b, rvec c) {
    c[XX] = a[YY] * b[ZZ] - a[ZZ] * b[YY];
    c[YY] = a[ZZ] * b[XX] - a[XX] * b[ZZ];
    c[ZZ] = a[XX] * b[YY] - a[YY] * b[XX];
}
static inline real iprod(const rvec a, const rvec b) {
    return a[XX] * b[XX] + a[YY] * b[YY] + a[ZZ] * b[ZZ];
}
static inline real norm(const rvec a) {
    return (real)sqrt(iprod(a, a));
}
#endif
