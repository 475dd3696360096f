/* oracle/shim/pargs.h -- TEST INFRASTRUCTURE ONLY. Opaque GROMACS type used
 * only in prototypes on the path (gta_tri.h:40-46, gkut_io.h:24). */
#ifndef ORACLE_SHIM_PARGS_H
#define ORACLE_SHIM_PARGS_H
typedef struct output_env *output_env_t;
#endif
