/* oracle/shim/smalloc.h -- TEST INFRASTRUCTURE ONLY.
 * GROMACS smalloc.h semantics: snew = zero-initialised allocation,
 * srenew = realloc, sfree = free. */
#ifndef ORACLE_SHIM_SMALLOC_H
#define ORACLE_SHIM_SMALLOC_H
#include <stdlib.h>
#define snew(ptr, n)   (ptr) = calloc((size_t)((n) > 0 ? (n) : 1), sizeof(*(ptr)))
#define srenew(ptr, n) (ptr) = realloc((ptr), (size_t)(n) * sizeof(*(ptr)))
#define sfree(ptr)     free(ptr)
#endif
