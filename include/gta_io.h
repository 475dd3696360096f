/* gta_io.h -- trajectory / index input behind the kept tessellate_area(), without GROMACS.
 *
 * Replaces the reference's gkut_io (extern/gkut/include/gkut_io.h:24-46, src/gkut_io.c:20-97), which
 * wraps GROMACS' read_first_x / read_next_x / rd_index. Same names and argument meaning; the formats
 * are decoded here: .gro (one or many frames), .pdb (CRYST1 + MODEL/ENDMDL, Angstrom -> nm) and .trr
 * (XDR, single or double precision) and .xtc (XDR + compressed coordinates; restated from the published
 * format, see csrc/gta_host.cu).
 */
#ifndef GTA_IO_H
#define GTA_IO_H
#include "gta_compat.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Whole trajectory into memory: (*x)[fr] is one malloc'd block of natoms rvecs, (*box)[fr] a 3x3 matrix
 * (reference gkut_io.c:20-46). Returns through the out-parameters; on error prints a message, sets
 * *nframes = 0 and returns. Caller frees every (*x)[fr], *x and *box with free(). */
void read_traj(const char *traj_fname, rvec ***x, matrix **box, int *nframes, int *natoms, output_env_t *oenv);

/* Keep the atoms listed in the FIRST group of the index file (1-based atom numbers, reference
 * gkut_io.c:80-97): (*new_x)[fr] holds the selected atoms in index order, *natoms becomes the group size. */
void ndx_filter_traj(const char *ndx_fname, rvec **pre_x, rvec ***new_x, int nframes, int *natoms);

#ifdef __cplusplus
}
#endif
#endif
