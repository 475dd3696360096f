/* g_tessla -- command-line driver with the reference's option surface (reference src/g_tessla.c:27-139),
 * without GROMACS: -f/-n/-o file options, -nthreads, -[no]corr, -espace, -[no]2d, -[no]print.
 * The experimental -dense weighted-grid mode (src/gta_grid.c, "not supported" per the reference README)
 * is outside this build; -dense / -width / -lin are recognised and refused with a message.
 * Everything numerical happens in libgtessla_b200.so (CUDA, sm_100a); this file only parses and prints. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gta_tri.h"

static void usage(const char *prog) {
    printf("%s calculates 3-d surface area using Delaunay tessellation (B200-native build).\n"
           "  -f  <traj>    trajectory: .gro .pdb .trr .xtc   (default traj.xtc)\n"
           "  -n  <ndx>     index file; the first group selects the atoms (optional)\n"
           "  -o  <dat>     output table                      (default tessellated_areas.dat)\n"
           "  -nthreads N   host staging streams (default: library choice)\n"
           "  -[no]corr     correct the area for periodic bounds with box-edge points\n"
           "  -espace X     spacing of the edge correction points in nm (default 0.8)\n"
           "  -[no]2d       also report the area of the XY projection\n"
           "  -[no]print    write triangles<k>.node/.ele per frame (showme format; one frame at a time)\n"
           "  environment:  GTA_B200_NGPUS = number of GPUs to shard frames over (default all)\n", prog);
}

/* GROMACS-style booleans: -name sets, -noname clears */
static int bool_opt(const char *arg, const char *name, int *val) {
    if (arg[0] != '-') return 0;
    if (!strcmp(arg + 1, name)) { *val = 1; return 1; }
    if (!strncmp(arg + 1, "no", 2) && !strcmp(arg + 3, name)) { *val = 0; return 1; }
    return 0;
}

int main(int argc, char *argv[]) {
    const char *traj = "traj.xtc", *ndx = NULL, *out = "tessellated_areas.dat";
    int nthreads = -1, dense = 0, corr = 0, a2D = 0, print = 0, linear = 0;
    real espace = 0.8, cell_width = 0.1;

    init_log("gta.log", argc, argv);
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        const int more = i + 1 < argc;
        if (!strcmp(a, "-h") || !strcmp(a, "-help") || !strcmp(a, "--help")) { usage(argv[0]); close_log(); return 0; }
        else if (!strcmp(a, "-f") && more) traj = argv[++i];
        else if (!strcmp(a, "-n") && more) ndx = argv[++i];
        else if (!strcmp(a, "-o") && more) out = argv[++i];
        else if (!strcmp(a, "-nthreads") && more) nthreads = atoi(argv[++i]);
        else if (!strcmp(a, "-espace") && more) espace = atof(argv[++i]);
        else if (!strcmp(a, "-width") && more) cell_width = atof(argv[++i]);
        else if (bool_opt(a, "corr", &corr) || bool_opt(a, "2d", &a2D) || bool_opt(a, "print", &print) ||
                 bool_opt(a, "dense", &dense) || bool_opt(a, "lin", &linear)) { }
        else { fprintf(stderr, "g_tessla: unknown or incomplete option '%s' (try -h)\n", a); close_log(); return 2; }
    }
    (void)cell_width; (void)linear;
    if (dense) {
        print_log("The -dense weighted-grid mode is experimental in the reference and is not part of this build.\n");
        close_log();
        return 2;
    }
    if (print) nthreads = 1;                                   /* reference g_tessla.c:116 */

    struct tri_area areas;
    unsigned long flags = ((int)corr * GTA_CORRECT) | ((int)a2D * GTA_2D) | ((int)print * GTA_PRINT);
    tessellate_area(traj, ndx, NULL, espace, nthreads, &areas, (unsigned char)flags);
    int rc = 0;
    if (gta_last_call_rc() != 0) { print_log("g_tessla: the tessellation failed (error %d); no table written.\n", gta_last_call_rc()); rc = 1; }
    else if (areas.nframes > 0 && areas.area) print_areas(out, &areas);
    else { print_log("No frames were tessellated.\n"); rc = 1; }
    free_tri_area(&areas);
    close_log();
    return rc;
}
