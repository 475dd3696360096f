/* oracle/oracle_port.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the reference hot path; see oracle_port.c. */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H

/* sign-exact predicates (restating extern/predicates/predicates.c:1620-1654, :3226-3270) */
double port_orient2d(const double *pa, const double *pb, const double *pc);
double port_incircle(const double *pa, const double *pb, const double *pc, const double *pd);
void port_orient2d_batch(const double *p, int n, double *out);
void port_incircle_batch(const double *p, int n, double *out);

/* statistics of the exact stage (how often the FP64 filter was inconclusive) */
void port_pred_counters(long long *out4, int reset);
void port_pred_count_enable(int on);

int port_triangulate(const double *points, int npoints, int *tri_out, int *nverts_out);
double port_tessellate(const double *xyz, const double *box_diag, int nframes, int natoms,
                       double espace, int nthreads, unsigned flags,
                       double *area, double *area2d, double *areabox,
                       double *corr_out, int corr_cap, int *ncorr_out);
int port_max_threads(void);
#endif
