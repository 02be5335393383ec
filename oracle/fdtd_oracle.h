/*
 * fdtd_oracle.h -- CPU restatement of the reference's FDTD hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may call it,
 * and only as the checker.  The product (libfdtd_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own
 * (SURVEY.md section 4), so the pin is the reference itself: oracle/_ref/libfdtd_ref.so is
 * the unmodified /root/reference/main.c compiled here (oracle/Makefile, oracle/ref_shim.c),
 * and tests/golden/ holds digests generated from it by oracle/make_golden.py.
 * tests/test_oracle.py checks this restatement bit-for-bit against both.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Must be compiled like the reference: gcc -std=c99 (no FP contraction), see oracle/Makefile.
 */
#ifndef FDTD_ORACLE_H
#define FDTD_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors `Parameters` (main.c:57-71): the three cavity sizes and the time limit are
 * single precision in the reference and are promoted to double wherever they are used. */
typedef struct oracle_params {
    float length;          /* x extent, 1st number of params.txt (main.c:226) */
    float width;           /* y extent, 2nd number (main.c:227) */
    float height;          /* z extent, 3rd number (main.c:228) */
    double spatial_step;   /* main.c:229 */
    double time_step;      /* main.c:230 */
    float simulation_time; /* main.c:231 */
    unsigned sampling_rate;/* main.c:232 */
    int mode;              /* main.c:233: 0 validation, 1 computation */
    size_t nx, ny, nz;     /* maxi, maxj, maxk (main.c:237-239) */
} oracle_params;

/* The six staggered arrays in the reference's dense layout (main.c:93-103, 379-407). */
typedef struct oracle_fields {
    double *ex, *ey, *ez, *hx, *hy, *hz;
} oracle_fields;

/* element counts of the six arrays, order ex ey ez hx hy hz (main.c:299,310,321,332,343,354) */
void oracle_field_sizes(const oracle_params *p, size_t out[6]);

/* main.c:216-242.  Returns 0, or -1 if the file cannot be opened. */
int oracle_load_parameters(const char *path, oracle_params *p);
/* same derivation from already-parsed numbers (main.c:237-239) */
void oracle_make_params(float length, float width, float height, double dx, double dt,
                        float simulation_time, unsigned sampling_rate, int mode,
                        oracle_params *p);

/* number of passes of the loop at main.c:765 */
size_t oracle_step_count(const oracle_params *p);

void oracle_set_initial_conditions(const oracle_params *p, double *ey);       /* main.c:416-424 */
void oracle_update_h(const oracle_params *p, const oracle_fields *f);         /* main.c:431-462 */
void oracle_update_e(const oracle_params *p, const oracle_fields *f);         /* main.c:469-500 */
void oracle_set_source(const oracle_params *p, const oracle_fields *f, double t); /* main.c:712-753 */

/* Source patch geometry and per-step amplitudes (main.c:720-751), exposed so the host side of
 * the product can be checked against it.  bounds = {i0, i1, j0, j1}. */
void oracle_source_bounds(const oracle_params *p, long bounds[4]);
double oracle_source_zte(const oracle_params *p);

/* `steps` passes of the loop body main.c:770-779, time accumulated as at main.c:765.
 * t_io: in = starting time counter, out = time counter after the last pass. */
void oracle_run(const oracle_params *p, const oracle_fields *f, size_t steps, double *t_io);

/* main.c:511-540 with the argument triples used at main.c:563-578.
 * var: 0 ex, 1 ey, 2 ez, 3 hx, 4 hy, 5 hz.  out has nx*ny*nz doubles. */
void oracle_aggregate(const oracle_params *p, const oracle_fields *f, int var, double *out);

/* main.c:670-710: analytic TE101 fields minus the computed ones. vey/vhx/vhz sized like ey/hx/hz. */
void oracle_validation_fields(const oracle_params *p, const oracle_fields *f,
                              double *vey, double *vhx, double *vhz, double t);

/* main.c:602-668 as coded (Ez indexed with Hz's strides, main.c:627): out = {electric, magnetic} */
void oracle_energy(const oracle_params *p, const oracle_fields *f, double out[2]);

#ifdef __cplusplus
}
#endif
#endif
