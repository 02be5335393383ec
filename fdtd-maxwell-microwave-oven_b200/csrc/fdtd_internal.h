/* Internal declarations shared by the host C part and the CUDA part of libfdtd_b200.so. */
#ifndef FDTD_INTERNAL_H
#define FDTD_INTERNAL_H

#include "fdtd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* printf-style; stores the message returned by fdtd_last_error() (per thread). */
void fdtd_set_error(const char *fmt, ...);

/* Physical constants exactly as the reference spells them (main.c:22-25), including the
 * truncated epsilon_0: the update factors must come out bit-identical. */
#define FDTD_MU 1.25663706143591729538505735331180115367886775975E-6
#define FDTD_EPSILON 8.854E-12
#define FDTD_PI 3.14159265358979323846264338327950288419716939937510582097494
#define FDTD_CELERITY 299792458.0

/* time_step / (MU * spatial_step), main.c:441, and time_step / (EPSILON * spatial_step),
 * main.c:479 -- evaluated on the host in double, passed to the kernels by value. */
double fdtd_factor_h(const fdtd_params *p);
double fdtd_factor_e(const fdtd_params *p);

/* set_initial_conditions(), main.c:416-424, for the node planes [k_first, k_first + nplanes) only */
int fdtd_initial_conditions_planes(const fdtd_params *p, size_t k_first, size_t nplanes, double *Ey);

#ifdef __cplusplus
}
#endif
#endif
