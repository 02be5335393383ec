/*
 * fdtd_host.c -- host-side (C99) half of libfdtd_b200.so: everything of the reference's path
 * that must be evaluated by the host C library to stay bit-identical with the reference
 * (parameter parsing through scanf, the float->double promotions that decide the grid, glibc
 * sin/sqrt/pow for the source amplitudes and the TE101 initial condition).
 *
 * Build with -std=c99 -ffp-contract=off like the reference (Makefile:12,18 of the reference).
 */
#include "fdtd_internal.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_error[512];

void fdtd_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

const char *fdtd_last_error(void) { return g_error; }

int fdtd_abi_version(void) { return FDTD_B200_ABI_VERSION; }

double fdtd_factor_h(const fdtd_params *p) { return p->time_step / (FDTD_MU * p->spatial_step); }
double fdtd_factor_e(const fdtd_params *p) { return p->time_step / (FDTD_EPSILON * p->spatial_step); }

/* main.c:237-239: (size_t)(float_size / double_step) with the float promoted first. */
static void derive_grid(fdtd_params *p)
{
    p->maxi = (size_t)(p->length / p->spatial_step);
    p->maxj = (size_t)(p->width / p->spatial_step);
    p->maxk = (size_t)(p->height / p->spatial_step);
}

int fdtd_make_params(float length, float width, float height, double spatial_step,
                     double time_step, float simulation_time, unsigned sampling_rate, int mode,
                     fdtd_params *out)
{
    if (!out) { fdtd_set_error("fdtd_make_params: out is NULL"); return FDTD_E_ARG; }
    memset(out, 0, sizeof *out);
    out->length = length;
    out->width = width;
    out->height = height;
    out->spatial_step = spatial_step;
    out->time_step = time_step;
    out->simulation_time = simulation_time;
    out->sampling_rate = sampling_rate;
    out->mode = mode;
    derive_grid(out);
    return FDTD_OK;
}

/* main.c:216-242.  The reference does not look at scanf's return value; a short or malformed
 * file therefore yields whatever was parsed so far -- here the rest stays zero. */
int fdtd_load_parameters(const char *path, fdtd_params *out)
{
    FILE *fp;
    unsigned mode = 0;
    int got = 0;
    if (!path || !out) { fdtd_set_error("fdtd_load_parameters: NULL argument"); return FDTD_E_ARG; }
    fp = fopen(path, "r");
    if (!fp) {
        fdtd_set_error("Unable to open parameters file!"); /* message of main.c:223 */
        return FDTD_E_IO;
    }
    memset(out, 0, sizeof *out);
    got += fscanf(fp, "%f", &out->length) == 1;
    got += fscanf(fp, "%f", &out->width) == 1;
    got += fscanf(fp, "%f", &out->height) == 1;
    got += fscanf(fp, "%lf", &out->spatial_step) == 1;
    got += fscanf(fp, "%lf", &out->time_step) == 1;
    got += fscanf(fp, "%f", &out->simulation_time) == 1;
    got += fscanf(fp, "%u", &out->sampling_rate) == 1;
    got += fscanf(fp, "%x", &mode) == 1;
    (void)got;
    fclose(fp);
    out->mode = (int)mode;
    derive_grid(out);
    return FDTD_OK;
}

int fdtd_field_sizes(const fdtd_params *p, size_t out[6])
{
    if (!p || !out) { fdtd_set_error("fdtd_field_sizes: NULL argument"); return FDTD_E_ARG; }
    out[0] = p->maxi * (p->maxj + 1) * (p->maxk + 1);
    out[1] = (p->maxi + 1) * p->maxj * (p->maxk + 1);
    out[2] = (p->maxi + 1) * (p->maxj + 1) * p->maxk;
    out[3] = (p->maxi + 1) * p->maxj * p->maxk;
    out[4] = p->maxi * (p->maxj + 1) * p->maxk;
    out[5] = p->maxi * p->maxj * (p->maxk + 1);
    return FDTD_OK;
}

/* main.c:765: `for (t = 0; t <= simulation_time; t += time_step)`. */
int fdtd_step_count(const fdtd_params *p, size_t *out)
{
    size_t n = 0;
    double t;
    if (!p || !out) { fdtd_set_error("fdtd_step_count: NULL argument"); return FDTD_E_ARG; }
    if (!(p->time_step > 0.0)) { fdtd_set_error("fdtd_step_count: time_step must be > 0"); return FDTD_E_ARG; }
    for (t = 0; t <= p->simulation_time; t += p->time_step)
        ++n;
    *out = n;
    return FDTD_OK;
}

/* main.c:720-739 */
int fdtd_source_plan_make(const fdtd_params *p, fdtd_source_plan *out)
{
    const double aprime = 0.005, bprime = 0.005;
    double min_y, max_y, min_x, max_x, f_mnl, omega;
    if (!p || !out) { fdtd_set_error("fdtd_source_plan_make: NULL argument"); return FDTD_E_ARG; }
    min_y = p->width / 2. - aprime / 2.;
    max_y = min_y + aprime;
    min_x = p->length / 2. - bprime / 2.;
    max_x = min_x + bprime;
    out->j0 = (long)((int)(min_y / p->spatial_step) - 1);
    out->j1 = (long)((int)(max_y / p->spatial_step) + 1);
    out->i0 = (long)((int)(min_x / p->spatial_step) - 1);
    out->i1 = (long)((int)(max_x / p->spatial_step) + 1);
    f_mnl = 0.5 * FDTD_CELERITY * sqrt(pow(FDTD_PI / p->width, 2) + pow(FDTD_PI / p->length, 2)) / FDTD_PI;
    omega = 2.0 * FDTD_PI * f_mnl;
    out->z_te = (omega * FDTD_MU) /
                sqrt(pow(omega, 2) * FDTD_MU * FDTD_EPSILON - pow(FDTD_PI / p->width, 2));
    return FDTD_OK;
}

/* main.c:748 and :751, products left to right; f = 2.45e10 as coded at main.c:735. */
int fdtd_source_values(const fdtd_params *p, const fdtd_source_plan *plan, double t,
                       double *ez_vals, double *hx_vals)
{
    const double aprime = 0.005;
    const double f = 2.45e10;
    long n, s;
    if (!p || !plan || !ez_vals || !hx_vals) { fdtd_set_error("fdtd_source_values: NULL argument"); return FDTD_E_ARG; }
    n = plan->i1 - plan->i0;
    for (s = 0; s < n; ++s) {
        const size_t shift_i = (size_t)s;
        ez_vals[s] = sin(2 * FDTD_PI * f * t) * sin(FDTD_PI * (shift_i * p->spatial_step) / aprime);
        hx_vals[s] = -(1.0 / plan->z_te) * sin(2 * FDTD_PI * f * t) *
                     sin(FDTD_PI * (shift_i * p->spatial_step) / aprime);
    }
    return FDTD_OK;
}

/* main.c:416-424, node planes [k_first, k_first + nplanes) only (a slab never needs more).
 * The product sin(pi k dx / height) * sin(pi i dx / length) has only one factor per plane and one
 * per column; each is evaluated once with the same expression and the same libm as the reference,
 * so every product is bit-identical to the reference's. */
int fdtd_initial_conditions_planes(const fdtd_params *p, size_t k_first, size_t nplanes, double *Ey)
{
    size_t i, j, k;
    double *si;
    if (!p || !Ey) { fdtd_set_error("fdtd_initial_conditions_host: NULL argument"); return FDTD_E_ARG; }
    if (k_first + nplanes > p->maxk + 1) { fdtd_set_error("fdtd_initial_conditions_host: planes out of range"); return FDTD_E_ARG; }
    si = (double *)malloc(sizeof(double) * (p->maxi + 1));
    if (!si) { fdtd_set_error("fdtd_initial_conditions_host: out of memory"); return FDTD_E_NOMEM; }
    for (i = 0; i < p->maxi + 1; ++i)
        si[i] = sin(FDTD_PI * i * p->spatial_step / p->length);
    for (k = k_first; k < k_first + nplanes; ++k) {
        const double sk = sin(FDTD_PI * k * p->spatial_step / p->height);
        for (j = 0; j < p->maxj; ++j) {
            double *row = Ey + (p->maxi + 1) * (j + p->maxj * (k - k_first));
            for (i = 0; i < p->maxi + 1; ++i)
                row[i] = sk * si[i];
        }
    }
    free(si);
    return FDTD_OK;
}

int fdtd_initial_conditions_host(const fdtd_params *p, double *Ey)
{
    if (!p) { fdtd_set_error("fdtd_initial_conditions_host: NULL argument"); return FDTD_E_ARG; }
    return fdtd_initial_conditions_planes(p, 0, p->maxk + 1, Ey);
}

int fdtd_slab_range(size_t maxk, int rank, int nranks, size_t *k0, size_t *k1)
{
    size_t base, extra, r;
    if (nranks < 1 || rank < 0 || rank >= nranks || !k0 || !k1) {
        fdtd_set_error("fdtd_slab_range: bad rank %d of %d", rank, nranks);
        return FDTD_E_ARG;
    }
    base = maxk / (size_t)nranks;
    extra = maxk % (size_t)nranks;
    r = (size_t)rank;
    *k0 = r * base + (r < extra ? r : extra);
    *k1 = *k0 + base + (r < extra ? 1 : 0);
    return FDTD_OK;
}
