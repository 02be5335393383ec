/*
 * fdtd_oracle.c -- CPU restatement of the reference's FDTD hot path (test infrastructure only;
 * see fdtd_oracle.h for the rules on who may call it and for the parity pin).
 *
 * The arithmetic of every statement is kept in the reference's operand order
 * (SURVEY.md Appendix B.2): t1 = a - b; t2 = c - d; t3 = t1 - t2; t4 = factor * t3; F = F + t4.
 * Build with -std=c99 so GCC does not contract a*b+c into an FMA (oracle/Makefile).
 */
#include "fdtd_oracle.h"

#include <math.h>
#include <stdio.h>

/* main.c:22-25 -- literal values, including the truncated epsilon_0 */
#define ORACLE_MU 1.25663706143591729538505735331180115367886775975E-6
#define ORACLE_EPSILON 8.854E-12
#define ORACLE_PI 3.14159265358979323846264338327950288419716939937510582097494
#define ORACLE_CELERITY 299792458.0

/* Row length / plane size of each staggered array (main.c:379-407):
 *   ex: nx     x (ny+1) x (nz+1)      hx: (nx+1) x ny     x nz
 *   ey: (nx+1) x ny     x (nz+1)      hy: nx     x (ny+1) x nz
 *   ez: (nx+1) x (ny+1) x nz          hz: nx     x ny     x (nz+1)            */
typedef struct shape { size_t row, plane; } shape;

static shape shape_ex(const oracle_params *p) { shape s = {p->nx, p->nx * (p->ny + 1)}; return s; }
static shape shape_ey(const oracle_params *p) { shape s = {p->nx + 1, (p->nx + 1) * p->ny}; return s; }
static shape shape_ez(const oracle_params *p) { shape s = {p->nx + 1, (p->nx + 1) * (p->ny + 1)}; return s; }
static shape shape_hx(const oracle_params *p) { return shape_ey(p); }
static shape shape_hy(const oracle_params *p) { return shape_ex(p); }
static shape shape_hz(const oracle_params *p) { shape s = {p->nx, p->nx * p->ny}; return s; }

void oracle_field_sizes(const oracle_params *p, size_t out[6])
{
    out[0] = p->nx * (p->ny + 1) * (p->nz + 1);
    out[1] = (p->nx + 1) * p->ny * (p->nz + 1);
    out[2] = (p->nx + 1) * (p->ny + 1) * p->nz;
    out[3] = (p->nx + 1) * p->ny * p->nz;
    out[4] = p->nx * (p->ny + 1) * p->nz;
    out[5] = p->nx * p->ny * (p->nz + 1);
}

/* main.c:237-239: float size promoted to double, divided by the double step, truncated. */
static void derive_grid(oracle_params *p)
{
    p->nx = (size_t)(p->length / p->spatial_step);
    p->ny = (size_t)(p->width / p->spatial_step);
    p->nz = (size_t)(p->height / p->spatial_step);
}

void oracle_make_params(float length, float width, float height, double dx, double dt,
                        float simulation_time, unsigned sampling_rate, int mode, oracle_params *p)
{
    p->length = length;
    p->width = width;
    p->height = height;
    p->spatial_step = dx;
    p->time_step = dt;
    p->simulation_time = simulation_time;
    p->sampling_rate = sampling_rate;
    p->mode = mode;
    derive_grid(p);
}

/* main.c:216-242: eight whitespace-separated numbers; conversions %f %f %f %lf %lf %f %u %x;
 * scanf results are not checked by the reference, so a short file leaves zeros here. */
int oracle_load_parameters(const char *path, oracle_params *p)
{
    FILE *fp = fopen(path, "r");
    unsigned mode = 0;
    if (!fp)
        return -1;
    oracle_make_params(0.f, 0.f, 0.f, 0., 0., 0.f, 0u, 0, p);
    if (fscanf(fp, "%f", &p->length) != 1) { /* ignored, as in the reference */ }
    if (fscanf(fp, "%f", &p->width) != 1) {}
    if (fscanf(fp, "%f", &p->height) != 1) {}
    if (fscanf(fp, "%lf", &p->spatial_step) != 1) {}
    if (fscanf(fp, "%lf", &p->time_step) != 1) {}
    if (fscanf(fp, "%f", &p->simulation_time) != 1) {}
    if (fscanf(fp, "%u", &p->sampling_rate) != 1) {}
    if (fscanf(fp, "%x", &mode) != 1) {}
    fclose(fp);
    p->mode = (int)mode;
    derive_grid(p);
    return 0;
}

/* main.c:765: double counter, repeated addition, float bound promoted to double, `<=`. */
size_t oracle_step_count(const oracle_params *p)
{
    size_t n = 0;
    double t;
    for (t = 0; t <= p->simulation_time; t += p->time_step)
        ++n;
    return n;
}

/* main.c:416-424 */
void oracle_set_initial_conditions(const oracle_params *p, double *ey)
{
    const shape s = shape_ey(p);
    for (size_t k = 0; k <= p->nz; ++k) {
        for (size_t j = 0; j < p->ny; ++j) {
            double *row = ey + k * s.plane + j * s.row;
            for (size_t i = 0; i <= p->nx; ++i)
                row[i] = sin(ORACLE_PI * k * p->spatial_step / p->height) *
                         sin(ORACLE_PI * i * p->spatial_step / p->length);
        }
    }
}

/* main.c:431-462.  Three sweeps, one per H component, each over that component's full extent. */
void oracle_update_h(const oracle_params *p, const oracle_fields *f)
{
    const size_t nx = p->nx, ny = p->ny, nz = p->nz;
    const shape sex = shape_ex(p), sey = shape_ey(p), sez = shape_ez(p);
    const shape shx = shape_hx(p), shy = shape_hy(p), shz = shape_hz(p);
    const double c = p->time_step / (ORACLE_MU * p->spatial_step); /* main.c:441 */

    /* Hx: main.c:445-449 */
    for (size_t k = 0; k < nz; ++k)
        for (size_t j = 0; j < ny; ++j) {
            double *h = f->hx + k * shx.plane + j * shx.row;
            const double *ey_up = f->ey + (k + 1) * sey.plane + j * sey.row;
            const double *ey_lo = f->ey + k * sey.plane + j * sey.row;
            const double *ez_up = f->ez + k * sez.plane + (j + 1) * sez.row;
            const double *ez_lo = f->ez + k * sez.plane + j * sez.row;
            for (size_t i = 0; i <= nx; ++i)
                h[i] = h[i] + c * ((ey_up[i] - ey_lo[i]) - (ez_up[i] - ez_lo[i]));
        }

    /* Hy: main.c:451-455 */
    for (size_t k = 0; k < nz; ++k)
        for (size_t j = 0; j <= ny; ++j) {
            double *h = f->hy + k * shy.plane + j * shy.row;
            const double *ez = f->ez + k * sez.plane + j * sez.row;
            const double *ex_up = f->ex + (k + 1) * sex.plane + j * sex.row;
            const double *ex_lo = f->ex + k * sex.plane + j * sex.row;
            for (size_t i = 0; i < nx; ++i)
                h[i] = h[i] + c * ((ez[i + 1] - ez[i]) - (ex_up[i] - ex_lo[i]));
        }

    /* Hz: main.c:457-461 */
    for (size_t k = 0; k <= nz; ++k)
        for (size_t j = 0; j < ny; ++j) {
            double *h = f->hz + k * shz.plane + j * shz.row;
            const double *ex_up = f->ex + k * sex.plane + (j + 1) * sex.row;
            const double *ex_lo = f->ex + k * sex.plane + j * sex.row;
            const double *ey = f->ey + k * sey.plane + j * sey.row;
            for (size_t i = 0; i < nx; ++i)
                h[i] = h[i] + c * ((ex_up[i] - ex_lo[i]) - (ey[i + 1] - ey[i]));
        }
}

/* main.c:469-500.  The loop bounds leave the tangential E on the six faces untouched: that is
 * the PEC wall. */
void oracle_update_e(const oracle_params *p, const oracle_fields *f)
{
    const size_t nx = p->nx, ny = p->ny, nz = p->nz;
    const shape sex = shape_ex(p), sey = shape_ey(p), sez = shape_ez(p);
    const shape shx = shape_hx(p), shy = shape_hy(p), shz = shape_hz(p);
    const double c = p->time_step / (ORACLE_EPSILON * p->spatial_step); /* main.c:479 */

    /* Ex: main.c:483-487 */
    for (size_t k = 1; k < nz; ++k)
        for (size_t j = 1; j < ny; ++j) {
            double *e = f->ex + k * sex.plane + j * sex.row;
            const double *hz_hi = f->hz + k * shz.plane + j * shz.row;
            const double *hz_lo = f->hz + k * shz.plane + (j - 1) * shz.row;
            const double *hy_hi = f->hy + k * shy.plane + j * shy.row;
            const double *hy_lo = f->hy + (k - 1) * shy.plane + j * shy.row;
            for (size_t i = 0; i < nx; ++i)
                e[i] = e[i] + c * ((hz_hi[i] - hz_lo[i]) - (hy_hi[i] - hy_lo[i]));
        }

    /* Ey: main.c:489-493 */
    for (size_t k = 1; k < nz; ++k)
        for (size_t j = 0; j < ny; ++j) {
            double *e = f->ey + k * sey.plane + j * sey.row;
            const double *hx_hi = f->hx + k * shx.plane + j * shx.row;
            const double *hx_lo = f->hx + (k - 1) * shx.plane + j * shx.row;
            const double *hz = f->hz + k * shz.plane + j * shz.row;
            for (size_t i = 1; i < nx; ++i)
                e[i] = e[i] + c * ((hx_hi[i] - hx_lo[i]) - (hz[i] - hz[i - 1]));
        }

    /* Ez: main.c:495-499 */
    for (size_t k = 0; k < nz; ++k)
        for (size_t j = 1; j < ny; ++j) {
            double *e = f->ez + k * sez.plane + j * sez.row;
            const double *hy = f->hy + k * shy.plane + j * shy.row;
            const double *hx_hi = f->hx + k * shx.plane + j * shx.row;
            const double *hx_lo = f->hx + k * shx.plane + (j - 1) * shx.row;
            for (size_t i = 1; i < nx; ++i)
                e[i] = e[i] + c * ((hy[i] - hy[i - 1]) - (hx_hi[i] - hx_lo[i]));
        }
}

/* main.c:720-733.  The reference keeps the bounds in doubles holding integer values and walks
 * them with size_t counters; for a patch inside the grid that is this integer range. */
void oracle_source_bounds(const oracle_params *p, long b[4])
{
    const double aprime = 0.005, bprime = 0.005;
    const double min_y = p->width / 2. - aprime / 2.;
    const double max_y = min_y + aprime;
    const double min_x = p->length / 2. - bprime / 2.;
    const double max_x = min_x + bprime;
    b[0] = (long)((int)(min_x / p->spatial_step) - 1);
    b[1] = (long)((int)(max_x / p->spatial_step) + 1);
    b[2] = (long)((int)(min_y / p->spatial_step) - 1);
    b[3] = (long)((int)(max_y / p->spatial_step) + 1);
}

/* main.c:737-739: wave impedance of the feeding guide, from width and length (not from f). */
double oracle_source_zte(const oracle_params *p)
{
    const double f_mnl = 0.5 * ORACLE_CELERITY *
                         sqrt(pow(ORACLE_PI / p->width, 2) + pow(ORACLE_PI / p->length, 2)) / ORACLE_PI;
    const double omega = 2.0 * ORACLE_PI * f_mnl;
    return (omega * ORACLE_MU) /
           sqrt(pow(omega, 2) * ORACLE_MU * ORACLE_EPSILON - pow(ORACLE_PI / p->width, 2));
}

/* main.c:712-753: hard source on the k = 0 plane; the profile depends on the x offset only. */
void oracle_set_source(const oracle_params *p, const oracle_fields *f, double t)
{
    const double aprime = 0.005;
    const double freq = 2.45e10; /* main.c:735, as coded */
    const double z_te = oracle_source_zte(p);
    const shape sex = shape_ex(p), sez = shape_ez(p), shx = shape_hx(p), shz = shape_hz(p);
    long b[4];
    oracle_source_bounds(p, b);

    for (long i = b[0]; i < b[1]; ++i) {
        const size_t s = (size_t)(i - b[0]);
        for (long j = b[2]; j < b[3]; ++j) {
            f->ez[(size_t)i + (size_t)j * sez.row] =
                sin(2 * ORACLE_PI * freq * t) * sin(ORACLE_PI * (s * p->spatial_step) / aprime);
            f->ex[(size_t)i + (size_t)j * sex.row] = 0;
            f->hz[(size_t)i + (size_t)j * shz.row] = 0;
            f->hx[(size_t)i + (size_t)j * shx.row] =
                -(1.0 / z_te) * sin(2 * ORACLE_PI * freq * t) *
                sin(ORACLE_PI * (s * p->spatial_step) / aprime);
        }
    }
}

/* main.c:765-779 loop body, `steps` times */
void oracle_run(const oracle_params *p, const oracle_fields *f, size_t steps, double *t_io)
{
    double t = *t_io;
    for (size_t n = 0; n < steps; ++n, t += p->time_step) {
        if (p->mode == 1)
            oracle_set_source(p, f, t);
        oracle_update_h(p, f);
        if (p->mode == 1)
            oracle_set_source(p, f, t);
        oracle_update_e(p, f);
    }
    *t_io = t;
}

/* main.c:511-521 / 532-540 through the generic index at main.c:374-377.  (oi,oj,ok) is the unit
 * vector of the component for H and its complement for E (main.c:563-578).  The E average is
 * kept as coded: points (0,0,0), (oi,oj,ok), (0,oj,ok), (oi,0,ok) -- for ex and ey that repeats
 * one corner and misses another. */
void oracle_aggregate(const oracle_params *p, const oracle_fields *f, int var, double *out)
{
    static const size_t off[6][3] = {{0, 1, 1}, {1, 0, 1}, {1, 1, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double *src[6] = {f->ex, f->ey, f->ez, f->hx, f->hy, f->hz};
    const double *a = src[var];
    const size_t oi = off[var][0], oj = off[var][1], ok = off[var][2];
    const size_t row = p->nx + oi, plane = row * (p->ny + oj);
    size_t n = 0;

    for (size_t k = 0; k < p->nz; ++k)
        for (size_t j = 0; j < p->ny; ++j)
            for (size_t i = 0; i < p->nx; ++i, ++n) {
                const size_t here = i + j * row + k * plane;
                if (var < 3)
                    out[n] = .25 * (a[here] +
                                    a[here + oi + oj * row + ok * plane] +
                                    a[here + oj * row + ok * plane] +
                                    a[here + oi + ok * plane]);
                else
                    out[n] = .5 * (a[here] + a[here + oi + oj * row + ok * plane]);
            }
}

/* main.c:670-710 */
void oracle_validation_fields(const oracle_params *p, const oracle_fields *f,
                              double *vey, double *vhx, double *vhz, double t)
{
    const double f_mnl = 0.5 * ORACLE_CELERITY *
                         sqrt(pow(ORACLE_PI / p->height, 2) + pow(ORACLE_PI / p->length, 2)) / ORACLE_PI;
    const double omega = 2.0 * ORACLE_PI * f_mnl;
    const double z_te = (omega * ORACLE_MU) /
                        sqrt(pow(omega, 2) * ORACLE_MU * ORACLE_EPSILON - pow(ORACLE_PI / p->length, 2));
    const shape sey = shape_ey(p), shx = shape_hx(p), shz = shape_hz(p);
    const double dx = p->spatial_step;

    for (size_t k = 0; k <= p->nz; ++k)
        for (size_t j = 0; j < p->ny; ++j)
            for (size_t i = 0; i <= p->nx; ++i) {
                const size_t n = i + j * sey.row + k * sey.plane;
                vey[n] = (cos(2 * ORACLE_PI * f_mnl * t) *
                          sin(ORACLE_PI * k * dx / p->height) *
                          sin(ORACLE_PI * i * dx / p->length)) - f->ey[n];
            }

    for (size_t k = 0; k < p->nz; ++k)
        for (size_t j = 0; j < p->ny; ++j)
            for (size_t i = 0; i <= p->nx; ++i) {
                const size_t n = i + j * shx.row + k * shx.plane;
                vhx[n] = ((1.0 / z_te) *
                          sin(2 * ORACLE_PI * f_mnl * t) *
                          sin(ORACLE_PI * k * dx / p->height) *
                          cos(ORACLE_PI * i * dx / p->length)) - f->hx[n];
            }

    for (size_t k = 0; k <= p->nz; ++k)
        for (size_t j = 0; j < p->ny; ++j)
            for (size_t i = 0; i < p->nx; ++i) {
                const size_t n = i + j * shz.row + k * shz.plane;
                vhz[n] = (-ORACLE_PI / (omega * ORACLE_MU * p->length) *
                          sin(2 * ORACLE_PI * f_mnl * t) *
                          cos(ORACLE_PI * k * dx / p->height) *
                          sin(ORACLE_PI * i * dx / p->length)) - f->hz[n];
            }
}

/* main.c:602-668, as coded: the Ez average at main.c:627 indexes Ez with Hz's strides (kHz).
 * out[0] = electric, out[1] = magnetic energy. */
void oracle_energy(const oracle_params *p, const oracle_fields *f, double out[2])
{
    const shape sex = shape_ex(p), sey = shape_ey(p), shx = shape_hx(p), shy = shape_hy(p), shz = shape_hz(p);
    const double dv = pow(p->spatial_step, 3);
    double ex = 0.0, ey = 0.0, ez = 0.0, hx = 0.0, hy = 0.0, hz = 0.0;
    for (size_t k = 0; k < p->nz; k++)
        for (size_t j = 0; j < p->ny; j++)
            for (size_t i = 0; i < p->nx; i++) {
                const size_t oex = i + j * sex.row + k * sex.plane, oey = i + j * sey.row + k * sey.plane;
                const size_t ohz = i + j * shz.row + k * shz.plane; /* used for Ez too, main.c:627 */
                const double mex = (f->ex[oex] + f->ex[oex + sex.plane] + f->ex[oex + sex.row] +
                                    f->ex[oex + sex.row + sex.plane]) / 4.;
                const double mey = (f->ey[oey] + f->ey[oey + 1] + f->ey[oey + sey.plane] +
                                    f->ey[oey + 1 + sey.plane]) / 4.;
                const double mez = (f->ez[ohz] + f->ez[ohz + shz.row] + f->ez[ohz + 1] + f->ez[ohz + 1 + shz.row]) / 4.;
                const size_t ohx = i + j * shx.row + k * shx.plane, ohy = i + j * shy.row + k * shy.plane;
                const double mhx = (f->hx[ohx] + f->hx[ohx + 1]) / 2.;
                const double mhy = (f->hy[ohy] + f->hy[ohy + shy.row]) / 2.;
                const double mhz = (f->hz[ohz] + f->hz[ohz + shz.plane]) / 2.;
                ex += pow(mex, 2) * dv; ey += pow(mey, 2) * dv; ez += pow(mez, 2) * dv;
                hx += pow(mhx, 2) * dv; hy += pow(mhy, 2) * dv; hz += pow(mhz, 2) * dv;
            }
    out[0] = (ex + ey + ez) * ORACLE_EPSILON / 2.;
    out[1] = (hx + hy + hz) * ORACLE_MU / 2.;
}
