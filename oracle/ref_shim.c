/*
 * ref_shim.c -- exposes the UNMODIFIED reference (/root/reference/main.c) as a shared library
 * for parity checks.  Test infrastructure; never linked into the product.
 *
 * main.c is included textually from where it lies (the Makefile passes -I/root/reference and
 * -Ioracle/silo_stub for its <silo.h>); nothing of it is copied into this repository.  Its
 * `main` is renamed so the file can live inside a library.  The wrappers below only marshal
 * plain arguments into the reference's own structs and call the reference's own functions.
 * Output goes to oracle/_ref/ (git-ignored).
 */
#define main reference_main
#include "main.c"
#undef main

#include "fdtd_oracle.h" /* oracle_params / oracle_fields: the plain structs shared with the tests */

/* The reference's allocation list leaves the first node's `previous` link unset (main.c:170-173),
 * so its freeAll() walks into whatever malloc handed back.  In a fresh process that is zeroed
 * heap; inside a long-lived test process it is not.  Seeding the list with a zeroed node before
 * the reference allocates keeps its own code unmodified and its teardown well defined. */
static void seed_allocation_list(void)
{
    if (!allocatedLs)
        allocatedLs = calloc(1, sizeof(ChainedAllocated));
}

static Parameters to_ref(const oracle_params *q)
{
    Parameters p;
    p.length = q->length;
    p.width = q->width;
    p.height = q->height;
    p.spatial_step = q->spatial_step;
    p.time_step = q->time_step;
    p.simulation_time = q->simulation_time;
    p.sampling_rate = q->sampling_rate;
    p.mode = (MODE)q->mode;
    p.maxi = q->nx;
    p.maxj = q->ny;
    p.maxk = q->nz;
    return p;
}

static Fields to_ref_fields(const oracle_fields *g)
{
    Fields f;
    f.Ex = g->ex; f.Ey = g->ey; f.Ez = g->ez;
    f.Hx = g->hx; f.Hy = g->hy; f.Hz = g->hz;
    return f;
}

/* reference load_parameters (main.c:216), result copied out, reference allocation released */
int ref_load_parameters(const char *path, oracle_params *out)
{
    FILE *probe = fopen(path, "r");
    Parameters *p;
    if (!probe)
        return -1; /* the reference would exit() here */
    fclose(probe);
    seed_allocation_list();
    p = load_parameters(path);
    out->length = p->length;
    out->width = p->width;
    out->height = p->height;
    out->spatial_step = p->spatial_step;
    out->time_step = p->time_step;
    out->simulation_time = p->simulation_time;
    out->sampling_rate = p->sampling_rate;
    out->mode = (int)p->mode;
    out->nx = p->maxi;
    out->ny = p->maxj;
    out->nz = p->maxk;
    freeAll();
    return 0;
}

void ref_set_initial_conditions(const oracle_params *q, double *ey)
{
    Parameters p = to_ref(q);
    set_initial_conditions(ey, &p);
}

void ref_update_h(const oracle_params *q, const oracle_fields *g)
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    update_H_field(&p, &f);
}

void ref_update_e(const oracle_params *q, const oracle_fields *g)
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    update_E_field(&p, &f);
}

void ref_set_source(const oracle_params *q, const oracle_fields *g, double t)
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    set_source(&p, &f, t);
}

/* the loop body of main.c:770-779, `steps` times, time accumulated like main.c:765 */
void ref_run(const oracle_params *q, const oracle_fields *g, size_t steps, double *t_io)
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    double t = *t_io;
    for (size_t n = 0; n < steps; ++n, t += p.time_step) {
        if (p.mode == COMPUTATION_MODE)
            set_source(&p, &f, t);
        update_H_field(&p, &f);
        if (p.mode == COMPUTATION_MODE)
            set_source(&p, &f, t);
        update_E_field(&p, &f);
    }
    *t_io = t;
}

void ref_aggregate(const oracle_params *q, const oracle_fields *g, int var, double *out)
{
    Parameters p = to_ref(q);
    switch (var) { /* argument triples of main.c:563-578 */
    case 0: aggregate_E_field(&p, g->ex, out, 0, 1, 1); break;
    case 1: aggregate_E_field(&p, g->ey, out, 1, 0, 1); break;
    case 2: aggregate_E_field(&p, g->ez, out, 1, 1, 0); break;
    case 3: aggregate_H_field(&p, g->hx, out, 1, 0, 0); break;
    case 4: aggregate_H_field(&p, g->hy, out, 0, 1, 0); break;
    default: aggregate_H_field(&p, g->hz, out, 0, 0, 1); break;
    }
}

void ref_validation_fields(const oracle_params *q, const oracle_fields *g,
                           double *vey, double *vhx, double *vhz, double t)
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    Fields v;
    v.Ex = v.Ez = v.Hy = NULL;
    v.Ey = vey; v.Hx = vhx; v.Hz = vhz;
    update_validation_fields_then_subfdtd(&p, &f, &v, t);
}

void ref_energy(const oracle_params *q, const oracle_fields *g, double out[2])
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    out[0] = calculate_E_energy(&f, &p);
    out[1] = calculate_H_energy(&f, &p);
}

/* The reference's whole stepping loop, propagate_fields (main.c:755-799), on caller-owned
 * arrays.  Every variable the reference hands to Silo is passed to `rec` (see silo_stub). */
void ref_propagate(const oracle_params *q, const oracle_fields *g, silo_stub_recorder rec)
{
    Parameters p = to_ref(q);
    Fields f = to_ref_fields(g);
    Fields v;
    Oven *oven;
    /* sizes as at main.c:310,332,354 */
    const size_t n_ey = (p.maxi + 1) * p.maxj * (p.maxk + 1);
    const size_t n_hx = (p.maxi + 1) * p.maxj * p.maxk;
    const size_t n_hz = p.maxi * p.maxj * (p.maxk + 1);
    seed_allocation_list();
    oven = compute_oven(&p);
    v.Ex = v.Ez = v.Hy = NULL;
    v.Ey = v.Hx = v.Hz = NULL;
    if (p.mode == VALIDATION_MODE) {
        v.Ey = Malloc(sizeof(double) * n_ey);
        v.Hx = Malloc(sizeof(double) * n_hx);
        v.Hz = Malloc(sizeof(double) * n_hz);
    }
    silo_stub_hook = rec;
    propagate_fields(&f, &v, &p, oven);
    silo_stub_hook = NULL;
    freeAll();
}
