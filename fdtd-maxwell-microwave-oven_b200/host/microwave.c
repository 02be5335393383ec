/*
 * microwave.c -- C99 host program with the reference's command line and console output
 * (main(), main.c:807-853), driving the hot path on a B200 through the C ABI of
 * libfdtd_b200.so instead of the CPU loops.
 *
 *     ./microwave params.txt          (same 8-number file, main.c:216-242)
 *
 * Output.  The reference writes one Silo file per dump, r/result%04d.silo (main.c:19, :550-598).
 * libsilo is not available in this image, so the dump sink below writes the same variables --
 * same names, same order, same zone-centred doubles, x fastest -- as a raw brick plus one
 * VisIt "BOV" header per variable:  r/result%04d.raw  and  r/result%04d.<var>.bov .
 * The sink is a three-function table (fdtd_dump_sink); a Silo-backed sink is a drop-in where
 * libsilo exists.  As in the reference the directory r/ must already exist;
 * if it does not, the run fails with the reference's message "Could not create DB".
 *
 * Environment: FDTD_B200_DEVICE (CUDA device index, default 0), FDTD_B200_GPUS=N (split the cavity
 * into N z-slabs on GPUs 0..N-1 of this box, or on the devices FDTD_B200_DEVICES=a,b,... names; still
 * one process, one thread: fdtd_group_*; every slab writes its own brick r/result%04d.slab<r>.raw
 * whose BOV header carries the z origin),
 * FDTD_B200_NO_DUMPS=1 (step without writing anything), FDTD_B200_REPORT=1 (timing summary on stderr).
 */
#define _POSIX_C_SOURCE 199309L
#include "fdtd_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* fail(), main.c:154-159: perror + exit(EXIT_FAILURE) */
static void fail(const char *msg)
{
    perror(msg);
    exit(EXIT_FAILURE);
}

static void fail_lib(const char *what)
{
    fprintf(stderr, "%s: %s\n", what, fdtd_last_error());
    exit(EXIT_FAILURE);
}

typedef struct file_sink {
    int slab, nslabs;
    FILE *raw;
    char base[128];
    size_t dims[3];
    size_t k0;
    size_t offset;
    int nvars;
    char names[16][8];
    size_t offsets[16];
    double dx;
} file_sink;

/* <-> DBCreate + DBPutQuadmesh, main.c:553-561 */
static int sink_begin(void *user, int iteration, const size_t dims[3], size_t k0)
{
    file_sink *s = (file_sink *)user;
    char name[160];
    if (s->nslabs > 1)
        snprintf(s->base, sizeof s->base, "r/result%04d.slab%d", iteration, s->slab);
    else
        snprintf(s->base, sizeof s->base, "r/result%04d", iteration);
    snprintf(name, sizeof name, "%s.raw", s->base);
    s->raw = fopen(name, "wb");
    if (!s->raw)
        return -1;
    memcpy(s->dims, dims, sizeof s->dims);
    s->k0 = k0;
    s->offset = 0;
    s->nvars = 0;
    return 0;
}

/* <-> DBPutQuadvar1, main.c:564-588 */
static int sink_variable(void *user, const char *name, const double *data, size_t count)
{
    file_sink *s = (file_sink *)user;
    if (s->nvars >= 16 || fwrite(data, sizeof(double), count, s->raw) != count)
        return -1;
    snprintf(s->names[s->nvars], sizeof s->names[0], "%s", name);
    s->offsets[s->nvars] = s->offset;
    s->offset += count * sizeof(double);
    s->nvars++;
    return 0;
}

/* <-> DBPutDefvars + DBClose, main.c:591-597 */
static int sink_end(void *user)
{
    file_sink *s = (file_sink *)user;
    int v;
    if (fclose(s->raw) != 0)
        return -1;
    s->raw = NULL;
    for (v = 0; v < s->nvars; ++v) {
        char name[200];
        FILE *h;
        snprintf(name, sizeof name, "%s.%s.bov", s->base, s->names[v]);
        h = fopen(name, "w");
        if (!h)
            return -1;
        fprintf(h, "DATA_FILE: %s.raw\nDATA_SIZE: %zu %zu %zu\nDATA_FORMAT: DOUBLE\nVARIABLE: %s\n"
                   "DATA_ENDIAN: LITTLE\nCENTERING: zonal\nBYTE_OFFSET: %zu\n"
                   "BRICK_ORIGIN: 0. 0. %.17g\nBRICK_SIZE: %.17g %.17g %.17g\n",
                strrchr(s->base, '/') + 1, s->dims[0], s->dims[1], s->dims[2], s->names[v], s->offsets[v],
                (double)s->k0 * s->dx, (double)s->dims[0] * s->dx, (double)s->dims[1] * s->dx,
                (double)s->dims[2] * s->dx);
        fclose(h);
    }
    return 0;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, const char *argv[])
{
    fdtd_params params;
    fdtd_ctx *ctx = NULL;
    fdtd_group *group = NULL;
    static file_sink files[64];
    fdtd_dump_sink sinks[64];
    size_t steps = 0;
    double t_end = 0.0, t0, t1;
    const char *dev_env = getenv("FDTD_B200_DEVICE");
    const char *gpus_env = getenv("FDTD_B200_GPUS");
    const char *no_dumps = getenv("FDTD_B200_NO_DUMPS");
    const char *report = getenv("FDTD_B200_REPORT");
    const char *devs_env = getenv("FDTD_B200_DEVICES");
    const int ngpus = gpus_env ? atoi(gpus_env) : 1;
    int devices[64];
    int rc, r;

    printf("Welcome into our microwave oven eletrico-magnetic field simulator! \n");

    if (argc != 2)
        fail("This program needs 1 argument: the parameters file (.txt). Eg.: ./microwave param.txt");

    printf("Loading the parameters...\n");
    rc = fdtd_load_parameters(argv[1], &params);
    if (rc == FDTD_E_IO)
        fail("Unable to open parameters file!");
    if (rc != FDTD_OK)
        fail_lib("load_parameters");
    if (params.time_step > params.simulation_time)
        fail("The time step must be lower than the simulation time!");

    printf("Initializing fields\n");
    if (ngpus < 1 || ngpus > 64) {
        fprintf(stderr, "FDTD_B200_GPUS must be between 1 and 64\n");
        return EXIT_FAILURE;
    }
    for (r = 0; r < 64; ++r)
        devices[r] = r;
    if (devs_env) { /* comma-separated device index per slab; indices may repeat */
        const char *q = devs_env;
        for (r = 0; r < 64 && *q; ++r) {
            devices[r] = atoi(q);
            q += strcspn(q, ",");
            if (*q == ',')
                ++q;
        }
    }
    if (ngpus > 1) {
        if (fdtd_group_create(&params, ngpus, devices, &group) != FDTD_OK)
            fail_lib("initialize_fields");
    } else if (fdtd_ctx_create(&params, dev_env ? atoi(dev_env) : 0, &ctx) != FDTD_OK) {
        fail_lib("initialize_fields");
    }
    if (params.mode == 0)
        printf("Validation mode activated. \n");

    printf("Creating mesh\n");

    printf("Setting initial conditions\n");
    if (params.mode == 0 &&
        (group ? fdtd_group_set_initial_conditions(group) : fdtd_set_initial_conditions(ctx)) != FDTD_OK)
        fail_lib("set_initial_conditions");

    printf("Launching simulation\n");
    fflush(stdout);
    memset(files, 0, sizeof files);
    for (r = 0; r < ngpus; ++r) {
        files[r].slab = r;
        files[r].nslabs = ngpus;
        files[r].dx = params.spatial_step;
        sinks[r].user = &files[r];
        sinks[r].begin = sink_begin;
        sinks[r].variable = sink_variable;
        sinks[r].end = sink_end;
    }
    t0 = now_s();
    if (group)
        rc = fdtd_group_propagate(group, (no_dumps && no_dumps[0] == '1') ? NULL : sinks, &steps, &t_end);
    else
        rc = fdtd_propagate(ctx, (no_dumps && no_dumps[0] == '1') ? NULL : &sinks[0], &steps, &t_end);
    t1 = now_s();
    if (rc == FDTD_E_IO)
        fail("Could not create DB\n"); /* main.c:556-559 */
    if (rc != FDTD_OK)
        fail_lib("propagate_fields");
    if (report && report[0] == '1')
        fprintf(stderr, "[fdtd_b200] %zu x %zu x %zu cells on %d GPU(s), %zu steps in %.3f s: %.3f Gcell-updates/s\n",
                params.maxi, params.maxj, params.maxk, ngpus, steps, t1 - t0,
                1e-9 * (double)params.maxi * (double)params.maxj * (double)params.maxk * (double)steps / (t1 - t0));

    printf("Freeing memory...\n");
    if (group)
        fdtd_group_destroy(group);
    else
        fdtd_ctx_destroy(ctx);

    printf("Simulation complete!\n");
    return 0;
}
