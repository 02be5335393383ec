/*
 * microwave.c -- C99 host program with the reference's command line and console output
 * (main(), main.c:807-853), driving the hot path on a B200 through the C ABI of
 * libfdtd_b200.so instead of the CPU loops.
 *
 *     ./microwave params.txt          (same 8-number file, main.c:216-242)
 *
 * Output.  Like the reference: one Silo file per dump, r/result%04d.silo (main.c:19, :550-598), in the
 * PDB driver's format, holding the quadmesh "mesh" (collinear, coordinates i * dx, dims I+1, J+1, K+1),
 * the zone-centred double quadvars "ex" .. "hz" (+ "aEy", "aHx", "aHz" in validation mode) and the
 * defvars "vecs" (E = {ex, ey, ez}, H = {hx, hy, hz}).  libsilo is not available in this image, so the
 * files are written by host/silo_pdb.c, a minimal writer of that format.  With FDTD_B200_GPUS=N every
 * slab writes its block r/result%04d.slab<r>.silo and the root file r/result%04d.silo holds a
 * multimesh "mesh" and multivars "ex" .. over the blocks (Silo's multi-block convention).
 * FDTD_B200_SINK=raw selects the earlier output instead: a raw brick r/result%04d.raw plus one VisIt
 * "BOV" header per variable.  The sink is a three-function table (fdtd_dump_sink); a libsilo-backed
 * sink is a drop-in where libsilo exists.  As in the reference the directory r/ must
 * already exist; if it does not, the run fails with the reference's message "Could not create DB".
 *
 * Environment: FDTD_B200_DEVICE (CUDA device index, default 0), FDTD_B200_GPUS=N (split the cavity
 * into N z-slabs on GPUs 0..N-1 of this box, or on the devices FDTD_B200_DEVICES=a,b,... names; still
 * one process, one thread: fdtd_group_*; every slab writes its own brick r/result%04d.slab<r>.raw
 * whose BOV header carries the z origin),
 * FDTD_B200_NO_DUMPS=1 (step without writing anything), FDTD_B200_REPORT=1 (timing summary on stderr).
 */
#define _POSIX_C_SOURCE 199309L
#include "fdtd_b200.h"
#include "silo_pdb.h"

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

/* ---- the Silo sink ---------------------------------------------------------------------------- */

typedef struct silo_sink {
    int slab, nslabs;
    spdb_file *db;
    int iteration;
    size_t dims[3], k0;
    double dx;
    int nvars;
    char names[16][8];
} silo_sink;

/* <-> DBCreate + DBPutQuadmesh, main.c:553-561 (coordinates as compute_oven() makes them, main.c:270-278) */
static int silo_begin(void *user, int iteration, const size_t dims[3], size_t k0)
{
    silo_sink *s = (silo_sink *)user;
    char name[160];
    double *xyz[3];
    const double *cxyz[3];
    int mdims[3], d, rc;
    size_t i;
    if (s->nslabs > 1)
        snprintf(name, sizeof name, "r/result%04d.slab%d.silo", iteration, s->slab);
    else
        snprintf(name, sizeof name, "r/result%04d.silo", iteration);
    s->db = spdb_create(name, NULL);
    if (!s->db)
        return -1;
    s->iteration = iteration;
    memcpy(s->dims, dims, sizeof s->dims);
    s->k0 = k0;
    s->nvars = 0;
    for (d = 0; d < 3; ++d) {
        mdims[d] = (int)dims[d] + 1;
        xyz[d] = (double *)malloc(sizeof(double) * (dims[d] + 1));
        if (!xyz[d])
            return -1;
        for (i = 0; i < dims[d] + 1; ++i)
            xyz[d][i] = (double)(int)(i + (d == 2 ? k0 : 0)) * s->dx;
        cxyz[d] = xyz[d];
    }
    rc = spdb_put_quadmesh(s->db, "mesh", cxyz, mdims);
    for (d = 0; d < 3; ++d)
        free(xyz[d]);
    return rc;
}

/* <-> DBPutQuadvar1, main.c:564-588 */
static int silo_variable(void *user, const char *name, const double *data, size_t count)
{
    silo_sink *s = (silo_sink *)user;
    const int zdims[3] = {(int)s->dims[0], (int)s->dims[1], (int)s->dims[2]};
    if (s->nvars < 16)
        snprintf(s->names[s->nvars++], sizeof s->names[0], "%s", name);
    if (spdb_quadvar_begin(s->db, name, "mesh", zdims) != 0 || spdb_quadvar_append(s->db, data, count) != 0)
        return -1;
    return spdb_quadvar_end(s->db);
}

static int put_vecs(spdb_file *db)
{
    const char *names[] = {"E", "H"};
    const char *defs[] = {"{ex, ey, ez}", "{hx, hy, hz}"};
    const int types[] = {SPDB_VARTYPE_VECTOR, SPDB_VARTYPE_VECTOR};
    return spdb_put_defvars(db, "vecs", 2, names, types, defs); /* main.c:591-595 */
}

/* <-> DBPutDefvars + DBClose, main.c:591-597; slab 0 of a multi-slab run also writes the root file */
static int silo_end(void *user)
{
    silo_sink *s = (silo_sink *)user;
    int rc = put_vecs(s->db);
    if (spdb_close(s->db) != 0)
        rc = -1;
    s->db = NULL;
    if (rc == 0 && s->nslabs > 1 && s->slab == 0) {
        char name[160], blocks[64][96];
        const char *ptrs[64];
        spdb_file *root;
        int r, v;
        snprintf(name, sizeof name, "r/result%04d.silo", s->iteration);
        root = spdb_create(name, NULL);
        if (!root)
            return -1;
        for (r = 0; r < s->nslabs; ++r) {
            snprintf(blocks[r], sizeof blocks[r], "result%04d.slab%d.silo:/mesh", s->iteration, r);
            ptrs[r] = blocks[r];
        }
        rc = spdb_put_multimesh(root, "mesh", s->nslabs, ptrs);
        for (v = 0; v < s->nvars && rc == 0; ++v) {
            for (r = 0; r < s->nslabs; ++r)
                snprintf(blocks[r], sizeof blocks[r], "result%04d.slab%d.silo:/%s", s->iteration, r, s->names[v]);
            rc = spdb_put_multivar(root, s->names[v], s->nslabs, ptrs);
        }
        if (rc == 0)
            rc = put_vecs(root);
        if (spdb_close(root) != 0)
            rc = -1;
    }
    return rc;
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
    static silo_sink silos[64];
    fdtd_dump_sink sinks[64];
    const char *sink_env = getenv("FDTD_B200_SINK");
    const int raw_sink = sink_env && !strcmp(sink_env, "raw");
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
    memset(silos, 0, sizeof silos);
    for (r = 0; r < ngpus; ++r) {
        files[r].slab = r;
        files[r].nslabs = ngpus;
        files[r].dx = params.spatial_step;
        silos[r].slab = r;
        silos[r].nslabs = ngpus;
        silos[r].dx = params.spatial_step;
        sinks[r].user = raw_sink ? (void *)&files[r] : (void *)&silos[r];
        sinks[r].begin = raw_sink ? sink_begin : silo_begin;
        sinks[r].variable = raw_sink ? sink_variable : silo_variable;
        sinks[r].end = raw_sink ? sink_end : silo_end;
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
