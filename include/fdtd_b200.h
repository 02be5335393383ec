/*
 * fdtd_b200.h -- C ABI of the B200-native FDTD hot path (libfdtd_b200.so).
 *
 * Drop-in boundary for the time-stepping path of Ethalides33/FDTD-Maxwell-microwave-oven:
 * the leapfrog H update, the E update with its implicit PEC walls, and the waveguide source,
 * i.e. the body of the loop at main.c:765-779 of the reference, plus what has to sit either
 * side of it (parameter parsing, field upload/download in the reference's dense layout, the
 * zone-centred dump variables).  Each entry point cites the reference interface it replaces
 * (file:line in the reference tree).  INTEGRATION.md shows the host-side binding.
 *
 * Conventions: plain C types only; every function returns 0 on success and a negative
 * FDTD_E_* code on failure, after which fdtd_last_error() describes the failure (the host keeps
 * the reference's convention of treating any failure as fatal, main.c:154-159).  One host
 * control thread per context.  There is no CPU fallback: without a usable CUDA device every
 * device entry point fails with FDTD_E_CUDA.
 */
#ifndef FDTD_B200_H
#define FDTD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDTD_B200_ABI_VERSION 1

enum {
    FDTD_OK = 0,
    FDTD_E_ARG = -1,     /* bad argument / unsupported geometry */
    FDTD_E_IO = -2,      /* parameter file could not be opened */
    FDTD_E_CUDA = -3,    /* CUDA runtime error (including: no device) */
    FDTD_E_NCCL = -4,    /* NCCL error */
    FDTD_E_NOMEM = -5,   /* host or device allocation failed */
    FDTD_E_STATE = -6    /* call not valid in the context's current state */
};

/* Replaces `Parameters`, main.c:57-71.  The three sizes and the time limit are single precision
 * exactly as in the reference: they are promoted to double at every use, which decides the grid
 * (0.04/0.001 -> 39 cells), the source phase and the number of steps. */
typedef struct fdtd_params {
    float length;           /* x, params.txt number 1 (main.c:226) */
    float width;            /* y, number 2 (main.c:227) */
    float height;           /* z, number 3 (main.c:228) */
    double spatial_step;    /* number 4 (main.c:229) */
    double time_step;       /* number 5 (main.c:230) */
    float simulation_time;  /* number 6 (main.c:231) */
    unsigned sampling_rate; /* number 7 (main.c:232) */
    int mode;               /* number 8 (main.c:233): 0 validation, 1 computation (main.c:37-41) */
    size_t maxi, maxj, maxk;/* derived, main.c:237-239 */
} fdtd_params;

/* Replaces `Fields`, main.c:93-103: six HOST arrays in the reference's dense layout, x fastest
 * (index helpers main.c:379-407).  Sizes: fdtd_field_sizes(). */
typedef struct fdtd_fields {
    double *Ex, *Ey, *Ez, *Hx, *Hy, *Hz;
} fdtd_fields;

/* Source patch on the k = 0 plane (main.c:720-739): i in [i0,i1), j in [j0,j1); the amplitude
 * depends on i - i0 only.  n = i1 - i0. */
typedef struct fdtd_source_plan {
    long i0, i1, j0, j1;
    double z_te;
} fdtd_source_plan;

typedef struct fdtd_ctx fdtd_ctx; /* opaque: device arrays, streams, NCCL communicator */

/* ---- host-side helpers (no device needed) ------------------------------------------------ */

int fdtd_abi_version(void);
const char *fdtd_last_error(void);

/* load_parameters(), main.c:216-242 (same scanf conversions, same grid derivation). */
int fdtd_load_parameters(const char *path, fdtd_params *out);
/* Same derivation from already parsed numbers. */
int fdtd_make_params(float length, float width, float height, double spatial_step,
                     double time_step, float simulation_time, unsigned sampling_rate, int mode,
                     fdtd_params *out);
/* Element counts of Ex,Ey,Ez,Hx,Hy,Hz (main.c:299,310,321,332,343,354). */
int fdtd_field_sizes(const fdtd_params *p, size_t out[6]);
/* Number of passes of the loop at main.c:765 (double counter, float bound, `<=`). */
int fdtd_step_count(const fdtd_params *p, size_t *out);
/* Patch bounds and wave impedance of set_source(), main.c:720-739. */
int fdtd_source_plan_make(const fdtd_params *p, fdtd_source_plan *out);
/* The values set_source() writes at time t (main.c:748,751), computed on the host with the C
 * library's sin: ez_vals[s] and hx_vals[s] for s in [0, i1-i0). */
int fdtd_source_values(const fdtd_params *p, const fdtd_source_plan *plan, double t,
                       double *ez_vals, double *hx_vals);
/* set_initial_conditions(), main.c:416-424, into a host Ey array (host sin, like the reference). */
int fdtd_initial_conditions_host(const fdtd_params *p, double *Ey);
/* z-slab owned by `rank` of `nranks`: cell planes [k0,k1); the first maxk % nranks ranks get one
 * plane more (the decomposition the reference's report describes, description.pdf 2.2). */
int fdtd_slab_range(size_t maxk, int rank, int nranks, size_t *k0, size_t *k1);

/* ---- device context ----------------------------------------------------------------------- */

/* initialize_fields(), main.c:294-364: allocates the six arrays in HBM, zero-filled.
 * Whole cavity on one GPU. */
int fdtd_ctx_create(const fdtd_params *p, int device, fdtd_ctx **out);
/* Same for one z-slab of a multi-GPU run (one process per GPU).  Halo exchange needs
 * fdtd_ctx_comm_init() unless nranks == 1. */
int fdtd_ctx_create_slab(const fdtd_params *p, int device, int rank, int nranks, fdtd_ctx **out);
int fdtd_ctx_destroy(fdtd_ctx *ctx);

/* Wiring the slabs of a one-process-per-GPU run to their neighbours.  Two ways, pick one, on every rank:
 *
 * (1) NCCL: rank 0 makes a 128-byte id, the host program distributes it (MPI, torchrun's store, a
 *     file ...), every rank then joins; halo planes travel with ncclSend / ncclRecv.
 * (2) Peer memory (one box): every rank exports a descriptor of its state (CUDA IPC handles), the host
 *     program gathers the descriptors of ALL ranks, in rank order, and hands the table to every rank;
 *     halo planes are then written straight into the neighbour's HBM over NVLink by the copy engines,
 *     ordered by sequence flags the streams wait on -- no NCCL, no proxy kernels on the SMs.  Ranks may
 *     share a device (that is how the slab logic is tested on a 1-GPU box).
 *
 * Both also agree, across all ranks, on whether the fused kernels can be used (they need the state
 * twice in HBM on every slab): select the "kernel" option BEFORE wiring. */
int fdtd_nccl_unique_id(void *id128);
int fdtd_ctx_comm_init(fdtd_ctx *ctx, const void *id128);
#define FDTD_PEER_BLOB_BYTES 384
int fdtd_ctx_peer_export(fdtd_ctx *ctx, void *blob /* FDTD_PEER_BLOB_BYTES */);
int fdtd_ctx_peer_connect(fdtd_ctx *ctx, const void *blobs /* nranks x FDTD_PEER_BLOB_BYTES, rank order */);

/* Tunables.
 *   "kernel"   0 = one thread per cell (the plain operators, source as separate launches);
 *              1 = z-marching register strips, one H and one E launch per step, in place;
 *              2 = H and E fused into one sweep per step (needs the state twice in HBM);
 *              3 = the same sweep with operands staged by TMA into a shared-memory ring;
 *              4 = TWO time steps per sweep over the TMA ring (default); an odd step and
 *                  slabs thinner than two planes take the single-step sweep of kernel 3.
 *              Kernels 2-4 keep the state twice in HBM.  When the second copy does not fit, kernel 4
 *              works in place on a rolling window of plane slots instead ("rolling" reads 1; 1.04 x
 *              the state, about two thirds of the speed) -- with a line on stderr.  Only
 *              where that is impossible (kernels 2, 3; slabs thinner than two planes) a context left at
 *              its defaults falls back to kernel 1 ("fallback" = 1) and a kernel chosen explicitly fails
 *              with FDTD_E_NOMEM.  On slabs, choose the kernel before wiring.
 *   "rolling"  1 = take the rolling-window form of kernel 4 even if a second copy would fit.
 *   "kchunk"   planes per block;  "stages" depth of the TMA ring (kernels 3, 4);
 *   "warps_y"  kernel 4: block = 32 x warps_y threads storing 28 x (2 warps_y - 3) sites (8, 12, 16);
 *   "strip", "warps_x", "warps_y" kernels 1-3: rows per thread (1..4) and block shape in warps (at most 8
 *              warps; 16 for kernel 3);  "prefetch" planes of L2 prefetch ahead of the sweep (kernel 2);
 *   "band", "l2promo", "cluster_x", "cluster_y"  measurement knobs of kernel 4 (tile numbering, TMA L2
 *              promotion, thread-block clusters of neighbouring tiles);
 *   "persistent" 1 = kernel 4 as one cooperative launch of as many blocks as the GPU holds, a producer warp
 *              per block and flow control between the blocks ("window" planes); reads each array ~once,
 *              bit-identical, not faster than the chunked form on a B200; default 0;
 *   "host_chunk", "host_pipeline"  fdtd_run_hosted (below).
 * Read-only: "k0", "k1" (owned cell planes), "launches" (kernels launched by this library so far),
 *   "fallback" (1 after the automatic fallback), "fused_ok" (second state copy available on every slab),
 *   "transport" (0 none, 1 NCCL, 2 peer copies of a group, 3 peer memory with flags). */
int fdtd_ctx_set_option(fdtd_ctx *ctx, const char *key, long value);
int fdtd_ctx_get_option(fdtd_ctx *ctx, const char *key, long *value);

/* Host <-> HBM, reference dense layout on the host side.  The pointers address the arrays of
 * the WHOLE cavity; a slab context copies only the planes it owns. */
int fdtd_upload(fdtd_ctx *ctx, const fdtd_fields *host);
int fdtd_download(fdtd_ctx *ctx, const fdtd_fields *host);
/* Same, but the pointers address dense arrays holding ONLY this slab's planes: cell planes
 * [k0,k1) for Ez, Hx, Hy; node planes [k0,k1) -- plus plane maxk on the last slab -- for Ex, Ey, Hz.
 * (No host has room for a whole 2048^3 cavity.) */
int fdtd_upload_slab(fdtd_ctx *ctx, const fdtd_fields *host);
int fdtd_download_slab(fdtd_ctx *ctx, const fdtd_fields *host);
/* set_initial_conditions(), main.c:416-424, applied to the device Ey. */
int fdtd_set_initial_conditions(fdtd_ctx *ctx);

/* The three operators exactly as the reference exposes them (in place, one call each): */
int fdtd_set_source(fdtd_ctx *ctx, double time_counter); /* set_source(),     main.c:712-753 */
int fdtd_update_H_field(fdtd_ctx *ctx);                  /* update_H_field(), main.c:431-462 */
int fdtd_update_E_field(fdtd_ctx *ctx);                  /* update_E_field(), main.c:469-500 */

/* `steps` passes of the loop body main.c:770-779 (source, H, source, E) with the source and the
 * PEC walls fused into the update kernels (by default two passes per kernel launch).  *time_counter is advanced by repeated addition
 * of time_step, as at main.c:765.  Asynchronous: returns once the work is queued. */
int fdtd_run(fdtd_ctx *ctx, size_t steps, double *time_counter);
/* Same, bracketed by CUDA events on the context's stream; blocks until done.
 * total_ms: the whole loop.  h_ms / e_ms: summed durations of the H / E update launches
 * (may be NULL). */
int fdtd_run_timed(fdtd_ctx *ctx, size_t steps, double *time_counter, float *total_ms,
                   float *h_ms, float *e_ms);
int fdtd_sync(fdtd_ctx *ctx);
/* The same loop for a caller whose fields live in HOST memory, as the reference's do (its update
 * functions mutate the host arrays in place, main.c:431-500): `host` holds this slab's planes in the
 * reference's dense layout (as for fdtd_upload_slab) and is advanced in place by `steps` passes.
 * On a single-slab context with the fused kernels the upload, the stepping and the download run as one
 * wavefront over z-chunks (chunk c of step s only needs chunks c-1, c, c+1 of step s-1), so the two
 * PCIe directions and the kernels overlap.  Slabs wired through peer memory do the same for runs of up
 * to 64 steps, with neighbouring slabs sweeping in opposite directions so that their wavefronts mesh
 * (call it on every rank).  Results are bit-identical to fdtd_upload_slab + fdtd_run +
 * fdtd_download_slab, which is also what other contexts do.  Use pinned arrays (fdtd_host_alloc).
 * Options: "host_chunk" planes per z-chunk (0 = automatic), "host_pipeline" 0 = always in sequence.
 * Blocks until the host arrays hold the result. */
int fdtd_run_hosted(fdtd_ctx *ctx, const fdtd_fields *host, size_t steps, double *time_counter);

/* Dump variables of write_silo(), main.c:563-579: zone-centred averages with the reference's
 * operand order.  var: 0 ex, 1 ey, 2 ez, 3 hx, 4 hy, 5 hz.  host_out: maxi*maxj*nk doubles for the
 * nk cell planes this context owns (the whole cavity on one GPU).  Blocking. */
int fdtd_aggregate(fdtd_ctx *ctx, int var, double *host_out);

/* Where the dump variables go: the three stages of write_silo(), main.c:550-598.  Called from the
 * context's writer thread, one dump at a time, in order:
 *   begin     <-> DBCreate + DBPutQuadmesh (main.c:553-561); file name is r/result%04d.silo
 *   variable  <-> DBPutQuadvar1 (main.c:564-588), once per variable in the reference's order:
 *                 "ex","ey","ez","hx","hy","hz" and, in validation mode, "aEy","aHx","aHz";
 *                 data is pinned host memory valid only during the call
 *   end       <-> DBPutDefvars + DBClose (main.c:591-597)
 * A non-zero return aborts the run (the reference treats output failures as fatal, main.c:556-559). */
typedef struct fdtd_dump_sink {
    void *user;
    int (*begin)(void *user, int iteration, const size_t dims[3] /* maxi, maxj, planes */,
                 size_t k0 /* first cell plane of this slab */);
    int (*variable)(void *user, const char *name, const double *data, size_t count);
    int (*end)(void *user);
} fdtd_dump_sink;

/* propagate_fields(), main.c:755-799: the initial dump as iteration 1, then the stepping loop
 * with a dump whenever iteration % sampling_rate == 0.  The dump variables are aggregated on the
 * compute stream into HBM scratch (about one step's worth of time), then copied to pinned host
 * buffers on a side stream and handed to `sink` from a writer thread while stepping continues.
 * sink may be NULL (no dumps).  steps_done / time_counter (may be NULL) receive the number of
 * passes and the final time counter.  Blocks until the last dump has been delivered. */
int fdtd_propagate(fdtd_ctx *ctx, const fdtd_dump_sink *sink, size_t *steps_done, double *time_counter);

/* Pinned host memory for the arrays passed to fdtd_upload / fdtd_download (plain malloc'ed
 * arrays work too, at pageable-copy speed). */
int fdtd_host_alloc(size_t bytes, void **out);
/* Same, with the pages taken from the NUMA node the GPU `device` hangs on (when the box tells:
 * /sys/bus/pci/devices/<id>/numa_node) -- on a two-socket box every rank should stream its slab through
 * memory of its own socket.  FDTD_B200_HOST_NUMA=interleave spreads the pages over all nodes instead,
 * =off keeps the default policy.  fdtd_host_numa_info reports what the box tells. */
int fdtd_host_alloc_near(int device, size_t bytes, void **out);
int fdtd_host_numa_info(int device, int *nodes, int *node_of_device);
int fdtd_host_free(void *ptr);

/* ---- all slabs of a cavity from ONE host thread ------------------------------------------------
 * The reference is a single-threaded C program (main.c:807-853); a group keeps its multi-GPU host
 * program one: n z-slab contexts, one per GPU of the box, wired with ncclCommInitAll -- no MPI, no
 * launcher.  Every call queues work on all slabs; they advance concurrently.  The per-slab contexts
 * stay accessible (fdtd_group_ctx) for calls that touch one slab only (fdtd_upload_slab,
 * fdtd_download_slab, fdtd_checksum, fdtd_validation_error, fdtd_ctx_info); calls that exchange halos
 * (stepping, fdtd_aggregate, fdtd_energy, fdtd_propagate) must go through the group, because one
 * thread has to queue every slab's transfers together -- on a slab of a group they return
 * FDTD_E_STATE.
 * devices: ngpus CUDA device indices, or NULL for 0 .. ngpus-1. */
typedef struct fdtd_group fdtd_group;
int fdtd_group_create(const fdtd_params *p, int ngpus, const int *devices, fdtd_group **out);
/* transport 0: peer copies over NVLink ordered by CUDA events (the default of fdtd_group_create; slabs
 * may share a device); 1: NCCL send/recv (also selected by FDTD_B200_TRANSPORT=nccl in the environment) */
int fdtd_group_create_transport(const fdtd_params *p, int ngpus, const int *devices, int transport,
                                fdtd_group **out);
int fdtd_group_destroy(fdtd_group *group);
int fdtd_group_size(fdtd_group *group);
int fdtd_group_ctx(fdtd_group *group, int rank, fdtd_ctx **out);
int fdtd_group_set_option(fdtd_group *group, const char *key, long value);
int fdtd_group_upload(fdtd_group *group, const fdtd_fields *whole_cavity);      /* each slab takes its planes */
int fdtd_group_download(fdtd_group *group, const fdtd_fields *whole_cavity);
int fdtd_group_set_initial_conditions(fdtd_group *group);                        /* main.c:416-424 */
int fdtd_group_run(fdtd_group *group, size_t steps, double *time_counter);      /* loop body main.c:770-779 */
int fdtd_group_sync(fdtd_group *group);
int fdtd_group_aggregate(fdtd_group *group, int var, double *host_out /* maxi*maxj*maxk */); /* main.c:511-540 */
int fdtd_group_energy(fdtd_group *group, int as_coded, double *e_energy, double *h_energy); /* main.c:602-668 */
/* propagate_fields(), main.c:755-799; sinks: one per slab (ngpus entries; each is called from its own
 * writer thread with that slab's planes and k0), or NULL for no dumps. */
int fdtd_group_propagate(fdtd_group *group, const fdtd_dump_sink *sinks, size_t *steps_done,
                         double *time_counter);

/* Diagnostics that sit next to the path in the reference (they never feed back into the fields).
 * Both are reductions and agree with the reference's sequential sums to rounding, not bit for bit.
 * A slab context returns its own zones' contribution; add the slabs up for the cavity.
 *
 * fdtd_energy: calculate_E_energy() / calculate_H_energy(), main.c:602-668.  as_coded != 0 keeps the
 *   reference's indexing of Ez with Hz's strides (main.c:627; single-GPU contexts only, the slip reaches
 *   across planes); 0 uses the intended zone average.
 * fdtd_validation_error: against the analytic TE101 fields of update_validation_fields_then_subfdtd(),
 *   main.c:670-710, at time_counter: sums[] = {sum (a-Ey)^2, sum a^2, same for Hx, same for Hz};
 *   rel_l2[] = sqrt(num/den) per field (the report's e_r, description.pdf eq. 2).  Either may be NULL. */
int fdtd_energy(fdtd_ctx *ctx, int as_coded, double *e_energy, double *h_energy);
int fdtd_validation_error(fdtd_ctx *ctx, double time_counter, double sums[6], double rel_l2[3]);

/* Full-size test support (no host copy of a 1024^3 state exists).  Both are pure functions of an
 * element's index in the reference's dense arrays (main.c:379-407), independent of pitch and slabs.
 * fdtd_fill_test_pattern: every element of the six arrays := 2u-1 with u = (splitmix64(seed ^
 *   array<<58 ^ dense_index) >> 11) / 2^53.
 * fdtd_checksum: out[a] = sum over this context's owned elements of
 *   splitmix64(bit pattern + dense_index) mod 2^64; slab checksums add up to the whole cavity's. */
int fdtd_fill_test_pattern(fdtd_ctx *ctx, unsigned long long seed);
int fdtd_checksum(fdtd_ctx *ctx, unsigned long long out[6]);

/* Bytes of HBM held by the context, and its geometry (pitch in doubles, rows per plane,
 * planes) -- for reports. */
int fdtd_ctx_info(fdtd_ctx *ctx, size_t *hbm_bytes, size_t *pitch, size_t *rows, size_t *planes);

#ifdef __cplusplus
}
#endif
#endif
