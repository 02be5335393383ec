/*
 * fdtd_ctx.hpp -- the device context and what the translation units of libfdtd_b200.so share:
 * fdtd_ctx.cu (context, transfers, launches, halo exchange, stepping), fdtd_dump.cu (dump variables,
 * fdtd_propagate), fdtd_diag.cu (energy, analytic-mode error, test pattern, checksum).
 */
#pragma once

#include "fdtd_internal.h"
#include "fdtd_types.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace fdtd;

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fdtd_set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return FDTD_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define NCCL_TRY(expr)                                                                         \
    do {                                                                                       \
        ncclResult_t r_ = (expr);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            fdtd_set_error("%s: %s (%s:%d)", #expr, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
            return FDTD_E_NCCL;                                                                \
        }                                                                                      \
    } while (0)

#define FDTD_TRY(expr)                                                                         \
    do {                                                                                       \
        int rc_ = (expr);                                                                      \
        if (rc_ != FDTD_OK)                                                                    \
            return rc_;                                                                        \
    } while (0)

namespace fdtdi {

/* NCCL is bound at run time, on first multi-GPU use, instead of at link time: a process that has
 * already loaded a libnccl.so.2 (e.g. the one PyTorch ships) keeps using that very library, and a
 * single-GPU run never loads NCCL at all. */struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};
extern NcclApi g_nccl;
int nccl_bind();

constexpr int kSrcRing = 512; /* source rows (steps) resident on the device at a time */

/* Every array is allocated as a ring of nk + 4 + kRollGap plane slots: local planes -1 .. nk + 2 sit in
 * slots 0 .. nk + 3, the gap behind them is what the in-place ("rolling") form of the two-step kernel
 * shifts the state into, one chunk at a time (fdtd_ctx.cu, launch_step2_rolling). */
constexpr int kRollGap = 40;

/* How the halo planes travel between neighbouring slabs (fdtd_halo.cu):
 *   TR_NCCL   ncclSend / ncclRecv on the halo stream (one process per GPU, or a group);
 *   TR_EVENT  slabs of one fdtd_group: the receiver pulls the planes with a peer copy (copy engines, no
 *             SM), ordered by CUDA events of the neighbour's streams -- also works when several slabs
 *             share one device (tests on a 1-GPU box);
 *   TR_FLAG   one process per GPU without NCCL: the sender pushes the planes into the neighbour's halo
 *             planes, mapped through CUDA IPC, and raises a sequence flag in the neighbour's memory;
 *             the streams wait on the flags (cuStreamWaitValue32). */
enum Transport { TR_NONE = 0, TR_NCCL = 1, TR_EVENT = 2, TR_FLAG = 3 };

/* flag words of TR_FLAG, 16 words apart */
enum { SIG_UP_DATA = 0, SIG_DOWN_DATA = 16, SIG_UP_ACK = 32, SIG_DOWN_ACK = 48, SIG_SCRATCH = 64, SIG_WORDS = 96 };

/* which planes an exchange moves */
struct Xchg {
    bool h;        /* Hx, Hy of my top cell plane -> plane 0 of rank+1 */
    bool h_with_e; /*   ... and Ex, Ey, Ez of that plane (fused step) */
    bool e;        /* Ex, Ey of my first node plane -> plane nk+1 of rank-1 */
    bool e_with_hz;/*   ... and Hz (dump variables only) */
    bool wide;     /* instead: two planes of all six arrays each way (two-step kernel); needs h and e */
};

struct DumpPipe;

} /* namespace fdtdi */

struct fdtd_ctx {
    fdtd_params p;
    int device, rank, nranks;
    size_t k0, k1;
    Geo g;
    Fld f;
    double *base;       /* the six arrays of the current state, inside raw (guard margins either side) */
    double *base2;      /* second set for the fused single-sweep step ("kernel" = 2), allocated on demand */
    double *raw, *raw2; /* what cudaMalloc returned */
    Fld f2;
    size_t array_elems; /* stride between the six arrays: P * R * (planes + 2), see create_impl */
    double ch, ce;      /* update factors, main.c:441 / :479 */

    cudaStream_t s_main, s_comm, s_dump;
    cudaEvent_t ev_bnd;             /* compute stream: the planes my neighbours need are final, my halos are consumed */
    cudaEvent_t ev_hhalo, ev_ehalo; /* halo stream: the last exchange of that kind has finished here */
    cudaEvent_t ev_sent;            /* TR_FLAG: my pushes have left my planes */
    bool e_halo_valid, h_halo_valid;
    bool low_e_halo_valid; /* fused step only: plane 0 also holds the lower neighbour's Ex, Ey, Ez */
    bool wide_halo_valid;  /* two-step kernel: planes -1, 0, nk+1, nk+2 of all six arrays are current */
    ncclComm_t comm;
    bool has_comm;
    bool in_group;      /* slab of an fdtd_group: calls that exchange halos go through the group */
    int transport;      /* fdtdi::Transport */
    fdtd_ctx *nb_lo, *nb_hi;         /* TR_EVENT: the neighbouring slabs of the group */
    double *peer_lo[2], *peer_hi[2]; /* TR_FLAG: the neighbours' two state sets; [flip] is their current one */
    unsigned *peer_sig_lo, *peer_sig_hi;
    int peer_nk_lo;                  /* cell planes of the lower neighbour (its upper halo is plane nk + 1) */
    size_t peer_elems_lo, peer_elems_hi; /* elements per array over there (slabs may differ by one plane) */
    void *ipc_mapped[6];
    int n_ipc;
    unsigned *sig;                   /* my flag words (device), SIG_WORDS of them */
    unsigned n_xh, n_xe;             /* exchanges of each kind so far: the sequence numbers of TR_FLAG */
    int flip;                        /* buffer swaps since the slabs were wired, modulo 2 */
    bool wired;                      /* comm_init / peer_connect / group wiring done */
    bool fused_ok;                   /* every slab holds the second state copy (agreed when wiring) */
    int fallback;                    /* 1: the automatically chosen fused kernel was replaced by the split ones */

    /* source */
    fdtd_source_plan plan;
    int src_n;          /* points per row */
    bool src_here;      /* computation mode and this slab holds k = 0 */
    bool src_staged;    /* a kernel of this slab touches the global plane k = 0 (it owns it, or recomputes it as halo) */
    int src_plane;      /* local index of that plane: 1, 0 or -1 */
    double *src_dev;    /* kSrcRing rows of 2*src_n doubles */
    double *src_host;   /* pinned mirror */
    cudaEvent_t ev_src; /* last upload of the ring finished */
    double *src_one_dev; /* single row for the operator-level fdtd_set_source */

    /* options */
    long opt_kernel, opt_strip, opt_kchunk, opt_wx, opt_wy, opt_prefetch, opt_stages, opt_band, opt_l2promo, opt_cluster_x, opt_cluster_y, opt_persistent, opt_window;
    long opt_rolling;       /* 1: the two-step kernel works in place on the rolling window even if a second set fits */
    bool rolling;           /* ... it does (chosen, or the second set did not fit) */
    bool roll_announced;
    int roll_rot;           /* ring slot that holds local plane -1 right now (0 = canonical) */
    int roll_peer_rot_lo, roll_peer_rot_hi, roll_peer_z_lo, roll_peer_z_hi; /* TR_FLAG: the neighbours' rings */
    double *roll_tmp;       /* one plane, for rotating a ring back */
    unsigned *progress_dev; /* planes-completed counters of the persistent two-step kernel */
    size_t progress_elems;
    long opt_host_chunk, opt_host_pipeline; /* fdtd_run_hosted: planes per z-chunk (0 = automatic), 0/1 */
    cudaStream_t s_h2d;                     /* uploads of fdtd_run_hosted (created on first use) */

    /* tensor maps of the TMA-staged fused step: [buffer set][array], valid for tma_bx x tma_by tiles */
    TmaMaps tma_maps[2];
    double *tma_base[2];
    int tma_bx, tma_by, tma_promo;
    int tma_mode;       /* 0: planes 0 .. nk+1; 1: with the spare planes (-1 .. nk+2); 2: the whole ring of slots */
    const void *smem_optin[16]; /* kernels already opted in to large dynamic shared memory on this device */
    int n_smem_optin;
    mutable long launches; /* kernels of this library launched so far (reports) */
    int launch_error;      /* first failure inside a launch helper, reported by queue_step */
    bool kernel_auto;      /* "kernel" was not chosen by the caller: may fall back to the split kernels */

    /* dump scratch for fdtd_aggregate */
    double *agg_dev;
    size_t agg_elems;

    /* asynchronous dump pipeline (fdtd_propagate) */
    fdtdi::DumpPipe *pipe;
};

namespace fdtdi {

int check_ctx(const fdtd_ctx *c, const char *who);
/* check_ctx + refuse slabs of a group: a call that exchanges halos must be issued for all slabs
 * inside one NCCL group, which only the fdtd_group_* entry points do */
int check_solo(const fdtd_ctx *c, const char *who);
int use_device(const fdtd_ctx *c);
double *field_ptr(const fdtd_ctx *c, int idx);

/* dense host shape of each array: row length, rows, planes (main.c:379-407); node = has K+1 planes */
struct DenseShape {
    size_t w, h, d;
    bool node_planes;
};
DenseShape dense_shape(const fdtd_params &p, int idx);

int create_impl(const fdtd_params *p, int device, int rank, int nranks, fdtd_ctx **out);
int copy_planes(fdtd_ctx *c, int idx, double *host_first_owned_plane, int kl0, int kl1, bool to_device, cudaStream_t st);
int copy_field(fdtd_ctx *c, int idx, double *host_first_owned_plane, bool to_device);
void launch_fused(fdtd_ctx *c, const fdtd::Src &s, int kl_begin, int kl_end, cudaStream_t st);
int launch_step2(fdtd_ctx *c, const fdtd::Src &s1, const fdtd::Src &s2, int kl_begin, int kl_end, cudaStream_t st);
void swap_buffers(fdtd_ctx *c);
int settle_kernel(fdtd_ctx *c);
int roll_canonicalise(fdtd_ctx *c);
int ring_slots(const fdtd_ctx *c);
int ensure_pong(fdtd_ctx *c);
void fall_back_to_split(fdtd_ctx *c);
fdtd::Src make_src(const fdtd_ctx *c, const double *row);
int stage_source_rows(fdtd_ctx *c, size_t count, double *t_io);

/* halo exchange (fdtd_halo.cu).  exchange_many moves the planes `x` names between the n slabs this
 * thread drives (n = 1: one process per GPU); on_comm: on the halo streams, after each slab's ev_bnd
 * (stepping) -- otherwise on the compute streams, in order with everything queued so far. */
int exchange_many(fdtd_ctx *const *cs, int n, const Xchg &x, bool on_comm);
/* the compute stream waits until the halos of the last exchanges have arrived and the planes they
 * sent have been read: before any kernel that reads halo planes or overwrites owned ones */
int wait_halos(fdtd_ctx *c);
/* bring the halos the selected kernels need up to date (after uploads and operator-level calls) */
int refresh_halos_many(fdtd_ctx *const *cs, int n, bool fused, bool wide = false);
int halo_side_wait(fdtd_ctx *c, bool top, unsigned seq);
int halo_side_push(fdtd_ctx *c, bool top, unsigned consumed, unsigned seq, double *src_base, int dst_parity);
void halo_destroy(fdtd_ctx *c);
int alloc_sig(fdtd_ctx *c);

/* a step is made of segments: one for the fused kernels, two for the split ones (see fdtd_ctx.cu) */
enum Segment { SEG_FUSED, SEG_H, SEG_E, SEG_STEP2 };
int seg_launch(fdtd_ctx *c, const fdtd::Src &s, Segment seg, const fdtd::Src *second = nullptr);
/* the two-step kernel can serve this cavity's slabs (every slab at least two planes thick) */
bool step2_usable(const fdtd_ctx *c);
Xchg seg_xchg(Segment seg);

/* one context, optionally timed (one process per GPU) */
int run_impl(fdtd_ctx *c, size_t steps, double *time_counter, float *total_ms, float *h_ms, float *e_ms);
/* n contexts driven by this thread: n == 1 -> run_impl, n > 1 -> all slabs of a group (fdtd_group.cu) */
int step_many(fdtd_ctx *const *cs, int n, size_t steps, double *time_counter);
int exchange_many_for_dump(fdtd_ctx *const *cs, int n);

int aggregate_many(fdtd_ctx *const *cs, int n, int var, double *const *host_out);
int energy_many(fdtd_ctx *const *cs, int n, int as_coded, double *e_energy, double *h_energy);

/* dump pipeline (fdtd_dump.cu) */
void pipe_destroy(fdtd_ctx *c);
int propagate_many(fdtd_ctx *const *cs, int n, const fdtd_dump_sink *sinks, size_t *steps_done, double *time_counter);

} /* namespace fdtdi */
