/*
 * fdtd_halo.cu -- how the halo planes travel between neighbouring z-slabs.
 *
 * What travels is fixed by the stencils (main.c:448-455 read Ex, Ey of plane k+1; main.c:486-493 read
 * Hx, Hy of plane k-1): after an H update the top cell plane of Hx, Hy goes up, after an E update the
 * first node plane of Ex, Ey goes down; the fused step sends both once per step, plus Ex, Ey, Ez of the
 * top plane (the slab above recomputes H_new of the plane below it).  Three transports move them:
 *
 *   TR_NCCL   ncclSend / ncclRecv in one NCCL group (bound with dlopen on first use).
 *   TR_EVENT  the slabs of an fdtd_group live in one process: the receiver PULLS the planes with peer
 *             copies on its own stream (copy engines over NVLink, no SM time, no proxy kernels), after
 *             waiting for the neighbour's "boundary planes final" event.  Works for slabs that share a
 *             device too, so the whole multi-slab logic is testable on a 1-GPU box.
 *   TR_FLAG   one process per GPU: every rank maps its neighbours' state through CUDA IPC
 *             (fdtd_ctx_peer_export / fdtd_ctx_peer_connect) and PUSHES its planes into their halo planes,
 *             then raises a sequence number in the neighbour's memory.  Streams wait on the numbers with
 *             cuStreamWaitValue32 (a one-thread polling kernel if the driver lacks it).  A halo slot is
 *             only overwritten after its owner has acknowledged that the previous content was consumed.
 *
 * Every exchange has the same three phases for the n slabs the calling thread drives, so that one
 * thread can serve a whole group:  A  each slab's stream is made to wait for what it depends on and
 * its "ready" event is recorded;  B  the transfers are queued;  C  the completion events are recorded.
 */
#include "fdtd_ctx.hpp"

#include <unistd.h>

using namespace fdtdi;

namespace fdtdi {

NcclApi g_nccl;

int nccl_bind()
{
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&mu);
    if (!g_nccl.ok) {
        const char *name = getenv("FDTD_B200_NCCL_LIB");
        void *h = name ? dlopen(name, RTLD_NOW | RTLD_GLOBAL) : dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h && !name)
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
#define BIND(field, sym) *(void **)(&g_nccl.field) = dlsym(h, sym)
            BIND(GetUniqueId, "ncclGetUniqueId");
            BIND(CommInitRank, "ncclCommInitRank");
            BIND(CommInitAll, "ncclCommInitAll");
            BIND(CommDestroy, "ncclCommDestroy");
            BIND(Send, "ncclSend");
            BIND(Recv, "ncclRecv");
            BIND(AllReduce, "ncclAllReduce");
            BIND(GroupStart, "ncclGroupStart");
            BIND(GroupEnd, "ncclGroupEnd");
            BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
            g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.Send &&
                        g_nccl.Recv && g_nccl.AllReduce && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
        }
    }
    const bool ok = g_nccl.ok;
    pthread_mutex_unlock(&mu);
    if (!ok) {
        fdtd_set_error("cannot load NCCL (libnccl.so.2): %s", dlerror() ? dlerror() : "symbols missing");
        return FDTD_E_NCCL;
    }
    return FDTD_OK;
}

/* ---- sequence flags (TR_FLAG) --------------------------------------------------------------- */

__global__ void k_post_flag(volatile unsigned *flag, unsigned value)
{
    __threadfence_system();
    *flag = value;
    __threadfence_system();
}

/* polling fallback for drivers without stream memory operations; the comparison is cyclic */
__global__ void k_wait_flag(const volatile unsigned *flag, unsigned value)
{
    while ((int)(*flag - value) < 0)
        __nanosleep(200);
    __threadfence_system();
}

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static WaitValue32Fn wait_value_fn()
{
    static bool looked = false;
    static WaitValue32Fn fn = nullptr;
    if (!looked) {
        looked = true;
        const char *off = getenv("FDTD_B200_NO_STREAM_MEMOPS");
        if (!(off && off[0] == '1')) {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && p &&
                q == cudaDriverEntryPointSuccess)
                fn = (WaitValue32Fn)p;
            cudaGetLastError();
        }
    }
    return fn;
}

static int flag_post(fdtd_ctx *c, cudaStream_t st, unsigned *flag, unsigned value)
{
    k_post_flag<<<1, 1, 0, st>>>(flag, value);
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
    return FDTD_OK;
}

static int flag_wait(fdtd_ctx *c, cudaStream_t st, unsigned *flag, unsigned value)
{
    if (WaitValue32Fn fn = wait_value_fn()) {
        const CUresult r = fn((CUstream)st, (CUdeviceptr)(uintptr_t)flag, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (r == CUDA_SUCCESS)
            return FDTD_OK;
        /* not supported on this device after all: poll from here on */
    }
    k_wait_flag<<<1, 1, 0, st>>>(flag, value);
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
    return FDTD_OK;
}

int alloc_sig(fdtd_ctx *c)
{
    if (c->sig)
        return FDTD_OK;
    CUDA_TRY(cudaMalloc((void **)&c->sig, SIG_WORDS * sizeof(unsigned)));
    CUDA_TRY(cudaMemset(c->sig, 0, SIG_WORDS * sizeof(unsigned)));
    return FDTD_OK;
}

/* ---- what travels ----------------------------------------------------------------------------- */

/* One direction of an exchange: `count` arrays (indices into Ex Ey Ez Hx Hy Hz), `planes` consecutive
 * planes starting at the sender's local plane `src`, landing at the receiver's local plane `dst`
 * (dst may be -1: the spare plane below plane 0). */
struct Lane {
    int count, arrays[6];
    int planes, src, dst;
};

/* upward: from a slab with nk_sender cell planes into the slab above it */
static Lane up_lane(const Xchg &x, int nk_sender)
{
    Lane l;
    if (x.wide) { /* two-step kernel: the top two owned planes of everything -> planes -1, 0 */
        l = Lane{6, {0, 1, 2, 3, 4, 5}, 2, nk_sender - 1, -1};
    } else {      /* Hx, Hy (+ Ex, Ey, Ez for the fused step) of the top owned plane -> plane 0 */
        l = Lane{x.h_with_e ? 5 : 2, {3, 4, 0, 1, 2, 0}, 1, nk_sender, 0};
    }
    return l;
}

/* downward: from a slab into the slab below it, which has nk_receiver cell planes */
static Lane down_lane(const Xchg &x, int nk_receiver)
{
    Lane l;
    if (x.wide) { /* the first two owned planes of everything -> planes nk + 1, nk + 2 */
        l = Lane{6, {0, 1, 2, 3, 4, 5}, 2, 1, nk_receiver + 1};
    } else {      /* Ex, Ey (+ Hz for dumps) of the first owned plane -> plane nk + 1 */
        l = Lane{x.e_with_hz ? 3 : 2, {0, 1, 5, 0, 0, 0}, 1, 1, nk_receiver + 1};
    }
    return l;
}

/* where local plane `plane` of array `array` is; Z > 0: the arrays are rings rotated by `rot` slots
 * (rolling form of the two-step kernel), else plane -1 is simply the spare plane below plane 0 */
static double *plane_ptr(double *base, size_t array_elems, long long PR, int array, int plane, int rot = 0, int Z = 0)
{
    const long long slot = Z ? (plane + 1 + rot) % Z : plane + 1;
    return base + (size_t)array * array_elems + (slot - 1) * PR;
}

static int own_z(const fdtd_ctx *c) { return c->rolling ? ring_slots(c) : 0; }

/* TR_NCCL: the sends and receives of one slab (inside the caller's NCCL group) */
static int nccl_exchange(fdtd_ctx *c, cudaStream_t st, const Xchg &x)
{
    if (!c->has_comm) {
        fdtd_set_error("multi-rank context without communicator: call fdtd_ctx_comm_init first");
        return FDTD_E_STATE;
    }
    const long long PR = c->g.PR;
    const bool up = c->rank + 1 < c->nranks, lo = c->rank > 0;
    if (x.h) {
        const Lane l = up_lane(x, c->g.nk);
        const size_t n = (size_t)PR; /* plane by plane: on a ring two planes need not be neighbours in memory */
        if (up)
            for (int i = 0; i < l.count; ++i)
                for (int q = 0; q < l.planes; ++q)
                    NCCL_TRY(g_nccl.Send(plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.src + q, c->roll_rot, own_z(c)),
                                         n, ncclDouble, c->rank + 1, c->comm, st));
        if (lo)
            for (int i = 0; i < l.count; ++i)
                for (int q = 0; q < l.planes; ++q)
                    NCCL_TRY(g_nccl.Recv(plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.dst + q, c->roll_rot, own_z(c)),
                                         n, ncclDouble, c->rank - 1, c->comm, st));
    }
    if (x.e) {
        const Lane l = down_lane(x, c->g.nk); /* as receiver: my own nk */
        const size_t n = (size_t)PR;
        if (lo)
            for (int i = 0; i < l.count; ++i)
                for (int q = 0; q < l.planes; ++q)
                    NCCL_TRY(g_nccl.Send(plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.src + q, c->roll_rot, own_z(c)),
                                         n, ncclDouble, c->rank - 1, c->comm, st));
        if (up)
            for (int i = 0; i < l.count; ++i)
                for (int q = 0; q < l.planes; ++q)
                    NCCL_TRY(g_nccl.Recv(plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.dst + q, c->roll_rot, own_z(c)),
                                         n, ncclDouble, c->rank + 1, c->comm, st));
    }
    return FDTD_OK;
}

/* TR_EVENT: this slab pulls what its neighbours hold for it */
static int event_pull(fdtd_ctx *c, cudaStream_t st, const Xchg &x)
{
    const long long PR = c->g.PR;
    if (x.h && c->nb_lo) {
        fdtd_ctx *s = c->nb_lo;
        CUDA_TRY(cudaStreamWaitEvent(st, s->ev_bnd, 0));
        const Lane l = up_lane(x, s->g.nk);
        const bool ring = c->rolling || s->rolling; /* on a ring two planes need not be neighbours in memory */
        for (int i = 0; i < l.count; ++i)
            for (int q = 0; q < l.planes; q += ring ? 1 : l.planes)
                CUDA_TRY(cudaMemcpyPeerAsync(plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.dst + q, c->roll_rot, own_z(c)),
                                             c->device,
                                             plane_ptr(s->base, s->array_elems, PR, l.arrays[i], l.src + q, s->roll_rot, own_z(s)),
                                             s->device, (size_t)PR * (ring ? 1 : l.planes) * sizeof(double), st));
    }
    if (x.e && c->nb_hi) {
        fdtd_ctx *s = c->nb_hi;
        CUDA_TRY(cudaStreamWaitEvent(st, s->ev_bnd, 0));
        const Lane l = down_lane(x, c->g.nk);
        const bool ring = c->rolling || s->rolling; /* on a ring two planes need not be neighbours in memory */
        for (int i = 0; i < l.count; ++i)
            for (int q = 0; q < l.planes; q += ring ? 1 : l.planes)
                CUDA_TRY(cudaMemcpyPeerAsync(plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.dst + q, c->roll_rot, own_z(c)),
                                             c->device,
                                             plane_ptr(s->base, s->array_elems, PR, l.arrays[i], l.src + q, s->roll_rot, own_z(s)),
                                             s->device, (size_t)PR * (ring ? 1 : l.planes) * sizeof(double), st));
    }
    return FDTD_OK;
}

/* TR_FLAG: this slab pushes into its neighbours' halo planes */
static int flag_push(fdtd_ctx *c, cudaStream_t st, const Xchg &x)
{
    const long long PR = c->g.PR;
    const bool up = c->rank + 1 < c->nranks, lo = c->rank > 0;
    if (x.h) {
        const unsigned seq = ++c->n_xh;
        if (lo) /* everything that read the previous content of my lower halo precedes this point of `st` */
            FDTD_TRY(flag_post(c, st, c->peer_sig_lo + SIG_UP_ACK, seq - 1));
        if (up) {
            FDTD_TRY(flag_wait(c, st, c->sig + SIG_UP_ACK, seq - 1));
            const Lane l = up_lane(x, c->g.nk);
            for (int i = 0; i < l.count; ++i)
                for (int q = 0; q < l.planes; q += c->rolling ? 1 : l.planes)
                    CUDA_TRY(cudaMemcpyAsync(plane_ptr(c->peer_hi[c->flip], c->peer_elems_hi, PR, l.arrays[i], l.dst + q,
                                                       c->roll_peer_rot_hi, c->rolling ? c->roll_peer_z_hi : 0),
                                             plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.src + q, c->roll_rot, own_z(c)),
                                             (size_t)PR * (c->rolling ? 1 : l.planes) * sizeof(double), cudaMemcpyDefault, st));
            FDTD_TRY(flag_post(c, st, c->peer_sig_hi + SIG_UP_DATA, seq));
        }
    }
    if (x.e) {
        const unsigned seq = ++c->n_xe;
        if (up)
            FDTD_TRY(flag_post(c, st, c->peer_sig_hi + SIG_DOWN_ACK, seq - 1));
        if (lo) {
            FDTD_TRY(flag_wait(c, st, c->sig + SIG_DOWN_ACK, seq - 1));
            const Lane l = down_lane(x, c->peer_nk_lo);
            for (int i = 0; i < l.count; ++i)
                for (int q = 0; q < l.planes; q += c->rolling ? 1 : l.planes)
                    CUDA_TRY(cudaMemcpyAsync(plane_ptr(c->peer_lo[c->flip], c->peer_elems_lo, PR, l.arrays[i], l.dst + q,
                                                       c->roll_peer_rot_lo, c->rolling ? c->roll_peer_z_lo : 0),
                                             plane_ptr(c->base, c->array_elems, PR, l.arrays[i], l.src + q, c->roll_rot, own_z(c)),
                                             (size_t)PR * (c->rolling ? 1 : l.planes) * sizeof(double), cudaMemcpyDefault, st));
            FDTD_TRY(flag_post(c, st, c->peer_sig_lo + SIG_DOWN_DATA, seq));
        }
    }
    return FDTD_OK;
}

/* ---- one side of a slab at a time (the hosted wavefront over several slabs, fdtd_hosted.cu) ----------
 * TR_FLAG only, everything on the compute stream.  `top`: the interface with the slab above (else below).
 * Sequence numbers continue the lanes' own (n_xh for the upward lane, n_xe for the downward one). */

/* the compute stream waits until the neighbour on that side has delivered its planes number `seq` */
int halo_side_wait(fdtd_ctx *c, bool top, unsigned seq)
{
    return flag_wait(c, c->s_main, c->sig + (top ? SIG_DOWN_DATA : SIG_UP_DATA), seq);
}

/* tell that neighbour that its planes number `consumed` have been used, then send it my two boundary planes
 * of the state in `src_base` (all six arrays) as number `seq`, into its buffer set `dst_parity` */
int halo_side_push(fdtd_ctx *c, bool top, unsigned consumed, unsigned seq, double *src_base, int dst_parity)
{
    const long long PR = c->g.PR;
    Xchg x{};
    x.h = x.e = x.wide = true;
    cudaStream_t st = c->s_main;
    if (top) {
        FDTD_TRY(flag_post(c, st, c->peer_sig_hi + SIG_DOWN_ACK, consumed));
        FDTD_TRY(flag_wait(c, st, c->sig + SIG_UP_ACK, seq - 2));
        const Lane l = up_lane(x, c->g.nk);
        for (int i = 0; i < l.count; ++i)
            CUDA_TRY(cudaMemcpyAsync(plane_ptr(c->peer_hi[dst_parity], c->peer_elems_hi, PR, l.arrays[i], l.dst),
                                     plane_ptr(src_base, c->array_elems, PR, l.arrays[i], l.src),
                                     (size_t)PR * l.planes * sizeof(double), cudaMemcpyDefault, st));
        FDTD_TRY(flag_post(c, st, c->peer_sig_hi + SIG_UP_DATA, seq));
    } else {
        FDTD_TRY(flag_post(c, st, c->peer_sig_lo + SIG_UP_ACK, consumed));
        FDTD_TRY(flag_wait(c, st, c->sig + SIG_DOWN_ACK, seq - 2));
        const Lane l = down_lane(x, c->peer_nk_lo);
        for (int i = 0; i < l.count; ++i)
            CUDA_TRY(cudaMemcpyAsync(plane_ptr(c->peer_lo[dst_parity], c->peer_elems_lo, PR, l.arrays[i], l.dst),
                                     plane_ptr(src_base, c->array_elems, PR, l.arrays[i], l.src),
                                     (size_t)PR * l.planes * sizeof(double), cudaMemcpyDefault, st));
        FDTD_TRY(flag_post(c, st, c->peer_sig_lo + SIG_DOWN_DATA, seq));
    }
    return FDTD_OK;
}

int wait_halos(fdtd_ctx *c)
{
    if (c->nranks == 1)
        return FDTD_OK;
    if (c->transport == TR_FLAG) {
        CUDA_TRY(cudaStreamWaitEvent(c->s_main, c->ev_sent, 0));
        if (c->rank > 0)
            FDTD_TRY(flag_wait(c, c->s_main, c->sig + SIG_UP_DATA, c->n_xh));
        if (c->rank + 1 < c->nranks)
            FDTD_TRY(flag_wait(c, c->s_main, c->sig + SIG_DOWN_DATA, c->n_xe));
        return FDTD_OK;
    }
    /* waiting for an event that was never recorded is a no-op */
    CUDA_TRY(cudaStreamWaitEvent(c->s_main, c->ev_ehalo, 0));
    CUDA_TRY(cudaStreamWaitEvent(c->s_main, c->ev_hhalo, 0));
    if (c->transport == TR_EVENT) {
        /* my neighbours pull from my planes: they must have finished before I overwrite them */
        for (fdtd_ctx *nb : {c->nb_lo, c->nb_hi})
            if (nb) {
                CUDA_TRY(cudaStreamWaitEvent(c->s_main, nb->ev_ehalo, 0));
                CUDA_TRY(cudaStreamWaitEvent(c->s_main, nb->ev_hhalo, 0));
            }
    }
    return FDTD_OK;
}

int exchange_many(fdtd_ctx *const *cs, int n, const Xchg &x, bool on_comm)
{
    if (cs[0]->nranks == 1 || (!x.h && !x.e))
        return FDTD_OK;
    const int tr = cs[0]->transport;
    if (tr == TR_NONE || !cs[0]->wired) {
        fdtd_set_error("multi-rank context is not wired to its neighbours: call fdtd_ctx_comm_init or "
                       "fdtd_ctx_peer_connect first");
        return FDTD_E_STATE;
    }
    /* A: dependencies of the transfers */
    for (int r = 0; r < n; ++r) {
        fdtd_ctx *c = cs[r];
        FDTD_TRY(use_device(c));
        if (on_comm) {
            CUDA_TRY(cudaStreamWaitEvent(c->s_comm, c->ev_bnd, 0));
        } else {
            FDTD_TRY(wait_halos(c));
            if (tr == TR_EVENT)
                CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
        }
    }
    /* B: the transfers */
    int rc = FDTD_OK;
    if (tr == TR_NCCL) {
        NCCL_TRY(g_nccl.GroupStart());
        for (int r = 0; r < n && rc == FDTD_OK; ++r) {
            rc = use_device(cs[r]);
            if (rc == FDTD_OK)
                rc = nccl_exchange(cs[r], on_comm ? cs[r]->s_comm : cs[r]->s_main, x);
        }
        const ncclResult_t e = g_nccl.GroupEnd();
        if (rc == FDTD_OK && e != ncclSuccess) {
            fdtd_set_error("ncclGroupEnd: %s", g_nccl.GetErrorString(e));
            rc = FDTD_E_NCCL;
        }
    } else {
        for (int r = 0; r < n && rc == FDTD_OK; ++r) {
            fdtd_ctx *c = cs[r];
            rc = use_device(c);
            if (rc != FDTD_OK)
                break;
            cudaStream_t st = on_comm ? c->s_comm : c->s_main;
            rc = tr == TR_EVENT ? event_pull(c, st, x) : flag_push(c, st, x);
        }
    }
    FDTD_TRY(rc);
    /* C: completion */
    for (int r = 0; r < n; ++r) {
        fdtd_ctx *c = cs[r];
        FDTD_TRY(use_device(c));
        cudaStream_t st = on_comm ? c->s_comm : c->s_main;
        if (tr == TR_FLAG) {
            CUDA_TRY(cudaEventRecord(c->ev_sent, st));
            if (!on_comm)
                FDTD_TRY(wait_halos(c)); /* what follows on the compute stream reads the new halos */
        } else {
            if (x.h)
                CUDA_TRY(cudaEventRecord(c->ev_hhalo, st));
            if (x.e)
                CUDA_TRY(cudaEventRecord(c->ev_ehalo, st));
        }
    }
    return FDTD_OK;
}

/* The validity flags are set identically on every slab, so the transfers always pair up. */
int refresh_halos_many(fdtd_ctx *const *cs, int n, bool fused, bool wide)
{
    fdtd_ctx *c0 = cs[0];
    if (c0->nranks == 1)
        return FDTD_OK;
    if (wide) { /* the two-step kernel's halos: two planes of everything, both ways */
        if (c0->wide_halo_valid)
            return FDTD_OK;
        Xchg w{};
        w.h = w.e = w.wide = true;
        FDTD_TRY(exchange_many(cs, n, w, false));
        for (int r = 0; r < n; ++r) /* planes 0 and nk + 1 of every array are current as well */
            cs[r]->wide_halo_valid = cs[r]->e_halo_valid = cs[r]->h_halo_valid = cs[r]->low_e_halo_valid = true;
        return FDTD_OK;
    }
    Xchg x{};
    x.e = !c0->e_halo_valid;
    x.h = !c0->h_halo_valid || (fused && !c0->low_e_halo_valid);
    x.h_with_e = fused;
    if (!x.e && !x.h)
        return FDTD_OK;
    FDTD_TRY(exchange_many(cs, n, x, false));
    for (int r = 0; r < n; ++r) {
        if (x.e)
            cs[r]->e_halo_valid = true;
        if (x.h) {
            cs[r]->h_halo_valid = true;
            cs[r]->low_e_halo_valid = fused;
        }
    }
    return FDTD_OK;
}

void halo_destroy(fdtd_ctx *c)
{
    for (int k = 0; k < c->n_ipc; ++k)
        if (c->ipc_mapped[k])
            cudaIpcCloseMemHandle(c->ipc_mapped[k]);
    c->n_ipc = 0;
    if (c->has_comm)
        g_nccl.CommDestroy(c->comm);
    c->has_comm = false;
    if (c->sig)
        cudaFree(c->sig);
    c->sig = nullptr;
}

/* ---- wiring one-process-per-GPU slabs ----------------------------------------------------------- */

struct PeerBlob {
    unsigned magic;
    int rank, nranks, device;
    long long pid;
    unsigned long long host;
    int nk, has_pong;
    unsigned long long array_elems, margin_front;
    unsigned long long raw, raw2, sig; /* addresses in the exporting process */
    cudaIpcMemHandle_t h_raw, h_raw2, h_sig;
};
static_assert(sizeof(PeerBlob) <= FDTD_PEER_BLOB_BYTES, "peer blob must fit the ABI's buffer");
constexpr unsigned kBlobMagic = 0xFD7DB200u;

static unsigned long long host_id()
{
    char name[256] = {0};
    gethostname(name, sizeof name - 1);
    unsigned long long h = 1469598103934665603ull;
    for (const char *p = name; *p; ++p)
        h = (h ^ (unsigned char)*p) * 1099511628211ull;
    if (FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r")) {
        int ch;
        while ((ch = fgetc(f)) != EOF)
            h = (h ^ (unsigned char)ch) * 1099511628211ull;
        fclose(f);
    }
    return h;
}

size_t margin_front_of(const fdtd_ctx *c);

/* the second state copy is a per-slab allocation; whether the fused kernels can be used has to be
 * the same answer on every slab, because their halo plan differs from the split kernels' */
static void settle_fused(fdtd_ctx *c, bool all_have_pong)
{
    c->fused_ok = all_have_pong;
    if (!all_have_pong && c->raw2) {
        cudaFree(c->raw2);
        c->raw2 = c->base2 = nullptr;
    }
}

} /* namespace fdtdi */

extern "C" {

int fdtd_nccl_unique_id(void *id128)
{
    if (!id128) {
        fdtd_set_error("fdtd_nccl_unique_id: NULL argument");
        return FDTD_E_ARG;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    FDTD_TRY(nccl_bind());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return FDTD_OK;
}

int fdtd_ctx_comm_init(fdtd_ctx *c, const void *id128)
{
    FDTD_TRY(check_ctx(c, "fdtd_ctx_comm_init"));
    if (!id128) {
        fdtd_set_error("fdtd_ctx_comm_init: NULL id");
        return FDTD_E_ARG;
    }
    if (c->wired) {
        fdtd_set_error("fdtd_ctx_comm_init: context is already wired to its neighbours");
        return FDTD_E_STATE;
    }
    FDTD_TRY(use_device(c));
    FDTD_TRY(nccl_bind());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&c->comm, c->nranks, id, c->rank));
    c->has_comm = true;
    c->transport = TR_NCCL;
    /* agree on the fused kernels: minimum over the ranks of "I hold the second state copy" */
    int have = 0;
    if (c->opt_kernel >= 2 && !c->opt_rolling) {
        const int rc = ensure_pong(c);
        if (rc != FDTD_OK && rc != FDTD_E_NOMEM)
            return rc;
        have = rc == FDTD_OK;
        cudaGetLastError();
    }
    FDTD_TRY(alloc_sig(c));
    int *scratch = (int *)(c->sig + SIG_SCRATCH);
    CUDA_TRY(cudaMemcpyAsync(scratch, &have, sizeof have, cudaMemcpyHostToDevice, c->s_main));
    NCCL_TRY(g_nccl.AllReduce(scratch, scratch + 1, 1, ncclInt, ncclMin, c->comm, c->s_main));
    int all = 0;
    CUDA_TRY(cudaMemcpyAsync(&all, scratch + 1, sizeof all, cudaMemcpyDeviceToHost, c->s_main));
    CUDA_TRY(cudaStreamSynchronize(c->s_main));
    settle_fused(c, all != 0);
    c->wired = true;
    return FDTD_OK;
}

int fdtd_ctx_peer_export(fdtd_ctx *c, void *blob)
{
    FDTD_TRY(check_ctx(c, "fdtd_ctx_peer_export"));
    if (!blob) {
        fdtd_set_error("fdtd_ctx_peer_export: NULL argument");
        return FDTD_E_ARG;
    }
    if (c->wired || c->in_group) {
        fdtd_set_error("fdtd_ctx_peer_export: context is already wired to its neighbours");
        return FDTD_E_STATE;
    }
    FDTD_TRY(use_device(c));
    if (c->opt_kernel >= 2 && !c->opt_rolling) {
        const int rc = ensure_pong(c);
        if (rc != FDTD_OK && rc != FDTD_E_NOMEM)
            return rc;
        cudaGetLastError();
    }
    FDTD_TRY(alloc_sig(c));
    CUDA_TRY(cudaStreamSynchronize(c->s_main));
    PeerBlob b;
    memset(&b, 0, sizeof b);
    b.magic = kBlobMagic;
    b.rank = c->rank;
    b.nranks = c->nranks;
    b.device = c->device;
    b.pid = (long long)getpid();
    b.host = host_id();
    b.nk = c->g.nk;
    b.has_pong = c->raw2 != nullptr;
    b.array_elems = c->array_elems;
    b.margin_front = margin_front_of(c);
    b.raw = (unsigned long long)(uintptr_t)c->raw;
    b.raw2 = (unsigned long long)(uintptr_t)c->raw2;
    b.sig = (unsigned long long)(uintptr_t)c->sig;
    CUDA_TRY(cudaIpcGetMemHandle(&b.h_raw, c->raw));
    if (c->raw2)
        CUDA_TRY(cudaIpcGetMemHandle(&b.h_raw2, c->raw2));
    CUDA_TRY(cudaIpcGetMemHandle(&b.h_sig, c->sig));
    memset(blob, 0, FDTD_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    c->flip = 0;
    return FDTD_OK;
}

static int map_peer(fdtd_ctx *c, const PeerBlob &b, bool fused, double *sets[2], unsigned **sig)
{
    if (b.host != host_id()) {
        fdtd_set_error("fdtd_ctx_peer_connect: rank %d runs on another host; peer memory needs one box (use NCCL)", b.rank);
        return FDTD_E_STATE;
    }
    sets[0] = sets[1] = nullptr;
    if (b.pid == (long long)getpid()) { /* same process (threads): plain pointers */
        if (b.device != c->device) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->device, b.device));
            if (!can) {
                fdtd_set_error("fdtd_ctx_peer_connect: device %d cannot access device %d", c->device, b.device);
                return FDTD_E_STATE;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                CUDA_TRY(e);
            cudaGetLastError();
        }
        sets[0] = (double *)(uintptr_t)b.raw + b.margin_front;
        if (fused)
            sets[1] = (double *)(uintptr_t)b.raw2 + b.margin_front;
        *sig = (unsigned *)(uintptr_t)b.sig;
        return FDTD_OK;
    }
    void *p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, b.h_raw, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_mapped[c->n_ipc++] = p;
    sets[0] = (double *)p + b.margin_front;
    if (fused) {
        CUDA_TRY(cudaIpcOpenMemHandle(&p, b.h_raw2, cudaIpcMemLazyEnablePeerAccess));
        c->ipc_mapped[c->n_ipc++] = p;
        sets[1] = (double *)p + b.margin_front;
    }
    CUDA_TRY(cudaIpcOpenMemHandle(&p, b.h_sig, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_mapped[c->n_ipc++] = p;
    *sig = (unsigned *)p;
    return FDTD_OK;
}

int fdtd_ctx_peer_connect(fdtd_ctx *c, const void *blobs)
{
    FDTD_TRY(check_ctx(c, "fdtd_ctx_peer_connect"));
    if (!blobs) {
        fdtd_set_error("fdtd_ctx_peer_connect: NULL argument");
        return FDTD_E_ARG;
    }
    if (c->wired || c->in_group) {
        fdtd_set_error("fdtd_ctx_peer_connect: context is already wired to its neighbours");
        return FDTD_E_STATE;
    }
    if (!c->sig) {
        fdtd_set_error("fdtd_ctx_peer_connect: call fdtd_ctx_peer_export first");
        return FDTD_E_STATE;
    }
    FDTD_TRY(use_device(c));
    std::vector<PeerBlob> all(c->nranks);
    bool every_pong = true;
    for (int r = 0; r < c->nranks; ++r) {
        memcpy(&all[r], (const char *)blobs + (size_t)r * FDTD_PEER_BLOB_BYTES, sizeof(PeerBlob));
        const PeerBlob &b = all[r];
        if (b.magic != kBlobMagic || b.rank != r || b.nranks != c->nranks || b.array_elems / (size_t)(b.nk + 4 + kRollGap) !=
                                                                                 c->array_elems / (size_t)(c->g.nk + 4 + kRollGap)) {
            fdtd_set_error("fdtd_ctx_peer_connect: entry %d is not the export of rank %d of this cavity", r, r);
            return FDTD_E_ARG;
        }
        every_pong = every_pong && b.has_pong;
    }
    settle_fused(c, every_pong);
    if (c->rank > 0) {
        FDTD_TRY(map_peer(c, all[c->rank - 1], every_pong, c->peer_lo, &c->peer_sig_lo));
        c->peer_nk_lo = all[c->rank - 1].nk;
        c->roll_peer_z_lo = all[c->rank - 1].nk + 4 + kRollGap;
        c->peer_elems_lo = (size_t)all[c->rank - 1].array_elems;
    }
    if (c->rank + 1 < c->nranks) {
        FDTD_TRY(map_peer(c, all[c->rank + 1], every_pong, c->peer_hi, &c->peer_sig_hi));
        c->peer_elems_hi = (size_t)all[c->rank + 1].array_elems;
        c->roll_peer_z_hi = all[c->rank + 1].nk + 4 + kRollGap;
    }
    c->transport = TR_FLAG;
    c->flip = 0;
    c->wired = true;
    return FDTD_OK;
}

} /* extern "C" */
