/*
 * fdtd_hosted.cu -- fdtd_run_hosted: the stepping loop of main.c:765-779 for a caller whose fields
 * live in HOST memory, like the reference's (its update functions mutate the host arrays in place).
 *
 * Done naively this is three phases in sequence -- upload 51.6 GB, step, download 51.6 GB at 1024^3 --
 * and for a short run the PCIe copies are most of the time.  But a time step only couples a z-plane to
 * its two neighbours, so the three phases can run as one wavefront over z-chunks:
 *
 *     step s of chunk c  needs  step s-1 of chunks c-1, c, c+1
 *
 * Chunks are uploaded bottom-up on a copy stream.  As soon as chunk u has landed, "wave" u runs on the
 * compute stream: step 1 of chunk u-1, step 2 of chunk u-2, ... step s of chunk u-s.  When step K of a
 * chunk is done, that chunk goes back to the host on a third stream -- the other PCIe direction --
 * while higher chunks are still arriving.  The download thus trails the upload by only K chunks and
 * the whole call takes about one upload time instead of upload + K steps + download.
 *
 * Hazards.  The fused step reads one buffer set and writes the other (fdtd_fused_tma.cuh): step s
 * reads X[s-1] and writes X[s], and X[s] shares its memory with X[s-2].  Within a wave the launches
 * are queued in the order s = 1, 2, ...: (s-1, c+1) -- the last reader of what (s, c) overwrites, and
 * the last producer of what it reads -- is the launch queued right before (s, c), and everything
 * else it depends on belongs to earlier waves of the same in-order stream.
 *
 * The arithmetic per element is untouched, so the result is bit-identical to upload + fdtd_run +
 * download (tests/test_gpu_hosted.py).  Long runs use the wavefront only to ramp in (first steps,
 * during the upload) and to ramp out (last steps, during the download); the steps in between run as
 * whole-grid launches.  Multi-slab contexts and the in-place kernels (0, 1) take the plain sequence.
 */
#include "fdtd_ctx.hpp"

using namespace fdtdi;

namespace fdtdi {

struct Chunks {
    int planes;              /* local planes 1 .. planes are swept (nk + 1 on the last slab) */
    int size, count;
    int last_extra = 0;      /* planes the last chunk takes on top of `size` (so that it is never a single plane) */
    int begin(int c) const { return 1 + c * size; }
    int end(int c) const { return c == count - 1 ? planes + 1 : std::min(1 + (c + 1) * size, planes + 1); }
};

static double *host_array(const fdtd_fields *h, int idx)
{
    double *a[6] = {h->Ex, h->Ey, h->Ez, h->Hx, h->Hy, h->Hz};
    return a[idx];
}

/* `nsteps` time steps as a wavefront; the state before them is c->f.  upload: chunk u's planes
 * arrive from `host` first; download: each chunk leaves for `host` after its last step.  With the
 * two-step kernel a wave element is a sweep of TWO steps (an odd count ends with a single-step sweep);
 * the dependency pattern is the same, a sweep reaches two planes into the neighbouring chunks and
 * chunks are at least two planes thick.  On return c->f is the final state. */
static int wavefront(fdtd_ctx *c, const fdtd_fields *host, const Chunks &ch, int nsteps_in, bool upload, bool download,
                     const double *src_rows)
{
    const int M = ch.count;
    const size_t row = 2 * (size_t)c->src_n;
    const bool pairs = c->opt_kernel == 4;
    const int nsteps = pairs ? (nsteps_in + 1) / 2 : nsteps_in; /* sweeps */
    struct Events { /* destroyed on every way out */
        std::vector<cudaEvent_t> v;
        ~Events()
        {
            for (cudaEvent_t e : v)
                if (e)
                    cudaEventDestroy(e);
        }
    } ups, dones;
    ups.v.assign(upload ? M : 0, nullptr);
    dones.v.assign(download ? M : 0, nullptr);
    std::vector<cudaEvent_t> &ev_up = ups.v, &ev_done = dones.v;
    for (auto &e : ev_up)
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : ev_done)
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = FDTD_OK;
    if (upload) {
        /* the uploads overwrite the current set: everything queued so far must have finished with it */
        cudaEvent_t ev0 = ev_up[0];
        CUDA_TRY(cudaEventRecord(ev0, c->s_main));
        CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, ev0, 0));
        for (int u = 0; u < M && rc == FDTD_OK; ++u) {
            for (int a = 0; a < 6 && rc == FDTD_OK; ++a)
                rc = copy_planes(c, a, host_array(host, a), ch.begin(u), ch.end(u), true, c->s_h2d);
            if (rc == FDTD_OK && cudaEventRecord(ev_up[u], c->s_h2d) != cudaSuccess)
                rc = FDTD_E_CUDA;
        }
    }
    /* X[0] is the current set; parity p: is c->f currently X[even] (0) or X[odd] (1)? */
    int parity = 0;
    for (int u = 1; u <= M - 1 + nsteps && rc == FDTD_OK; ++u) {
        /* step 1 of chunk u-1 reads chunks u-2 .. u: the copy stream is in order, so one event does */
        if (upload && u <= M)
            CUDA_TRY(cudaStreamWaitEvent(c->s_main, ev_up[std::min(u, M - 1)], 0));
        for (int s = std::max(1, u - M + 1); s <= std::min(nsteps, u); ++s) {
            const int k = u - s; /* chunk */
            if (((s - 1) & 1) != parity) { /* step s reads X[s-1] */
                swap_buffers(c);
                parity ^= 1;
            }
            if (pairs && 2 * s <= nsteps_in) {
                const int rc2 = launch_step2(c, make_src(c, src_rows + (size_t)(2 * s - 2) * row),
                                             make_src(c, src_rows + (size_t)(2 * s - 1) * row), ch.begin(k), ch.end(k), c->s_main);
                if (rc2 != FDTD_OK)
                    c->launch_error = rc2;
            } else {
                const int step = pairs ? nsteps_in : s; /* the odd last step of a two-step run */
                launch_fused(c, make_src(c, src_rows + (size_t)(step - 1) * row), ch.begin(k), ch.end(k), c->s_main);
            }
            if (c->launch_error != FDTD_OK) {
                rc = c->launch_error;
                c->launch_error = FDTD_OK;
                break;
            }
            if (s == nsteps && download) {
                /* chunk k is final: it lives in the set step nsteps wrote, i.e. c->f2 right now */
                if (cudaEventRecord(ev_done[k], c->s_main) != cudaSuccess ||
                    cudaStreamWaitEvent(c->s_dump, ev_done[k], 0) != cudaSuccess) {
                    rc = FDTD_E_CUDA;
                    break;
                }
                swap_buffers(c); /* copy_planes reads c->f */
                for (int a = 0; a < 6 && rc == FDTD_OK; ++a)
                    rc = copy_planes(c, a, host_array(host, a), ch.begin(k), ch.end(k), false, c->s_dump);
                swap_buffers(c);
            }
        }
    }
    if (rc == FDTD_OK && cudaGetLastError() != cudaSuccess)
        rc = FDTD_E_CUDA;
    /* make c->f the final state X[nsteps] */
    if ((nsteps & 1) != parity)
        swap_buffers(c);
    if (rc == FDTD_OK && download) {
        /* later work on the compute stream must not overwrite what is still leaving */
        cudaEvent_t e = ev_done[0];
        if (cudaEventRecord(e, c->s_dump) != cudaSuccess || cudaStreamWaitEvent(c->s_main, e, 0) != cudaSuccess)
            rc = FDTD_E_CUDA;
    }
    if (rc == FDTD_E_CUDA)
        fdtd_set_error("fdtd_run_hosted: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

/* The same wavefront for a slab with neighbours (one process per GPU, peer-memory halos).  A sweep of the
 * chunk next to an interface needs the neighbour's boundary planes of the previous sweep, so the slabs'
 * wavefronts have to mesh: neighbouring slabs sweep their chunks in OPPOSITE directions (even ranks
 * bottom-up, odd ranks top-down).  Then every interface is either the first thing both of its slabs
 * touch in a wave or the last, the two sides reach it in the same wave, and what one side needs from
 * the other is always one sweep older -- each slab streams upload, sweeps and download at its own PCIe
 * rate, and the halo planes of sweep s travel right after the boundary chunk of sweep s.  Everything
 * is queued on the compute stream; sequence flags order it against the neighbours (fdtd_halo.cu). */
static int wavefront_slab(fdtd_ctx *c, const fdtd_fields *host, const Chunks &ch, int nsteps_in, const double *src_rows)
{
    const int M = ch.count;
    const size_t row = 2 * (size_t)c->src_n;
    const int S = (nsteps_in + 1) / 2; /* sweeps of two steps (the last one single when the count is odd) */
    const bool up_first = (c->rank & 1) == 0;
    const bool has_lo = c->rank > 0, has_hi = c->rank + 1 < c->nranks;
    auto chunk_of = [&](int q) { return up_first ? q : M - 1 - q; };
    const unsigned base_h = c->n_xh, base_e = c->n_xe; /* lane numbering so far; push p of a lane = base + 1 + p */
    const int flip0 = c->flip;
    struct Events {
        std::vector<cudaEvent_t> v;
        ~Events()
        {
            for (cudaEvent_t e : v)
                if (e)
                    cudaEventDestroy(e);
        }
    } ups, dones;
    ups.v.assign(M, nullptr);
    dones.v.assign(M, nullptr);
    for (auto &e : ups.v)
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : dones.v)
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    FDTD_TRY(wait_halos(c)); /* nothing of earlier exchanges is still in flight */
    CUDA_TRY(cudaEventRecord(ups.v[0], c->s_main));
    CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, ups.v[0], 0));
    for (int q = 0; q < M; ++q) {
        const int k = chunk_of(q);
        for (int a = 0; a < 6; ++a)
            FDTD_TRY(copy_planes(c, a, host_array(host, a), ch.begin(k), ch.end(k), true, c->s_h2d));
        CUDA_TRY(cudaEventRecord(ups.v[q], c->s_h2d));
    }
    /* which interfaces does the chunk at order position q touch? */
    auto touches_top = [&](int q) { return has_hi && chunk_of(q) == M - 1; };
    auto touches_bottom = [&](int q) { return has_lo && chunk_of(q) == 0; };
    auto push = [&](bool top, int p, double *src_base) { /* my boundary planes of X[p] */
        const unsigned b = top ? base_h : base_e;          /* outgoing lane: upward at the top, downward at the bottom */
        const unsigned bin = top ? base_e : base_h;        /* incoming lane of that side */
        return halo_side_push(c, top, bin + (unsigned)p, b + 1 + (unsigned)p, src_base, (flip0 + p) & 1);
    };
    int parity = 0;
    bool pushed0_top = false, pushed0_bottom = false;
    for (int u = 1; u <= M - 1 + S; ++u) {
        if (u <= M)
            CUDA_TRY(cudaStreamWaitEvent(c->s_main, ups.v[std::min(u, M - 1)], 0));
        /* the freshly uploaded boundary planes are the neighbours' halo of the initial state */
        const int landed = std::min(u, M - 1);
        for (int q = 0; q <= landed; ++q) {
            if (touches_top(q) && !pushed0_top && q <= landed) {
                if (parity != 0) { swap_buffers(c); parity = 0; }
                FDTD_TRY(push(true, 0, c->base));
                pushed0_top = true;
            }
            if (touches_bottom(q) && !pushed0_bottom && q <= landed) {
                if (parity != 0) { swap_buffers(c); parity = 0; }
                FDTD_TRY(push(false, 0, c->base));
                pushed0_bottom = true;
            }
        }
        for (int s = std::max(1, u - M + 1); s <= std::min(S, u); ++s) {
            const int q = u - s, k = chunk_of(q);
            if (((s - 1) & 1) != parity) { /* sweep s reads X[s-1] */
                swap_buffers(c);
                parity ^= 1;
            }
            /* the neighbours' boundary planes of X[s-1] */
            if (touches_top(q))
                FDTD_TRY(halo_side_wait(c, true, base_e + (unsigned)s));
            if (touches_bottom(q))
                FDTD_TRY(halo_side_wait(c, false, base_h + (unsigned)s));
            if (2 * s <= nsteps_in) {
                const int rc2 = launch_step2(c, make_src(c, src_rows + (size_t)(2 * s - 2) * row),
                                             make_src(c, src_rows + (size_t)(2 * s - 1) * row), ch.begin(k), ch.end(k), c->s_main);
                if (rc2 != FDTD_OK)
                    return rc2;
            } else {
                launch_fused(c, make_src(c, src_rows + (size_t)(nsteps_in - 1) * row), ch.begin(k), ch.end(k), c->s_main);
                if (c->launch_error != FDTD_OK) {
                    const int rc2 = c->launch_error;
                    c->launch_error = FDTD_OK;
                    return rc2;
                }
            }
            /* my boundary planes of X[s] (just written into the other set) for the neighbours' sweep s + 1 */
            if (touches_top(q))
                FDTD_TRY(push(true, s, c->base2));
            if (touches_bottom(q))
                FDTD_TRY(push(false, s, c->base2));
            if (s == S) {
                CUDA_TRY(cudaEventRecord(dones.v[q], c->s_main));
                CUDA_TRY(cudaStreamWaitEvent(c->s_dump, dones.v[q], 0));
                swap_buffers(c); /* copy_planes reads c->f */
                int rc2 = FDTD_OK;
                for (int a = 0; a < 6 && rc2 == FDTD_OK; ++a)
                    rc2 = copy_planes(c, a, host_array(host, a), ch.begin(k), ch.end(k), false, c->s_dump);
                swap_buffers(c);
                FDTD_TRY(rc2);
            }
        }
    }
    CUDA_TRY(cudaGetLastError());
    if ((S & 1) != parity)
        swap_buffers(c);
    /* the lanes have seen one initial push and S more; what the neighbours sent last is the halo of X[S] */
    c->n_xh = base_h + 1 + (unsigned)S;
    c->n_xe = base_e + 1 + (unsigned)S;
    CUDA_TRY(cudaEventRecord(c->ev_sent, c->s_main));
    c->e_halo_valid = c->h_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = true;
    CUDA_TRY(cudaEventRecord(dones.v[0], c->s_dump));
    CUDA_TRY(cudaStreamWaitEvent(c->s_main, dones.v[0], 0));
    return FDTD_OK;
}

} /* namespace fdtdi */

extern "C" {

int fdtd_run_hosted(fdtd_ctx *c, const fdtd_fields *host, size_t steps, double *time_counter)
{
    FDTD_TRY(check_solo(c, "fdtd_run_hosted"));
    if (!host || !host->Ex || !host->Ey || !host->Ez || !host->Hx || !host->Hy || !host->Hz || !time_counter) {
        fdtd_set_error("fdtd_run_hosted: NULL argument");
        return FDTD_E_ARG;
    }
    FDTD_TRY(use_device(c));
    FDTD_TRY(settle_kernel(c));
    const int kWave = 32; /* steps ramped in / out as a wavefront in a long run */
    const bool pipelined = c->opt_host_pipeline && c->nranks == 1 && c->opt_kernel >= 2 && !c->rolling && steps >= 1 &&
                           steps <= (size_t)1 << 30;
    /* slabs: the meshed wavefront needs the two-step kernel's wide halos, peer-memory flags, a second buffer
     * set and a run short enough to be one wavefront (the same decision on every rank) */
    const bool meshed = c->opt_host_pipeline && c->nranks > 1 && c->transport == TR_FLAG && c->opt_kernel == 4 &&
                        !c->rolling && step2_usable(c) && steps >= 1 && steps <= 2 * (size_t)kWave;
    if (meshed) {
        if (!c->s_h2d)
            CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        Chunks ch;
        ch.planes = c->g.nk + c->g.top;
        ch.size = c->opt_host_chunk > 0 ? (int)c->opt_host_chunk : std::max(8, (ch.planes + 255) / 256);
        ch.size = std::max(ch.size, 2);
        ch.count = (ch.planes + ch.size - 1) / ch.size;
        if (ch.count >= 2 && ch.planes - (ch.count - 1) * ch.size < 2) /* the last chunk carries two boundary planes */
            ch.count -= 1, ch.last_extra = ch.planes - ch.count * ch.size;
        double t = *time_counter;
        FDTD_TRY(stage_source_rows(c, steps, &t));
        FDTD_TRY(wavefront_slab(c, host, ch, (int)steps, c->src_dev));
        CUDA_TRY(cudaStreamSynchronize(c->s_dump));
        CUDA_TRY(cudaStreamSynchronize(c->s_main));
        *time_counter = t;
        return FDTD_OK;
    }
    if (!pipelined) {
        FDTD_TRY(fdtd_upload_slab(c, host));
        FDTD_TRY(run_impl(c, steps, time_counter, nullptr, nullptr, nullptr));
        return fdtd_download_slab(c, host);
    }
    if (!c->s_h2d)
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    Chunks ch;
    ch.planes = c->g.nk + c->g.top;
    /* 8 planes per chunk keep the download close behind the upload; at most 256 chunks */
    ch.size = c->opt_host_chunk > 0 ? (int)c->opt_host_chunk : std::max(8, (ch.planes + 255) / 256);
    ch.size = std::max(ch.size, 2);
    ch.count = (ch.planes + ch.size - 1) / ch.size;

    double t = *time_counter;
    auto stage = [&](size_t count) { /* source rows of the next `count` steps -> c->src_dev */
        return stage_source_rows(c, count, &t);
    };
    int rc;
    if (steps <= 2 * (size_t)kWave) {
        FDTD_TRY(stage(steps));
        rc = wavefront(c, host, ch, (int)steps, true, true, c->src_dev);
    } else {
        FDTD_TRY(stage(kWave));
        rc = wavefront(c, host, ch, kWave, true, false, c->src_dev);
        if (rc == FDTD_OK)
            rc = run_impl(c, steps - 2 * kWave, &t, nullptr, nullptr, nullptr);
        if (rc == FDTD_OK)
            rc = stage(kWave);
        if (rc == FDTD_OK)
            rc = wavefront(c, host, ch, kWave, false, true, c->src_dev);
    }
    FDTD_TRY(rc);
    CUDA_TRY(cudaStreamSynchronize(c->s_dump));
    CUDA_TRY(cudaStreamSynchronize(c->s_main));
    *time_counter = t;
    return FDTD_OK;
}

} /* extern "C" */
