/*
 * fdtd_group.cu -- all z-slabs of a cavity driven by ONE host thread (fdtd_group_*).
 *
 * The reference is a single-threaded C program; this keeps its host program one: no MPI, no
 * launcher.  A group is n slab contexts, normally one per GPU of the box.  Every call only queues
 * work; the slabs advance concurrently on their own streams.  Halo planes travel by peer copies
 * ordered with CUDA events (TR_EVENT, fdtd_halo.cu) or, on request, by NCCL send/recv through a
 * communicator made by ncclCommInitAll; because one thread issues the transfers of all slabs, every
 * exchange is queued in phases over all slabs (exchange_many).
 *
 * The one-process-per-GPU route (fdtd_ctx_create_slab + fdtd_ctx_comm_init / fdtd_ctx_peer_connect)
 * runs the same segments for its one slab.
 */
#include "fdtd_ctx.hpp"

#include <vector>

using namespace fdtdi;

struct fdtd_group {
    std::vector<fdtd_ctx *> ctx;
};

namespace fdtdi {

/* dumps read node plane k1 of Ex, Ey and Hz (main.c:517-520, 538-539) */
int exchange_many_for_dump(fdtd_ctx *const *cs, int n)
{
    if (cs[0]->nranks == 1)
        return FDTD_OK;
    Xchg x{};
    x.e = x.e_with_hz = true;
    for (int r = 0; r < n; ++r)
        cs[r]->e_halo_valid = false;
    FDTD_TRY(exchange_many(cs, n, x, false));
    for (int r = 0; r < n; ++r)
        cs[r]->e_halo_valid = true;
    return FDTD_OK;
}

static int group_run(fdtd_ctx *const *cs, int n, size_t steps, double *time_counter)
{
    /* kernel choice must be the same on every slab: the halo plans differ */
    bool want_fused = cs[0]->opt_kernel >= 2;
    const bool can_roll = cs[0]->opt_kernel == 4 && step2_usable(cs[0]);
    for (int r = 0; r < n; ++r)
        cs[r]->rolling = can_roll && cs[0]->opt_rolling;
    if (want_fused && !cs[0]->rolling) {
        bool nomem = false;
        for (int r = 0; r < n; ++r) {
            FDTD_TRY(use_device(cs[r]));
            const int rc = ensure_pong(cs[r]);
            if (rc == FDTD_E_NOMEM && (cs[r]->kernel_auto || can_roll))
                nomem = true;
            else if (rc != FDTD_OK)
                return rc;
        }
        if (nomem && can_roll) { /* in place on the rolling window instead of a second set */
            if (!cs[0]->rolling)
                fprintf(stderr, "[fdtd_b200] a second copy of the state does not fit in HBM on every slab: the two-step kernel "
                                "works in place on a rolling window of plane slots\n");
            for (int r = 0; r < n; ++r)
                cs[r]->rolling = true;
            cudaGetLastError();
        } else if (nomem) {
            for (int r = 0; r < n; ++r) {
                FDTD_TRY(use_device(cs[r]));
                fall_back_to_split(cs[r]);
            }
            want_fused = false;
        }
    }
    const bool pairs = want_fused && cs[0]->opt_kernel == 4 && step2_usable(cs[0]);
    if (!(pairs && steps >= 2))
        FDTD_TRY(refresh_halos_many(cs, n, want_fused));

    const Segment fused_plan[1] = {SEG_FUSED}, split_plan[2] = {SEG_H, SEG_E};
    const Segment *plan = want_fused ? fused_plan : split_plan;
    const int nseg = want_fused ? 1 : 2;
    double t = *time_counter;
    for (size_t done = 0; done < steps;) {
        const size_t chunk = std::min(steps - done, (size_t)kSrcRing);
        double t_chunk = t;
        for (int r = 0; r < n; ++r) {
            t_chunk = t;
            FDTD_TRY(use_device(cs[r]));
            FDTD_TRY(stage_source_rows(cs[r], chunk, &t_chunk));
        }
        for (size_t s = 0; s < chunk; ++s) {
            if (pairs && s + 1 < chunk) { /* two steps in one sweep on every slab */
                FDTD_TRY(refresh_halos_many(cs, n, true, true));
                for (int r = 0; r < n; ++r) {
                    fdtd_ctx *c = cs[r];
                    const size_t row = 2 * (size_t)c->src_n;
                    const fdtd::Src s1 = make_src(c, c->src_dev + s * row), s2 = make_src(c, c->src_dev + (s + 1) * row);
                    FDTD_TRY(seg_launch(c, s1, SEG_STEP2, &s2));
                }
                FDTD_TRY(exchange_many(cs, n, seg_xchg(SEG_STEP2), true));
                ++s;
                continue;
            }
            if (cs[0]->rolling) {
                /* the odd step of a rolling run: in-place split kernels on the canonical layout */
                long keep[5] = {cs[0]->opt_kernel, cs[0]->opt_strip, cs[0]->opt_kchunk, cs[0]->opt_wx, cs[0]->opt_wy};
                int rc1 = FDTD_OK;
                for (int r = 0; r < n && rc1 == FDTD_OK; ++r) {
                    rc1 = use_device(cs[r]);
                    if (rc1 == FDTD_OK)
                        rc1 = roll_canonicalise(cs[r]);
                    cs[r]->opt_kernel = 1; cs[r]->opt_strip = 2; cs[r]->opt_kchunk = 8; cs[r]->opt_wx = 2; cs[r]->opt_wy = 2;
                }
                if (rc1 == FDTD_OK)
                    rc1 = refresh_halos_many(cs, n, false);
                for (int g = 0; g < 2 && rc1 == FDTD_OK; ++g) {
                    for (int r = 0; r < n && rc1 == FDTD_OK; ++r) {
                        fdtd_ctx *c = cs[r];
                        rc1 = seg_launch(c, make_src(c, c->src_dev + s * 2 * (size_t)c->src_n), split_plan[g]);
                    }
                    if (rc1 == FDTD_OK)
                        rc1 = exchange_many(cs, n, seg_xchg(split_plan[g]), true);
                }
                for (int r = 0; r < n; ++r) {
                    cs[r]->opt_kernel = keep[0]; cs[r]->opt_strip = keep[1]; cs[r]->opt_kchunk = keep[2];
                    cs[r]->opt_wx = keep[3]; cs[r]->opt_wy = keep[4];
                }
                FDTD_TRY(rc1);
                continue;
            }
            if (pairs)
                FDTD_TRY(refresh_halos_many(cs, n, true));
            for (int g = 0; g < nseg; ++g) {
                for (int r = 0; r < n; ++r) {
                    fdtd_ctx *c = cs[r];
                    const fdtd::Src src = make_src(c, c->src_dev + s * 2 * (size_t)c->src_n);
                    FDTD_TRY(seg_launch(c, src, plan[g]));
                }
                FDTD_TRY(exchange_many(cs, n, seg_xchg(plan[g]), true));
            }
        }
        t = t_chunk;
        done += chunk;
    }
    *time_counter = t;
    for (int r = 0; r < n; ++r) { /* (rolling form only) back to the canonical layout */
        FDTD_TRY(use_device(cs[r]));
        FDTD_TRY(roll_canonicalise(cs[r]));
    }
    return FDTD_OK;
}

int step_many(fdtd_ctx *const *cs, int n, size_t steps, double *time_counter)
{
    if (n == 1)
        return run_impl(cs[0], steps, time_counter, nullptr, nullptr, nullptr);
    return group_run(cs, n, steps, time_counter);
}

} /* namespace fdtdi */

/* ============================================ C ABI ========================================= */

extern "C" {

/* transport: 0 = peer copies ordered by events (TR_EVENT; the default, also for slabs that share a
 * device), 1 = NCCL send/recv (needs one distinct GPU per slab) */
int fdtd_group_create_transport(const fdtd_params *p, int ngpus, const int *devices, int transport, fdtd_group **out)
{
    if (!p || !out || ngpus < 1 || ngpus > 64 || transport < 0 || transport > 1) {
        fdtd_set_error("fdtd_group_create: bad argument (ngpus %d, transport %d)", ngpus, transport);
        return FDTD_E_ARG;
    }
    fdtd_group *g = new (std::nothrow) fdtd_group();
    if (!g) {
        fdtd_set_error("fdtd_group_create: out of host memory");
        return FDTD_E_NOMEM;
    }
    std::vector<int> dev(ngpus);
    for (int r = 0; r < ngpus; ++r)
        dev[r] = devices ? devices[r] : r;
    int rc = FDTD_OK;
    for (int r = 0; r < ngpus && rc == FDTD_OK; ++r) {
        fdtd_ctx *c = nullptr;
        rc = create_impl(p, dev[r], r, ngpus, &c);
        if (rc == FDTD_OK)
            g->ctx.push_back(c);
    }
    if (rc == FDTD_OK && ngpus > 1 && transport == 1) {
        rc = nccl_bind();
        if (rc == FDTD_OK) {
            std::vector<ncclComm_t> comms(ngpus);
            ncclResult_t e = g_nccl.CommInitAll(comms.data(), ngpus, dev.data());
            if (e != ncclSuccess) {
                fdtd_set_error("ncclCommInitAll: %s", g_nccl.GetErrorString(e));
                rc = FDTD_E_NCCL;
            } else {
                for (int r = 0; r < ngpus; ++r) {
                    g->ctx[r]->comm = comms[r];
                    g->ctx[r]->has_comm = true;
                    g->ctx[r]->transport = TR_NCCL;
                }
            }
        }
    } else if (rc == FDTD_OK && ngpus > 1) {
        for (int r = 0; r < ngpus; ++r) {
            fdtd_ctx *c = g->ctx[r];
            c->transport = TR_EVENT;
            c->nb_lo = r > 0 ? g->ctx[r - 1] : nullptr;
            c->nb_hi = r + 1 < ngpus ? g->ctx[r + 1] : nullptr;
            /* direct NVLink copies where the devices can reach each other; otherwise the peer copy is
             * staged by the driver */
            for (fdtd_ctx *nb : {c->nb_lo, c->nb_hi}) {
                if (!nb || nb->device == c->device)
                    continue;
                int can = 0;
                if (cudaSetDevice(c->device) == cudaSuccess &&
                    cudaDeviceCanAccessPeer(&can, c->device, nb->device) == cudaSuccess && can)
                    cudaDeviceEnablePeerAccess(nb->device, 0);
                cudaGetLastError(); /* "already enabled" is fine */
            }
        }
    }
    if (rc == FDTD_OK)
        for (fdtd_ctx *c : g->ctx) {
            c->in_group = ngpus > 1;
            c->wired = true;
        }
    if (rc != FDTD_OK) {
        for (fdtd_ctx *c : g->ctx)
            fdtd_ctx_destroy(c);
        delete g;
        return rc;
    }
    *out = g;
    return FDTD_OK;
}

int fdtd_group_create(const fdtd_params *p, int ngpus, const int *devices, fdtd_group **out)
{
    const char *env = getenv("FDTD_B200_TRANSPORT");
    return fdtd_group_create_transport(p, ngpus, devices, env && !strcmp(env, "nccl") ? 1 : 0, out);
}

int fdtd_group_destroy(fdtd_group *g)
{
    if (!g)
        return FDTD_OK;
    for (fdtd_ctx *c : g->ctx) /* a slab's streams may still wait on a neighbour's events */
        fdtd_sync(c);
    for (fdtd_ctx *c : g->ctx) {
        c->wired = false; /* the neighbours are about to go away */
        c->nb_lo = c->nb_hi = nullptr;
    }
    for (fdtd_ctx *c : g->ctx)
        fdtd_ctx_destroy(c);
    delete g;
    return FDTD_OK;
}

int fdtd_group_size(fdtd_group *g)
{
    return g ? (int)g->ctx.size() : 0;
}

int fdtd_group_ctx(fdtd_group *g, int rank, fdtd_ctx **out)
{
    if (!g || !out || rank < 0 || rank >= (int)g->ctx.size()) {
        fdtd_set_error("fdtd_group_ctx: bad argument (rank %d)", rank);
        return FDTD_E_ARG;
    }
    *out = g->ctx[rank];
    return FDTD_OK;
}

#define GROUP_CHECK(who)                                                                       \
    if (!g || g->ctx.empty()) {                                                                \
        fdtd_set_error(who ": group is NULL");                                                 \
        return FDTD_E_ARG;                                                                     \
    }

int fdtd_group_set_option(fdtd_group *g, const char *key, long value)
{
    GROUP_CHECK("fdtd_group_set_option");
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(fdtd_ctx_set_option(c, key, value));
    return FDTD_OK;
}

int fdtd_group_upload(fdtd_group *g, const fdtd_fields *whole_cavity)
{
    GROUP_CHECK("fdtd_group_upload");
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(fdtd_upload(c, whole_cavity));
    return FDTD_OK;
}

int fdtd_group_download(fdtd_group *g, const fdtd_fields *whole_cavity)
{
    GROUP_CHECK("fdtd_group_download");
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(fdtd_download(c, whole_cavity));
    return FDTD_OK;
}

int fdtd_group_set_initial_conditions(fdtd_group *g)
{
    GROUP_CHECK("fdtd_group_set_initial_conditions");
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(fdtd_set_initial_conditions(c));
    return FDTD_OK;
}

int fdtd_group_run(fdtd_group *g, size_t steps, double *time_counter)
{
    GROUP_CHECK("fdtd_group_run");
    if (!time_counter) {
        fdtd_set_error("fdtd_group_run: time_counter is NULL");
        return FDTD_E_ARG;
    }
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(check_ctx(c, "fdtd_group_run"));
    return step_many(g->ctx.data(), (int)g->ctx.size(), steps, time_counter);
}

/* fdtd_aggregate for the whole cavity: host_out has maxi*maxj*maxk doubles (main.c:511-540) */
int fdtd_group_aggregate(fdtd_group *g, int var, double *host_out)
{
    GROUP_CHECK("fdtd_group_aggregate");
    if (!host_out) {
        fdtd_set_error("fdtd_group_aggregate: host_out is NULL");
        return FDTD_E_ARG;
    }
    std::vector<double *> out(g->ctx.size());
    for (size_t r = 0; r < g->ctx.size(); ++r)
        out[r] = host_out + g->ctx[r]->k0 * (size_t)g->ctx[r]->g.I * (size_t)g->ctx[r]->g.J;
    return aggregate_many(g->ctx.data(), (int)g->ctx.size(), var, out.data());
}

/* fdtd_energy summed over the slabs (main.c:602-668) */
int fdtd_group_energy(fdtd_group *g, int as_coded, double *e_energy, double *h_energy)
{
    GROUP_CHECK("fdtd_group_energy");
    return energy_many(g->ctx.data(), (int)g->ctx.size(), as_coded, e_energy, h_energy);
}

int fdtd_group_sync(fdtd_group *g)
{
    GROUP_CHECK("fdtd_group_sync");
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(fdtd_sync(c));
    return FDTD_OK;
}

int fdtd_group_propagate(fdtd_group *g, const fdtd_dump_sink *sinks, size_t *steps_done, double *time_counter)
{
    GROUP_CHECK("fdtd_group_propagate");
    for (fdtd_ctx *c : g->ctx)
        FDTD_TRY(check_ctx(c, "fdtd_group_propagate"));
    return propagate_many(g->ctx.data(), (int)g->ctx.size(), sinks, steps_done, time_counter);
}

} /* extern "C" */
