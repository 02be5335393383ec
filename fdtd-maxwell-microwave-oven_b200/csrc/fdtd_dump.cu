/*
 * fdtd_dump.cu -- the dump side of the path: the zone-centred variables of write_silo (main.c:550-598)
 * and fdtd_propagate, the reference's stepping loop with its dump cadence (main.c:755-799).
 */
#include "fdtd_ctx.hpp"
#include "fdtd_dump_kernels.cuh"

using namespace fdtdi;

namespace fdtdi {


constexpr int kMaxDumpVars = 9;
const char *const kVarNames[kMaxDumpVars] = {"ex", "ey", "ez", "hx", "hy", "hz", "aEy", "aHx", "aHz"};

/* Writer-side state.  The compute thread aggregates every variable of one dump into HBM scratch
 * (dev[v]) on the compute stream and posts the iteration number; the writer thread drains the
 * scratch through two pinned buffers on the dump stream and feeds the sink.  The compute thread
 * blocks only if the next dump is due before the previous one has left HBM. */
struct DumpPipe {
    fdtd_ctx *ctx;
    fdtd_dump_sink sink;
    size_t n;                   /* doubles per variable */
    int nvars;                  /* 6 or 9 */
    double *dev[7];             /* ex..hz, aEy (aHx/aHz alias hx/hz, main.c:585-588) */
    double *pinned[2];
    cudaEvent_t ev_agg, ev_copy[2];
    double *sk_dev, *si_dev;    /* analytic factors for aEy */
    double f_mnl;

    pthread_t thread;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int pending_iteration;      /* -1: none */
    bool scratch_busy;          /* dev[] still being drained */
    bool writer_busy;           /* a dump is between sink.begin and sink.end */
    bool stop;
    int error;                  /* first sink / CUDA failure seen by the writer */
    char error_msg[256];
};

void *writer_main(void *arg)
{
    DumpPipe *dp = (DumpPipe *)arg;
    fdtd_ctx *c = dp->ctx;
    cudaSetDevice(c->device);
    for (;;) {
        pthread_mutex_lock(&dp->mu);
        while (dp->pending_iteration < 0 && !dp->stop)
            pthread_cond_wait(&dp->cv, &dp->mu);
        if (dp->pending_iteration < 0 && dp->stop) {
            pthread_mutex_unlock(&dp->mu);
            return nullptr;
        }
        const int iteration = dp->pending_iteration;
        dp->pending_iteration = -1;
        dp->writer_busy = true;
        pthread_mutex_unlock(&dp->mu);

        int err = 0;
        const char *what = "";
        const size_t dims[3] = {(size_t)c->g.I, (size_t)c->g.J, (size_t)c->g.nk};
        if (cudaStreamWaitEvent(c->s_dump, dp->ev_agg, 0) != cudaSuccess) { err = FDTD_E_CUDA; what = "wait for aggregation"; }
        if (!err && dp->sink.begin && dp->sink.begin(dp->sink.user, iteration, dims, c->k0) != 0) { err = FDTD_E_IO; what = "sink.begin"; }
        auto source_of = [&](int v) { return v < 7 ? dp->dev[v] : (v == 7 ? dp->dev[3] : dp->dev[5]); };
        auto start_copy = [&](int v) {
            if (cudaMemcpyAsync(dp->pinned[v & 1], source_of(v), dp->n * sizeof(double), cudaMemcpyDeviceToHost, c->s_dump) != cudaSuccess ||
                cudaEventRecord(dp->ev_copy[v & 1], c->s_dump) != cudaSuccess) {
                err = FDTD_E_CUDA;
                what = "device-to-host copy";
            }
        };
        if (!err)
            start_copy(0);
        for (int v = 0; v < dp->nvars && !err; ++v) {
            if (cudaEventSynchronize(dp->ev_copy[v & 1]) != cudaSuccess) { err = FDTD_E_CUDA; what = "device-to-host copy"; break; }
            if (v + 1 < dp->nvars)
                start_copy(v + 1); /* overlaps with the sink consuming variable v */
            else {
                /* the last variable has left HBM: the compute thread may aggregate the next dump */
                pthread_mutex_lock(&dp->mu);
                dp->scratch_busy = false;
                pthread_cond_broadcast(&dp->cv);
                pthread_mutex_unlock(&dp->mu);
            }
            if (!err && dp->sink.variable &&
                dp->sink.variable(dp->sink.user, kVarNames[v], dp->pinned[v & 1], dp->n) != 0) { err = FDTD_E_IO; what = "sink.variable"; }
        }
        if (!err && dp->sink.end && dp->sink.end(dp->sink.user) != 0) { err = FDTD_E_IO; what = "sink.end"; }
        pthread_mutex_lock(&dp->mu);
        if (err && !dp->error) {
            dp->error = err;
            snprintf(dp->error_msg, sizeof dp->error_msg, "dump of iteration %d failed in %s", iteration, what);
        }
        /* scratch_busy was released when the last variable left HBM; by now the compute thread may
         * already own the scratch again for the next dump, so it must not be touched here */
        dp->writer_busy = false;
        if (err)
            dp->scratch_busy = false;
        pthread_cond_broadcast(&dp->cv);
        pthread_mutex_unlock(&dp->mu);
    }
}

int pipe_create(fdtd_ctx *c, const fdtd_dump_sink *sink)
{
    DumpPipe *dp = new (std::nothrow) DumpPipe();
    if (!dp) {
        fdtd_set_error("fdtd_propagate: out of host memory");
        return FDTD_E_NOMEM;
    }
    memset(dp, 0, sizeof *dp);
    c->pipe = dp;
    dp->ctx = c;
    dp->sink = *sink;
    dp->n = (size_t)c->g.I * c->g.J * c->g.nk;
    dp->nvars = c->p.mode == 0 ? 9 : 6;
    dp->pending_iteration = -1;
    const int ndev = c->p.mode == 0 ? 7 : 6;
    for (int v = 0; v < ndev; ++v)
        CUDA_TRY(cudaMalloc((void **)&dp->dev[v], dp->n * sizeof(double)));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaHostAlloc((void **)&dp->pinned[b], dp->n * sizeof(double), cudaHostAllocDefault));
        CUDA_TRY(cudaEventCreateWithFlags(&dp->ev_copy[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&dp->ev_agg, cudaEventDisableTiming));
    if (c->p.mode == 0) {
        /* factors of the analytic TE101 solution, main.c:672 and :688-690, with the host libm */
        const fdtd_params &p = c->p;
        std::vector<double> sk(p.maxk + 2), si(p.maxi + 2);
        for (size_t k = 0; k < p.maxk + 1; ++k)
            sk[k] = sin(FDTD_PI * k * p.spatial_step / p.height);
        for (size_t i = 0; i < p.maxi + 1; ++i)
            si[i] = sin(FDTD_PI * i * p.spatial_step / p.length);
        dp->f_mnl = 0.5 * FDTD_CELERITY * sqrt(pow(FDTD_PI / p.height, 2) + pow(FDTD_PI / p.length, 2)) / FDTD_PI;
        CUDA_TRY(cudaMalloc((void **)&dp->sk_dev, sk.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc((void **)&dp->si_dev, si.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(dp->sk_dev, sk.data(), sk.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dp->si_dev, si.data(), si.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    pthread_mutex_init(&dp->mu, nullptr);
    pthread_cond_init(&dp->cv, nullptr);
    if (pthread_create(&dp->thread, nullptr, writer_main, dp) != 0) {
        fdtd_set_error("fdtd_propagate: cannot start the writer thread");
        return FDTD_E_STATE;
    }
    return FDTD_OK;
}

/* let the writer finish what is queued, then stop it */
void pipe_join(DumpPipe *dp)
{
    if (!dp->thread)
        return;
    pthread_mutex_lock(&dp->mu);
    dp->stop = true;
    pthread_cond_broadcast(&dp->cv);
    pthread_mutex_unlock(&dp->mu);
    pthread_join(dp->thread, nullptr);
    dp->thread = 0;
    pthread_mutex_destroy(&dp->mu);
    pthread_cond_destroy(&dp->cv);
}

void pipe_destroy(fdtd_ctx *c)
{
    DumpPipe *dp = c->pipe;
    if (!dp)
        return;
    pipe_join(dp);
    for (int v = 0; v < 7; ++v)
        if (dp->dev[v]) cudaFree(dp->dev[v]);
    for (int b = 0; b < 2; ++b) {
        if (dp->pinned[b]) cudaFreeHost(dp->pinned[b]);
        if (dp->ev_copy[b]) cudaEventDestroy(dp->ev_copy[b]);
    }
    if (dp->ev_agg) cudaEventDestroy(dp->ev_agg);
    if (dp->sk_dev) cudaFree(dp->sk_dev);
    if (dp->si_dev) cudaFree(dp->si_dev);
    delete dp;
    c->pipe = nullptr;
}

/* write_silo(), main.c:550-598, device side, in three parts so that several slabs driven by one
 * thread can share one NCCL group for the middle one:
 *   dump_prepare   wait until the previous dump has left the HBM scratch;
 *   (exchange)     node plane k1 of Ex, Ey, Hz for the top zone plane -- exchange_many_for_dump();
 *   dump_launch    aggregate every variable of the current state, hand the job to the writer.
 * t_validation is the time the validation fields were last evaluated for (main.c:762, :783). */
int dump_prepare(fdtd_ctx *c)
{
    DumpPipe *dp = c->pipe;
    pthread_mutex_lock(&dp->mu);
    while (dp->scratch_busy && !dp->error)
        pthread_cond_wait(&dp->cv, &dp->mu);
    const int err = dp->error;
    if (!err)
        dp->scratch_busy = true;
    pthread_mutex_unlock(&dp->mu);
    if (err) {
        fdtd_set_error("%s", dp->error_msg);
        return err;
    }
    return FDTD_OK;
}

int dump_launch(fdtd_ctx *c, int iteration, double t_validation)
{
    DumpPipe *dp = c->pipe;
    FDTD_TRY(use_device(c));
    dim3 block(64, 4);
    dim3 grid((c->g.I + 63) / 64, (c->g.J + 3) / 4, c->g.nk);
    for (int v = 0; v < 6; ++v)
        k_aggregate<<<grid, block, 0, c->s_main>>>(c->g, field_ptr(c, v), v, dp->dev[v]);
    if (c->p.mode == 0) {
        const double ct = cos(2 * FDTD_PI * dp->f_mnl * t_validation); /* main.c:688 */
        k_aggregate_aey<<<grid, block, 0, c->s_main>>>(c->g, c->f.ey, ct, dp->sk_dev, dp->si_dev, dp->dev[6]);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(dp->ev_agg, c->s_main));
    pthread_mutex_lock(&dp->mu);
    dp->pending_iteration = iteration;
    pthread_cond_broadcast(&dp->cv);
    pthread_mutex_unlock(&dp->mu);
    return FDTD_OK;
}

int post_dump_many(fdtd_ctx *const *cs, int n, int iteration, double t_validation)
{
    for (int r = 0; r < n; ++r)
        FDTD_TRY(dump_prepare(cs[r]));
    FDTD_TRY(exchange_many_for_dump(cs, n));
    for (int r = 0; r < n; ++r)
        FDTD_TRY(dump_launch(cs[r], iteration, t_validation));
    return FDTD_OK;
}

/* fdtd_aggregate for the n slabs this thread drives: host_out[r] receives slab r's zones */
int aggregate_many(fdtd_ctx *const *cs, int n, int var, double *const *host_out)
{
    if (var < 0 || var > 5) {
        fdtd_set_error("fdtd_aggregate: bad variable index %d", var);
        return FDTD_E_ARG;
    }
    for (int r = 0; r < n; ++r) {
        fdtd_ctx *c = cs[r];
        FDTD_TRY(use_device(c));
        const size_t cnt = (size_t)c->g.I * c->g.J * c->g.nk;
        if (c->agg_elems < cnt) {
            if (c->agg_dev) cudaFree(c->agg_dev);
            c->agg_dev = nullptr;
            c->agg_elems = 0;
            CUDA_TRY(cudaMalloc((void **)&c->agg_dev, cnt * sizeof(double)));
            c->agg_elems = cnt;
        }
    }
    /* zone plane k1-1 of ex, ey, hz reads node plane k1 (main.c:517-520, 538-539) */
    FDTD_TRY(exchange_many_for_dump(cs, n));
    for (int r = 0; r < n; ++r) {
        fdtd_ctx *c = cs[r];
        FDTD_TRY(use_device(c));
        const size_t cnt = (size_t)c->g.I * c->g.J * c->g.nk;
        dim3 block(64, 4);
        dim3 grid((c->g.I + 63) / 64, (c->g.J + 3) / 4, c->g.nk);
        k_aggregate<<<grid, block, 0, c->s_main>>>(c->g, field_ptr(c, var), var, c->agg_dev);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(host_out[r], c->agg_dev, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->s_main));
    }
    for (int r = 0; r < n; ++r) {
        FDTD_TRY(use_device(cs[r]));
        CUDA_TRY(cudaStreamSynchronize(cs[r]->s_main));
    }
    return FDTD_OK;
}

/* propagate_fields(), main.c:755-799, for the n slabs this thread drives (n = 1: one context) */
int propagate_many(fdtd_ctx *const *cs, int n, const fdtd_dump_sink *sinks, size_t *steps_done, double *time_counter)
{
    fdtd_ctx *c0 = cs[0];
    if (sinks && c0->p.sampling_rate == 0) {
        /* the reference divides by zero at main.c:794 */
        fdtd_set_error("fdtd_propagate: sampling_rate must be >= 1");
        return FDTD_E_ARG;
    }
    if (sinks) {
        /* scratch, pinned buffers and the writer thread are kept for the life of the context:
         * pinning two variable-sized host buffers is the expensive part (about 0.25 s per GB) */
        for (int r = 0; r < n; ++r) {
            fdtd_ctx *c = cs[r];
            FDTD_TRY(use_device(c));
            if (c->pipe && !c->pipe->error) {
                c->pipe->sink = sinks[r];
            } else {
                pipe_destroy(c);
                int rc = pipe_create(c, &sinks[r]);
                if (rc != FDTD_OK) {
                    pipe_destroy(c);
                    return rc;
                }
            }
        }
    }
    int rc = FDTD_OK;
    int iteration = 1; /* main.c:758 */
    size_t steps = 0;
    double t = 0.0;
    const float t_limit = c0->p.simulation_time;
    const int rate = (int)c0->p.sampling_rate;
    if (sinks)
        rc = post_dump_many(cs, n, iteration, 0.0); /* main.c:762-764 */
    /* main.c:765: double counter, repeated addition, float bound promoted to double, `<=` */
    while (rc == FDTD_OK && t <= t_limit) {
        /* queue every step up to the next dump in one go */
        size_t batch = 0;
        double t_probe = t, t_last = t;
        int it_probe = iteration;
        while (t_probe <= t_limit) {
            ++batch;
            t_last = t_probe;
            t_probe += c0->p.time_step;
            if (sinks && it_probe % rate == 0)
                break;
            ++it_probe;
            if (!sinks && batch >= 4096)
                break;
        }
        rc = step_many(cs, n, batch, &t);
        if (rc != FDTD_OK)
            break;
        steps += batch;
        iteration += (int)batch;
        /* main.c:794: the test runs before `iteration++`, i.e. on the index of the pass just done */
        if (sinks && (iteration - 1) % rate == 0)
            rc = post_dump_many(cs, n, iteration - 1, t_last);
    }
    for (int r = 0; r < n && rc == FDTD_OK; ++r)
        rc = fdtd_sync(cs[r]);
    for (int r = 0; r < n; ++r) {
        DumpPipe *dp = cs[r]->pipe;
        if (!dp || !sinks)
            continue;
        /* wait until the writer has delivered the last dump (its final sink.end()) */
        pthread_mutex_lock(&dp->mu);
        while ((dp->writer_busy || dp->pending_iteration >= 0) && !dp->error)
            pthread_cond_wait(&dp->cv, &dp->mu);
        if (dp->error && rc == FDTD_OK) {
            rc = dp->error;
            fdtd_set_error("%s", dp->error_msg);
        }
        pthread_mutex_unlock(&dp->mu);
    }
    if (steps_done) *steps_done = steps;
    if (time_counter) *time_counter = t;
    return rc;
}

} /* namespace fdtdi */

extern "C" {

int fdtd_aggregate(fdtd_ctx *c, int var, double *host_out)
{
    FDTD_TRY(check_solo(c, "fdtd_aggregate"));
    if (!host_out) {
        fdtd_set_error("fdtd_aggregate: host_out is NULL");
        return FDTD_E_ARG;
    }
    return aggregate_many(&c, 1, var, &host_out);
}

int fdtd_propagate(fdtd_ctx *c, const fdtd_dump_sink *sink, size_t *steps_done, double *time_counter)
{
    FDTD_TRY(check_solo(c, "fdtd_propagate"));
    FDTD_TRY(use_device(c));
    return propagate_many(&c, 1, sink, steps_done, time_counter);
}

} /* extern "C" */
