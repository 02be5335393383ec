/*
 * fdtd_diag.cu -- reductions next to the path (energy, analytic-mode error: SURVEY.md 8(f) ranks 2, 3)
 * and the test pattern / checksum used where no host copy of the state exists.
 */
#include "fdtd_ctx.hpp"
#include "fdtd_diag_kernels.cuh"

using namespace fdtdi;

namespace fdtdi {

/* fdtd_energy over the n slabs this thread drives (n = 1: one context); the slabs' sums add up */
int energy_many(fdtd_ctx *const *cs, int n, int as_coded, double *e_energy, double *h_energy)
{
    if (as_coded && cs[0]->nranks > 1) {
        /* main.c:627 reads Ez at Hz's offsets, i.e. from other planes -- possibly another slab's */
        fdtd_set_error("fdtd_energy: as_coded is only available on a single-GPU context");
        return FDTD_E_ARG;
    }
    /* the top zone plane of a slab averages with node plane k1 of Ex, Ey, Hz */
    FDTD_TRY(exchange_many_for_dump(cs, n));
    double tot[6] = {0, 0, 0, 0, 0, 0};
    std::vector<double *> dev(n, nullptr);
    std::vector<double> host(6 * (size_t)n, 0.0);
    int rc = FDTD_OK;
    for (int r = 0; r < n && rc == FDTD_OK; ++r) {
        fdtd_ctx *c = cs[r];
        rc = use_device(c);
        if (rc != FDTD_OK)
            break;
        if (cudaMalloc((void **)&dev[r], 6 * sizeof(double)) != cudaSuccess) {
            fdtd_set_error("fdtd_energy: cudaMalloc failed");
            rc = FDTD_E_CUDA;
            break;
        }
        cudaMemsetAsync(dev[r], 0, 6 * sizeof(double), c->s_main);
        dim3 block(64, 4);
        dim3 grid((c->g.I + 63) / 64, (c->g.J + 3) / 4, c->g.nk);
        k_energy<<<grid, block, 0, c->s_main>>>(c->g, c->f, as_coded, dev[r]);
        cudaMemcpyAsync(&host[6 * (size_t)r], dev[r], 6 * sizeof(double), cudaMemcpyDeviceToHost, c->s_main);
    }
    for (int r = 0; r < n; ++r) {
        if (!dev[r])
            continue;
        cudaSetDevice(cs[r]->device);
        if (cudaStreamSynchronize(cs[r]->s_main) != cudaSuccess && rc == FDTD_OK) {
            fdtd_set_error("fdtd_energy: %s", cudaGetErrorString(cudaGetLastError()));
            rc = FDTD_E_CUDA;
        }
        cudaFree(dev[r]);
        for (int v = 0; v < 6; ++v)
            tot[v] += host[6 * (size_t)r + v];
    }
    if (rc != FDTD_OK)
        return rc;
    const double dv = pow(cs[0]->p.spatial_step, 3); /* main.c:613 */
    if (e_energy) *e_energy = (tot[0] * dv + tot[1] * dv + tot[2] * dv) * FDTD_EPSILON / 2.; /* main.c:631 */
    if (h_energy) *h_energy = (tot[3] * dv + tot[4] * dv + tot[5] * dv) * FDTD_MU / 2.;      /* main.c:665 */
    return FDTD_OK;
}

} /* namespace fdtdi */

extern "C" {

int fdtd_energy(fdtd_ctx *c, int as_coded, double *e_energy, double *h_energy)
{
    FDTD_TRY(check_solo(c, "fdtd_energy"));
    return energy_many(&c, 1, as_coded, e_energy, h_energy);
}

int fdtd_validation_error(fdtd_ctx *c, double t, double sums[6], double rel_l2[3])
{
    FDTD_TRY(check_ctx(c, "fdtd_validation_error"));
    FDTD_TRY(use_device(c));
    const fdtd_params &p = c->p;
    const size_t nk = p.maxk + 2, ni = p.maxi + 2;
    std::vector<double> tab(2 * nk + 2 * ni);
    double *sk = tab.data(), *ck = sk + nk, *si = ck + nk, *ci = si + ni;
    for (size_t k = 0; k < p.maxk + 1; ++k) {
        sk[k] = sin(FDTD_PI * k * p.spatial_step / p.height);
        ck[k] = cos(FDTD_PI * k * p.spatial_step / p.height);
    }
    for (size_t i = 0; i < p.maxi + 1; ++i) {
        si[i] = sin(FDTD_PI * i * p.spatial_step / p.length);
        ci[i] = cos(FDTD_PI * i * p.spatial_step / p.length);
    }
    /* main.c:672-675 */
    const double f_mnl = 0.5 * FDTD_CELERITY * sqrt(pow(FDTD_PI / p.height, 2) + pow(FDTD_PI / p.length, 2)) / FDTD_PI;
    const double omega = 2.0 * FDTD_PI * f_mnl;
    const double z_te = (omega * FDTD_MU) / sqrt(pow(omega, 2) * FDTD_MU * FDTD_EPSILON - pow(FDTD_PI / p.length, 2));
    const double a_ey = cos(2 * FDTD_PI * f_mnl * t);
    const double a_hx = (1.0 / z_te) * sin(2 * FDTD_PI * f_mnl * t);
    const double a_hz = -FDTD_PI / (omega * FDTD_MU * p.length) * sin(2 * FDTD_PI * f_mnl * t);
    double *dev = nullptr;
    CUDA_TRY(cudaMalloc((void **)&dev, (tab.size() + 6) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(dev + 6, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->s_main);
    cudaMemsetAsync(dev, 0, 6 * sizeof(double), c->s_main);
    dim3 block(64, 4);
    dim3 grid((c->g.I + 1 + 63) / 64, (c->g.J + 3) / 4, c->g.nk + c->g.top);
    k_validation_error<<<grid, block, 0, c->s_main>>>(c->g, c->f, dev + 6, dev + 6 + nk, dev + 6 + 2 * nk,
                                                     dev + 6 + 2 * nk + ni, a_ey, a_hx, a_hz, c->g.top, dev);
    double s[6];
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s, dev, sizeof s, cudaMemcpyDeviceToHost, c->s_main);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->s_main);
    cudaFree(dev);
    if (e != cudaSuccess) {
        fdtd_set_error("fdtd_validation_error: %s", cudaGetErrorString(e));
        return FDTD_E_CUDA;
    }
    for (int v = 0; v < 6; ++v)
        if (sums) sums[v] = s[v];
    for (int v = 0; v < 3; ++v)
        if (rel_l2) rel_l2[v] = s[2 * v + 1] > 0.0 ? sqrt(s[2 * v] / s[2 * v + 1]) : 0.0;
    return FDTD_OK;
}

static DenseView dense_view(const fdtd_ctx *c, int idx)
{
    const DenseShape s = dense_shape(c->p, idx);
    DenseView v;
    v.w = (int)s.w;
    v.h = (int)s.h;
    v.np = c->g.nk + ((s.node_planes && c->g.top) ? 1 : 0);
    v.kd0 = (long long)c->k0;
    return v;
}

int fdtd_fill_test_pattern(fdtd_ctx *c, unsigned long long seed)
{
    FDTD_TRY(check_ctx(c, "fdtd_fill_test_pattern"));
    FDTD_TRY(use_device(c));
    FDTD_TRY(wait_halos(c)); /* nothing of the last exchange may still be in flight */
    CUDA_TRY(cudaMemsetAsync(c->base, 0, 6 * c->array_elems * sizeof(double), c->s_main));
    for (int a = 0; a < 6; ++a) {
        const DenseView v = dense_view(c, a);
        dim3 block(64, 4);
        dim3 grid((v.w + 63) / 64, (v.h + 3) / 4, v.np);
        k_fill_pattern<<<grid, block, 0, c->s_main>>>(c->g, field_ptr(c, a), v, seed, a);
    }
    CUDA_TRY(cudaGetLastError());
    c->e_halo_valid = c->h_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = (c->nranks == 1);
    return FDTD_OK;
}

int fdtd_checksum(fdtd_ctx *c, unsigned long long out[6])
{
    FDTD_TRY(check_ctx(c, "fdtd_checksum"));
    if (!out) {
        fdtd_set_error("fdtd_checksum: NULL argument");
        return FDTD_E_ARG;
    }
    FDTD_TRY(use_device(c));
    unsigned long long *dev = nullptr;
    CUDA_TRY(cudaMalloc((void **)&dev, 6 * sizeof(unsigned long long)));
    cudaMemsetAsync(dev, 0, 6 * sizeof(unsigned long long), c->s_main);
    for (int a = 0; a < 6; ++a) {
        const DenseView v = dense_view(c, a);
        dim3 block(64, 4);
        dim3 grid((v.w + 63) / 64, (v.h + 3) / 4, v.np);
        k_checksum<<<grid, block, 0, c->s_main>>>(c->g, field_ptr(c, a), v, dev + a);
    }
    cudaError_t e = cudaMemcpyAsync(out, dev, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->s_main);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->s_main);
    cudaFree(dev);
    if (e != cudaSuccess) {
        fdtd_set_error("fdtd_checksum: %s", cudaGetErrorString(e));
        return FDTD_E_CUDA;
    }
    return FDTD_OK;
}
} /* extern "C" */
