/*
 * fdtd_diag_kernels.cuh -- reductions and test support next to the path.  Included by fdtd_diag.cu only.
 */
#pragma once

#include "fdtd_types.cuh"

namespace fdtd {

/* ------------------------------------------------------------------------------------------
 * Diagnostics of the reference that sit next to the path.  They are
 * reductions: the reference accumulates sequentially, a GPU cannot, so these agree with the CPU to
 * rounding (about 1e-13 relative), not bit for bit.  Neither feeds back into the fields.
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void block_accumulate(double *vals, int n, double *out)
{
    __shared__ double warp_part[8][8];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int v = 0; v < n; ++v) {
        double x = vals[v];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((tid & 31) == 0)
            warp_part[tid >> 5][v] = x;
    }
    __syncthreads();
    if (tid < n) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x * blockDim.y + 31) / 32; ++w)
            t += warp_part[w][tid];
        atomicAdd(out + tid, t);
    }
}

/* calculate_E_energy / calculate_H_energy, main.c:602-668: sums over the zones of the squared
 * zone-averaged components; out[0..2] = ex, ey, ez, out[3..5] = hx, hy, hz (the host applies dv and
 * eps/2, mu/2).  as_coded != 0 reproduces main.c:627, which indexes Ez with Hz's strides. */
__global__ void __launch_bounds__(256) k_energy(Geo g, Fld f, int as_coded, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kc = blockIdx.z;
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (i < g.I && j < g.J) {
        const long long P = g.P, PR = g.PR;
        const long long o = i + P * (j + (long long)g.R * (kc + 1));
        const double mex = (f.ex[o] + f.ex[o + PR] + f.ex[o + P] + f.ex[o + P + PR]) / 4.;
        const double mey = (f.ey[o] + f.ey[o + 1] + f.ey[o + PR] + f.ey[o + 1 + PR]) / 4.;
        double mez;
        if (!as_coded) {
            mez = (f.ez[o] + f.ez[o + P] + f.ez[o + 1] + f.ez[o + 1 + P]) / 4.;
        } else {
            /* Ez[kHz(i,j,k)]: dense offset i + j*I + k*I*J reinterpreted in Ez's own dense shape */
            const long long I = g.I, J = g.J, k = kc + g.kbase;
            const long long m[4] = {i + j * I + k * I * J, i + (j + 1) * I + k * I * J,
                                    (i + 1) + j * I + k * I * J, (i + 1) + (j + 1) * I + k * I * J};
            double s = 0.0;
            for (int t = 0; t < 4; ++t) {
                const long long ii = m[t] % (I + 1), jj = (m[t] / (I + 1)) % (J + 1), kk = m[t] / ((I + 1) * (J + 1));
                s += f.ez[ii + P * (jj + (long long)g.R * (kk - g.kbase + 1))];
            }
            mez = s / 4.;
        }
        const double mhx = (f.hx[o] + f.hx[o + 1]) / 2.;
        const double mhy = (f.hy[o] + f.hy[o + P]) / 2.;
        const double mhz = (f.hz[o] + f.hz[o + PR]) / 2.;
        v[0] = mex * mex; v[1] = mey * mey; v[2] = mez * mez;
        v[3] = mhx * mhx; v[4] = mhy * mhy; v[5] = mhz * mhz;
    }
    block_accumulate(v, 6, out);
}

/* Relative L2 error against the analytic TE101 cavity mode (main.c:670-710; description.pdf eq. 2):
 * out = {sum (a-Ey)^2, sum a^2} for Ey, then Hx, then Hz.  The analytic factors are separable and
 * come from the host: sk/ck = sin/cos(pi k dx / height), si/ci = sin/cos(pi i dx / length),
 * amp = {cos(2 pi f t), sin(2 pi f t) / Z_te, -pi / (omega mu length) * sin(2 pi f t)}. */
__global__ void __launch_bounds__(256) k_validation_error(Geo g, Fld f, const double *__restrict__ sk,
                                                          const double *__restrict__ ck,
                                                          const double *__restrict__ si,
                                                          const double *__restrict__ ci, double a_ey,
                                                          double a_hx, double a_hz, int top, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = blockIdx.z + 1; /* planes 1 .. nk (+1 on the top slab for Ey, Hz) */
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (i <= g.I && j < g.J) {
        const long long o = i + (long long)g.P * (j + (long long)g.R * kl);
        const int k = kl - 1 + g.kbase;
        const bool cell = kl <= g.nk;
        if (cell || top) { /* Ey: k <= K, i <= I */
            const double a = a_ey * sk[k] * si[i];
            const double d = a - f.ey[o];
            v[0] = d * d; v[1] = a * a;
        }
        if (cell) { /* Hx: k < K, i <= I */
            const double a = a_hx * sk[k] * ci[i];
            const double d = a - f.hx[o];
            v[2] = d * d; v[3] = a * a;
        }
        if ((cell || top) && i < g.I) { /* Hz: k <= K, i < I */
            const double a = a_hz * ck[k] * si[i];
            const double d = a - f.hz[o];
            v[4] = d * d; v[5] = a * a;
        }
    }
    block_accumulate(v, 6, out);
}

/* ------------------------------------------------------------------------------------------
 * Test pattern and checksum: let full-size runs (1024^3 and up, where no host copy of the state
 * exists) start from non-trivial data and be compared between kernel variants, slab counts and
 * the CPU oracle.  Both are pure functions of the element's index in the reference's DENSE
 * array (main.c:379-407), so they do not depend on pitch, slab or launch shape.
 * ------------------------------------------------------------------------------------------ */
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* one array: w x h dense rows/columns, local planes [1, 1 + np) <-> dense planes [kd0, kd0 + np) */
__global__ void __launch_bounds__(256) k_fill_pattern(Geo g, double *__restrict__ a, DenseView v,
                                                      unsigned long long seed, int array)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int pl = blockIdx.z;
    if (i >= v.w || j >= v.h)
        return;
    const unsigned long long dense = (unsigned long long)i + (unsigned long long)v.w * ((unsigned long long)j + (unsigned long long)v.h * (unsigned long long)(v.kd0 + pl));
    const unsigned long long r = splitmix64(seed ^ ((unsigned long long)array << 58) ^ dense);
    const double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
    a[i + (long long)g.P * (j + (long long)g.R * (pl + 1))] = __dsub_rn(__dmul_rn(2.0, u), 1.0);
}

/* sum over the owned elements of splitmix64(bits(value) + dense index), modulo 2^64 */
__global__ void __launch_bounds__(256) k_checksum(Geo g, const double *__restrict__ a, DenseView v,
                                                  unsigned long long *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int pl = blockIdx.z;
    unsigned long long s = 0;
    if (i < v.w && j < v.h) {
        const unsigned long long dense = (unsigned long long)i + (unsigned long long)v.w * ((unsigned long long)j + (unsigned long long)v.h * (unsigned long long)(v.kd0 + pl));
        const double x = a[i + (long long)g.P * (j + (long long)g.R * (pl + 1))];
        s = splitmix64((unsigned long long)__double_as_longlong(x) + dense);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, d);
    __shared__ unsigned long long warp_sum[8];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0)
        warp_sum[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x * blockDim.y + 31) / 32; ++w)
            t += warp_sum[w];
        atomicAdd(out, t);
    }
}

} /* namespace fdtd */
