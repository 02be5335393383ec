/*
 * fdtd_dump_kernels.cuh -- the dump variables of write_silo (main.c:550-598).  Included by fdtd_dump.cu only.
 */
#pragma once

#include "fdtd_types.cuh"

namespace fdtd {

/* ------------------------------------------------------------------------------------------
 * Dump variables: aggregate_E_field / aggregate_H_field, main.c:511-540, with the offset triples
 * of main.c:563-578.  One thread per zone; out is dense I x J x nk, x fastest.
 * The E average is kept exactly as coded -- (0,0,0) + (oi,oj,ok) + (0,oj,ok) + (oi,0,ok), summed
 * left to right, times .25 -- which for ex and ey counts one corner twice.
 * ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_aggregate(Geo g, const double *__restrict__ a, int var,
                                                   double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kc = blockIdx.z;
    if (i >= g.I || j >= g.J)
        return;
    const long long oi = (var == 1 || var == 2 || var == 3) ? 1 : 0;
    const long long oj = (var == 0 || var == 2 || var == 4) ? g.P : 0;
    const long long ok = (var == 0 || var == 1 || var == 5) ? g.PR : 0;
    const long long o = i + (long long)g.P * (j + (long long)g.R * (kc + 1));
    double v;
    if (var < 3) {
        v = __dadd_rn(a[o], a[o + oi + oj + ok]);
        v = __dadd_rn(v, a[o + oj + ok]);
        v = __dadd_rn(v, a[o + oi + ok]);
        v = __dmul_rn(.25, v);
    } else {
        v = __dmul_rn(.5, __dadd_rn(a[o], a[o + oi + oj + ok]));
    }
    out[i + (long long)g.I * (j + (long long)g.J * kc)] = v;
}

/* Validation-mode dump variable aEy (main.c:583): the zone average, with ey's offsets, of
 * "analytic TE101 Ey minus computed Ey" (main.c:688-691).  The analytic factor is separable;
 * ct = cos(2 pi f t), sk[k] = sin(pi k dx / height), si[i] = sin(pi i dx / length) are evaluated on
 * the host with glibc, so the device only multiplies and subtracts: (ct * sk[k]) * si[i] - Ey. */
__global__ void __launch_bounds__(256) k_aggregate_aey(Geo g, const double *__restrict__ ey, double ct,
                                                       const double *__restrict__ sk,
                                                       const double *__restrict__ si,
                                                       double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kc = blockIdx.z;
    if (i >= g.I || j >= g.J)
        return;
    const long long o = i + (long long)g.P * (j + (long long)g.R * (kc + 1));
    const int k = kc + g.kbase;
    const double a0 = __dmul_rn(ct, sk[k]), a1 = __dmul_rn(ct, sk[k + 1]);
    const double v000 = __dsub_rn(__dmul_rn(a0, si[i]), ey[o]);
    const double v101 = __dsub_rn(__dmul_rn(a1, si[i + 1]), ey[o + 1 + g.PR]);
    const double v001 = __dsub_rn(__dmul_rn(a1, si[i]), ey[o + g.PR]);
    double v = __dadd_rn(v000, v101);
    v = __dadd_rn(v, v001);
    v = __dadd_rn(v, v101);
    out[i + (long long)g.I * (j + (long long)g.J * kc)] = __dmul_rn(.25, v);
}

} /* namespace fdtd */
