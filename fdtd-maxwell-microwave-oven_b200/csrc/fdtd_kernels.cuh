/*
 * fdtd_kernels.cuh -- sm_100a kernels of the FDTD hot path.
 *
 * Device layout ("pitched slab").  All six arrays share one geometry so that one 64-bit offset
 * addresses the same (i, j, k) in every array:
 *
 *     offset(i, j, kl) = i + P * (j + R * kl)          doubles
 *
 *   P  = pitch, (I + 1) rounded up to 16 doubles: every row starts on a 128-byte line, which the
 *        reference's dense rows (I or I+1 doubles, main.c:379-407) do not;
 *   R  = J + 1 rows per plane for every array;
 *   kl = local plane index.  A slab that owns the cell planes [k0, k1) stores global plane k at
 *        kl = k - k0 + 1.  Plane 0 receives the lower neighbour's Hx/Hy halo, plane nk + 1 the
 *        upper neighbour's Ex/Ey halo -- or, on the last slab, it IS the global node plane K.
 *   Padding (columns beyond an array's extent, the unused row/plane) is zero and never stored to.
 *
 * Arithmetic.  Every update is evaluated with explicit round-to-nearest intrinsics in the
 * reference's operand order (main.c:448-461, 486-499):
 *     F = F + c * ((a - b) - (d - e))
 * so no FMA can be formed whatever the compiler flags; the build also passes -fmad=false.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fdtd {

struct Geo {
    int I, J, K;      /* global cell counts maxi, maxj, maxk */
    int P, R;         /* pitch in doubles, rows per plane */
    long long PR;     /* plane stride in doubles */
    int nk;           /* cell planes owned by this slab */
    int kbase;        /* global k of local plane 1 */
    int top;          /* 1: local plane nk+1 is the global node plane K */
    int planes;       /* nk + 2 */
};

struct Fld {
    double *ex, *ey, *ez, *hx, *hy, *hz;
};

/* Waveguide source (main.c:712-753) in fused form.  vals[0..n) are the Ez amplitudes and
 * vals[n..2n) the Hx amplitudes of this step, computed on the host with glibc (fdtd_source_values).
 * on == 0 in validation mode and on slabs that do not hold the global plane k = 0. */
struct Src {
    int on;
    int i0, i1, j0, j1;
    int n;
    const double *vals;
};

__device__ __forceinline__ double yee(double f, double c, double a, double b, double d, double e)
{
    return __dadd_rn(f, __dmul_rn(c, __dsub_rn(__dsub_rn(a, b), __dsub_rn(d, e))));
}

__device__ __forceinline__ bool in_patch(const Src &s, int i, int j)
{
    return i >= s.i0 && i < s.i1 && j >= s.j0 && j < s.j1;
}

__device__ __forceinline__ double ldp(const double *__restrict__ p, long long off, bool ok)
{
    return ok ? __ldg(p + off) : 0.0;
}

/* ------------------------------------------------------------------------------------------
 * Variant 0: one thread per cell.  Plain operators, no source fusion: the host launches
 * k_set_source around them exactly where the reference calls set_source (main.c:770-778).
 * ------------------------------------------------------------------------------------------ */

/* update_H_field, main.c:431-462.  grid.z walks the local planes 1 .. nk + top. */
__global__ void __launch_bounds__(256) k_update_h_cell(Geo g, Fld f, double c)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = blockIdx.z + 1;
    if (i > g.I || j > g.J)
        return;
    const long long o = i + (long long)g.P * (j + (long long)g.R * kl);
    const bool cell = kl <= g.nk;
    if (cell && j < g.J) /* Hx: k < K, j < J, i <= I */
        f.hx[o] = yee(f.hx[o], c, f.ey[o + g.PR], f.ey[o], f.ez[o + g.P], f.ez[o]);
    if (cell && i < g.I) /* Hy: k < K, j <= J, i < I */
        f.hy[o] = yee(f.hy[o], c, f.ez[o + 1], f.ez[o], f.ex[o + g.PR], f.ex[o]);
    if (i < g.I && j < g.J) /* Hz: k <= K, j < J, i < I */
        f.hz[o] = yee(f.hz[o], c, f.ex[o + g.P], f.ex[o], f.ey[o + 1], f.ey[o]);
}

/* update_E_field, main.c:469-500; the skipped faces are the PEC wall.  grid.z: planes 1 .. nk. */
__global__ void __launch_bounds__(256) k_update_e_cell(Geo g, Fld f, double c)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = blockIdx.z + 1;
    if (i > g.I || j > g.J)
        return;
    const long long o = i + (long long)g.P * (j + (long long)g.R * kl);
    const bool kin = (kl - 1 + g.kbase) >= 1; /* global k >= 1; k < K holds for every owned plane */
    if (kin && j >= 1 && j < g.J && i < g.I) /* Ex */
        f.ex[o] = yee(f.ex[o], c, f.hz[o], f.hz[o - g.P], f.hy[o], f.hy[o - g.PR]);
    if (kin && j < g.J && i >= 1 && i < g.I) /* Ey */
        f.ey[o] = yee(f.ey[o], c, f.hx[o], f.hx[o - g.PR], f.hz[o], f.hz[o - 1]);
    if (j >= 1 && j < g.J && i >= 1 && i < g.I) /* Ez: k < K */
        f.ez[o] = yee(f.ez[o], c, f.hy[o], f.hy[o - 1], f.hx[o], f.hx[o - g.P]);
}

/* set_source, main.c:745-752: one thread per patch point, plane kl = 1 (global k = 0). */
__global__ void k_set_source(Geo g, Fld f, Src s)
{
    const int i = s.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = s.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= s.i1 || j >= s.j1)
        return;
    const long long o = i + (long long)g.P * (j + (long long)g.R);
    f.ez[o] = s.vals[i - s.i0];
    f.ex[o] = 0.0;
    f.hz[o] = 0.0;
    f.hx[o] = s.vals[s.n + i - s.i0];
}

/* ------------------------------------------------------------------------------------------
 * Variant 1: z-marching register strips, source and PEC fused.
 *
 * A thread owns the column (i, jb .. jb+TY-1) and walks a chunk of planes upwards.  A warp is 32
 * consecutive i, so every row access is one fully coalesced 256-byte request.  What makes each
 * element come from HBM once per half-step:
 *   - the k+-1 neighbour is the value the same thread loaded one plane ago (registers);
 *   - the j+-1 neighbour is the next row of the same thread's strip (registers; one extra row per
 *     strip comes from L1/L2);
 *   - the i+-1 neighbour comes from the adjacent lane by shuffle (the edge lane re-reads one
 *     element that the neighbouring warp has just pulled into L1/L2).
 * Block = (32*WX) x WY threads = 32*WX columns x WY*TY rows; grid.z = plane chunks.
 * ------------------------------------------------------------------------------------------ */

struct Span {
    int kl_begin, kl_end; /* local planes [kl_begin, kl_end) handled by this launch */
    int kchunk;           /* planes per block */
    int prefetch;         /* planes ahead to pull into L2 (0 = off) */
};

/* non-blocking hint: bring the line holding p into L2 (no register, no scoreboard entry) */
__device__ __forceinline__ void prefetch_l2(const double *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

/* resident blocks per SM the register budget is capped for: 256 threads x 2 blocks x 128 registers
 * fill the register file for TY = 4; shorter strips need fewer registers and fit more blocks */
template <int TY>
struct MarchCfg {
    static constexpr int kMinBlocks = TY >= 4 ? 2 : (TY == 2 ? 3 : 4);
};

template <int TY>
__global__ void __launch_bounds__(256, MarchCfg<TY>::kMinBlocks) k_update_h_march(Geo g, Fld f, double c, Src s, Span sp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int jb = (blockIdx.y * blockDim.y + threadIdx.y) * TY;
    const int kl0 = sp.kl_begin + blockIdx.z * sp.kchunk;
    const int kl1 = min(kl0 + sp.kchunk, sp.kl_end);
    if (jb > g.J)
        return; /* warp-uniform: threadIdx.y is constant inside a warp */
    const unsigned lane = threadIdx.x & 31u;
    const bool xn = i <= g.I; /* column exists in Ey, Ez, Hx */
    const bool xc = i < g.I;  /* column exists in Ex, Hy, Hz */
    const bool xn1 = i + 1 <= g.I;

    bool rn[TY + 1], rc[TY + 1]; /* row exists in (Ex, Ez, Hy) / (Ey, Hx, Hz) */
#pragma unroll
    for (int r = 0; r <= TY; ++r) {
        rn[r] = jb + r <= g.J;
        rc[r] = jb + r < g.J;
    }

    long long o = i + (long long)g.P * (jb + (long long)g.R * kl0);
    const bool src_chunk = s.on && kl0 == 1; /* this chunk starts on the global plane k = 0 */

    /* plane kl0 of Ex (TY+1 rows) and Ey (TY rows) */
    double exk[TY + 1], eyk[TY];
#pragma unroll
    for (int r = 0; r <= TY; ++r)
        exk[r] = ldp(f.ex, o + (long long)r * g.P, xc && rn[r]);
#pragma unroll
    for (int r = 0; r < TY; ++r)
        eyk[r] = ldp(f.ey, o + (long long)r * g.P, xn && rc[r]);
    if (src_chunk) { /* first set_source of the step: Ex = 0 on the patch (main.c:749) */
#pragma unroll
        for (int r = 0; r <= TY; ++r)
            if (in_patch(s, i, jb + r))
                exk[r] = 0.0;
    }

    for (int kl = kl0; kl < kl1; ++kl, o += g.PR) {
        const bool cell = kl <= g.nk;           /* Hx, Hy, Ez exist on this plane */
        const bool srck = s.on && kl == 1;      /* global plane k = 0 carries the source */

        double exn[TY + 1], eyn[TY], ezk[TY + 1], hx[TY], hy[TY], hz[TY];
#pragma unroll
        for (int r = 0; r <= TY; ++r) {
            exn[r] = ldp(f.ex, o + g.PR + (long long)r * g.P, cell && xc && rn[r]);
            ezk[r] = ldp(f.ez, o + (long long)r * g.P, cell && xn && rn[r]);
        }
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            eyn[r] = ldp(f.ey, o + g.PR + (long long)r * g.P, cell && xn && rc[r]);
            hx[r] = (cell && xn && rc[r]) ? f.hx[o + (long long)r * g.P] : 0.0;
            hy[r] = (cell && xc && rn[r]) ? f.hy[o + (long long)r * g.P] : 0.0;
            hz[r] = (xc && rc[r]) ? f.hz[o + (long long)r * g.P] : 0.0;
        }
        if (srck) { /* first set_source of the step: Ez on the patch (main.c:748) */
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                if (in_patch(s, i, jb + r))
                    ezk[r] = s.vals[i - s.i0];
        }

        /* i+1 neighbours of Ez and Ey on plane k: next lane, or a direct read on the warp edge */
        double ezi[TY], eyi[TY];
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            ezi[r] = __shfl_down_sync(0xffffffffu, ezk[r], 1);
            eyi[r] = __shfl_down_sync(0xffffffffu, eyk[r], 1);
        }
        if (lane == 31u) {
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                ezi[r] = ldp(f.ez, o + 1 + (long long)r * g.P, cell && xn1 && rn[r]);
                eyi[r] = ldp(f.ey, o + 1 + (long long)r * g.P, xn1 && rc[r]);
                if (srck && in_patch(s, i + 1, jb + r))
                    ezi[r] = s.vals[i + 1 - s.i0];
            }
        }

#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const long long orow = o + (long long)r * g.P;
            const bool patch = srck && in_patch(s, i, jb + r);
            if (cell && xn && rc[r]) { /* Hx, main.c:448 */
                double v = yee(hx[r], c, eyn[r], eyk[r], ezk[r + 1], ezk[r]);
                if (patch)
                    v = s.vals[s.n + i - s.i0]; /* second set_source overwrites it, main.c:751 */
                f.hx[orow] = v;
            }
            if (cell && xc && rn[r]) /* Hy, main.c:454 */
                f.hy[orow] = yee(hy[r], c, ezi[r], ezk[r], exn[r], exk[r]);
            if (xc && rc[r]) { /* Hz, main.c:460 */
                double v = yee(hz[r], c, exk[r + 1], exk[r], eyi[r], eyk[r]);
                if (patch)
                    v = 0.0; /* main.c:750 */
                f.hz[orow] = v;
            }
        }
#pragma unroll
        for (int r = 0; r <= TY; ++r)
            exk[r] = exn[r];
#pragma unroll
        for (int r = 0; r < TY; ++r)
            eyk[r] = eyn[r];
    }
}

template <int TY>
__global__ void __launch_bounds__(256, MarchCfg<TY>::kMinBlocks) k_update_e_march(Geo g, Fld f, double c, Src s, Span sp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int jb = (blockIdx.y * blockDim.y + threadIdx.y) * TY;
    const int kl0 = sp.kl_begin + blockIdx.z * sp.kchunk;
    const int kl1 = min(kl0 + sp.kchunk, sp.kl_end);
    if (jb > g.J)
        return;
    const unsigned lane = threadIdx.x & 31u;
    const bool xn = i <= g.I;
    const bool xc = i < g.I;
    const bool xi = i >= 1 && i < g.I; /* interior column: Ey, Ez are updated here */
    const bool xm = i >= 1;            /* column i-1 exists */

    /* rows jb-1 .. jb+TY-1, index r+1 */
    bool rn[TY + 1], rc[TY + 1], rint[TY + 1];
#pragma unroll
    for (int r = 0; r <= TY; ++r) {
        const int j = jb - 1 + r;
        rn[r] = j >= 0 && j <= g.J;
        rc[r] = j >= 0 && j < g.J;
        rint[r] = j >= 1 && j < g.J; /* interior row: Ex, Ez are updated here */
    }

    long long o = i + (long long)g.P * (jb + (long long)g.R * kl0);

    /* plane kl0 - 1 of Hx, Hy.  Local plane 0 is the lower halo; on the slab that starts at the
     * global bottom it holds nothing and nothing reads it (k = 0 is PEC for Ex, Ey). */
    double hxm[TY], hym[TY];
    const bool below = kl0 >= 2 || g.kbase > 0;
#pragma unroll
    for (int r = 0; r < TY; ++r) {
        hxm[r] = ldp(f.hx, o - g.PR + (long long)r * g.P, below && xn && rc[r + 1]);
        hym[r] = ldp(f.hy, o - g.PR + (long long)r * g.P, below && xc && rn[r + 1]);
    }

    for (int kl = kl0; kl < kl1; ++kl, o += g.PR) {
        const bool kin = (kl - 1 + g.kbase) >= 1; /* Ex, Ey are updated on this plane */
        const bool srck = s.on && kl == 1;

        double hxk[TY + 1], hzk[TY + 1], hyk[TY], ex[TY], ey[TY], ez[TY];
#pragma unroll
        for (int r = 0; r <= TY; ++r) {
            const long long orow = o + (long long)(r - 1) * g.P;
            hxk[r] = ldp(f.hx, orow, xn && rc[r]);
            hzk[r] = ldp(f.hz, orow, xc && rc[r]);
        }
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const long long orow = o + (long long)r * g.P;
            hyk[r] = ldp(f.hy, orow, xc && rn[r + 1]);
            ex[r] = (kin && xc && rint[r + 1]) ? f.ex[orow] : 0.0;
            ey[r] = (kin && xi && rc[r + 1]) ? f.ey[orow] : 0.0;
            ez[r] = (xi && rint[r + 1]) ? f.ez[orow] : 0.0;
        }

        /* i-1 neighbours of Hy and Hz on plane k */
        double hyi[TY], hzi[TY];
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            hyi[r] = __shfl_up_sync(0xffffffffu, hyk[r], 1);
            hzi[r] = __shfl_up_sync(0xffffffffu, hzk[r + 1], 1);
        }
        if (lane == 0u) {
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                const long long orow = o - 1 + (long long)r * g.P;
                hyi[r] = ldp(f.hy, orow, xm && i - 1 < g.I && rn[r + 1]);
                hzi[r] = ldp(f.hz, orow, xm && i - 1 < g.I && rc[r + 1]);
            }
        }

#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const long long orow = o + (long long)r * g.P;
            const bool patch = srck && in_patch(s, i, jb + r);
            if (kin && xc && rint[r + 1]) /* Ex, main.c:486 */
                f.ex[orow] = yee(ex[r], c, hzk[r + 1], hzk[r], hyk[r], hym[r]);
            else if (patch && xc)
                f.ex[orow] = 0.0; /* set_source left Ex = 0 on the patch (main.c:749) */
            if (kin && xi && rc[r + 1]) /* Ey, main.c:492 */
                f.ey[orow] = yee(ey[r], c, hxk[r + 1], hxm[r], hzk[r + 1], hzi[r]);
            {   /* Ez, main.c:498.  On the patch the old value is the source amplitude that the
                   second set_source of this step wrote (main.c:748). */
                const bool upd = xi && rint[r + 1];
                double old = ez[r];
                if (patch)
                    old = s.vals[i - s.i0];
                if (upd)
                    f.ez[orow] = yee(old, c, hyk[r], hyi[r], hxk[r + 1], hxk[r]);
                else if (patch && xn)
                    f.ez[orow] = old;
            }
        }
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            hxm[r] = hxk[r + 1];
            hym[r] = hyk[r];
        }
    }
}

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
struct DenseView {
    int w, h, np;
    long long kd0;
};

/* value = 2 * u - 1, u = top 53 bits of splitmix64(seed ^ array<<58 ^ dense index) / 2^53 */
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
