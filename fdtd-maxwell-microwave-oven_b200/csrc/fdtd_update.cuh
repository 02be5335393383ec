/*
 * fdtd_update.cuh -- the split update kernels: one launch per half-step, in place (144 B per
 * cell-update).  Variant 0 mirrors the reference's three operators one to one; variant 1 is the
 * z-marching register-strip form with source and PEC fused.  Layout and arithmetic: fdtd_types.cuh.
 * Included by fdtd_ctx.cu only.
 */
#pragma once

#include "fdtd_types.cuh"

namespace fdtd {

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
    const bool src_chunk = s.on && kl0 == s.kl; /* this chunk starts on the global plane k = 0 */

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
        const bool srck = s.on && kl == s.kl;      /* global plane k = 0 carries the source */

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
        const bool srck = s.on && kl == s.kl;

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

} /* namespace fdtd */
