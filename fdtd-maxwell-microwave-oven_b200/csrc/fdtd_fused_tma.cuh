/*
 * fdtd_fused_tma.cuh -- the fused single-sweep step (fdtd_fused.cuh) with its operands staged
 * through shared memory by the Tensor Memory Accelerator.
 *
 * Why.  k_step_fused reads its operands with ordinary loads; at 128 registers per thread only 16
 * warps fit on an SM and ncu shows them parked on the long scoreboard (63 % of the issue latency):
 * not enough loads in flight to cover HBM latency.  Here one elected thread per block issues, a
 * few planes ahead of the sweep, six cp.async.bulk.tensor loads per plane -- the block's tile of
 * Ex, Ey, Ez, Hx, Hy, Hz plus a one-element halo ring -- into a ring of shared-memory stages, each
 * guarded by an mbarrier that counts the landed bytes.  The bytes in flight are then set by the
 * ring depth (stages x 6 boxes), not by registers, and the compute warps only ever wait on shared
 * memory.  TMA zero-fills whatever part of a box lies outside the tensor, so halo cells beyond a
 * wall need no special case and no address is ever out of bounds.
 *
 * Everything else is the fused step: same register strips (thread = column i, rows jb..jb+TY-1),
 * same recomputation of the H ring (row jb-1 by the thread, column i-1 by lane 0 of each warp),
 * same double-buffered state (reads `a` through the tensor maps, writes `b` with plain coalesced
 * stores), same un-fused arithmetic, same fused source / PEC semantics.
 *
 * Stage s of the ring holds, for one plane p, the six boxes
 *     [bx0-2, bx0+BX+1] x [by0-1, by0+BY]   (W = BX+4 columns, HH = BY+2 rows, x fastest)
 * -- two halo columns either side although one is needed: the byte address of a box's first element
 * must be 16-byte aligned, so with 8-byte elements the innermost start coordinate has to be even
 * (an odd one raises "illegal instruction"; measured with tools/probe/tma_probe.cu).
 * Iteration p of the sweep reads Ez, Hx, Hy, Hz of plane p and Ex, Ey of plane p+1 (Ex, Ey of plane p
 * are still in registers from the previous iteration), so a stage is dead after iteration p and is
 * refilled, after the block barrier that ends the iteration, with plane p + STAGES.
 */
#pragma once

#include "fdtd_types.cuh"

namespace fdtd {

namespace tma {

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

/* one box of a 3-D tensor (x, y, plane) into shared memory; completion is counted on `bar` */
__device__ __forceinline__ void load_box(void *dst, const CUtensorMap *map, int x, int y, int z,
                                         unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

} /* namespace tma */


/* CWX, CWY: block shape in warps (32*CWX x CWY threads) known at compile time, so that every
 * shared-memory offset is an immediate; 0 = read the shape from blockDim (any launch shape). */
template <int TY, bool EDGE, int CWX, int CWY>
__device__ __forceinline__ void fused_tma_sweep(const Geo &g, const TmaMaps &maps, const Fld &b,
                                                const double cH, const double cE, const Src &s,
                                                const int stages, double *ring, unsigned long long *full,
                                                const int bx0, const int by0, const int kl0, const int kl1)
{
    constexpr int NR = TY + 2;
    const int BX = CWX ? 32 * CWX : (int)blockDim.x, BY = (CWY ? CWY : (int)blockDim.y) * TY;
    const int W = BX + 4;
    const int box = tma_box_doubles(BX, BY);
    const int stage_doubles = 6 * box;
    const unsigned lane = threadIdx.x & 31u;
    const int i = bx0 + threadIdx.x;
    const int jb = by0 + threadIdx.y * TY;
    const int P = g.P;
    const bool leader = threadIdx.x == 0 && threadIdx.y == 0;

    bool st_nc[TY], st_cn[TY], st_cc[TY], st_nn[TY], up_x[TY], up_y[TY], up_z[TY];
    if (EDGE) {
        const bool xn = i <= g.I, xc = i < g.I, xi = i >= 1 && i < g.I;
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const int j = jb + r;
            const bool jn = j <= g.J, jc = j < g.J, ji = j >= 1 && j < g.J;
            st_nc[r] = xn && jc;
            st_cn[r] = xc && jn;
            st_cc[r] = xc && jc;
            st_nn[r] = xn && jn;
            up_x[r] = xc && ji;
            up_y[r] = xi && jc;
            up_z[r] = xi && ji;
        }
    }

    const bool below = (kl0 - 1 + g.kbase) >= 1;
    const int kstart = below ? kl0 - 1 : kl0;
    const int nplanes = kl1 - kstart + 1; /* the last one only for its Ex, Ey */
    const unsigned full_bytes = 6u * (unsigned)(W * (BY + 2)) * 8u;
    const unsigned last_bytes = 2u * (unsigned)(W * (BY + 2)) * 8u;

    /* producer: plane kstart + n into ring slot `slot` (= n mod stages, tracked without dividing) */
    auto issue = [&](int slot, int n) {
        double *dst = ring + (size_t)slot * stage_doubles;
        unsigned long long *bar = full + slot;
        const int plane = kstart + n;
        const bool last = n == nplanes - 1;
        tma::mbar_expect_tx(bar, last ? last_bytes : full_bytes);
        tma::load_box(dst + 0 * box, &maps.m[0], bx0 - 2, by0 - 1, plane, bar);
        tma::load_box(dst + 1 * box, &maps.m[1], bx0 - 2, by0 - 1, plane, bar);
        if (!last) {
            tma::load_box(dst + 2 * box, &maps.m[2], bx0 - 2, by0 - 1, plane, bar);
            tma::load_box(dst + 3 * box, &maps.m[3], bx0 - 2, by0 - 1, plane, bar);
            tma::load_box(dst + 4 * box, &maps.m[4], bx0 - 2, by0 - 1, plane, bar);
            tma::load_box(dst + 5 * box, &maps.m[5], bx0 - 2, by0 - 1, plane, bar);
        }
    };
    if (leader) {
        for (int n = 0; n < stages && n < nplanes; ++n)
            issue(n, n);
    }

    /* position of (i, jb-1) inside a box: column tx+2, row ty*TY */
    const int c0 = (int)threadIdx.x + 2 + W * ((int)threadIdx.y * TY);
    const int q = i + P * jb;

    /* E of plane kstart */
    double exk[NR], eyk[NR - 1], exl[TY + 1], eyl[TY];
    tma::mbar_wait(full + 0, 0);
    {
        const double *sex = ring, *sey = ring + box;
#pragma unroll
        for (int rr = 0; rr < NR; ++rr)
            exk[rr] = sex[c0 + rr * W];
#pragma unroll
        for (int rr = 0; rr < NR - 1; ++rr)
            eyk[rr] = sey[c0 + rr * W];
        if (lane == 0u) {
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                exl[r] = sex[c0 - 1 + (r + 1) * W];
#pragma unroll
            for (int r = 0; r < TY; ++r)
                eyl[r] = sey[c0 - 1 + (r + 1) * W];
        }
        if (s.on && kstart == s.kl) {
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
                if (in_patch(s, i, jb - 1 + rr))
                    exk[rr] = 0.0;
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                if (in_patch(s, i - 1, jb + r))
                    exl[r] = 0.0;
        }
    }
    double hxm[TY], hym[TY];
#pragma unroll
    for (int r = 0; r < TY; ++r)
        hxm[r] = hym[r] = 0.0;

    /* A warp whose 32 columns all lie beyond column I (the last block in x usually holds the single
     * column i = I) has nothing to compute: it only keeps the block barriers company. */
    const bool warp_has_work = !EDGE || (i - (int)lane) <= g.I;

    long long pl = (long long)kstart * g.PR;
    int slot = 0;        /* ring slot of plane kl */
    unsigned phase = 0;  /* mbarrier phase parity of that slot's current fill */
    for (int kl = kstart, n = 0; kl < kl1; ++kl, ++n, pl += g.PR) {
        const bool srck = s.on && kl == s.kl;
        int slot1 = slot + 1;
        unsigned phase1 = phase;
        if (slot1 == stages) {
            slot1 = 0;
            phase1 ^= 1u;
        }
        const int slot_now = slot;
        slot = slot1;
        phase = phase1;
        if (!warp_has_work) {
            __syncthreads();
            if (leader && n + stages < nplanes)
                issue(slot_now, n + stages);
            continue;
        }
        /* stage n was awaited one iteration ago (for its Ex, Ey) or above (n = 0) */
        tma::mbar_wait(full + slot1, phase1);
        const double *cur = ring + (size_t)slot_now * stage_doubles;
        const double *nxt = ring + (size_t)slot1 * stage_doubles;
        const double *sexn = nxt, *seyn = nxt + box;
        const double *sez = cur + 2 * box, *shx = cur + 3 * box, *shy = cur + 4 * box, *shz = cur + 5 * box;

        double exn[NR], eyn[NR - 1], ezk[NR], hxo[TY + 1], hyo[TY], hzo[TY + 1];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
            exn[rr] = sexn[c0 + rr * W];
            ezk[rr] = sez[c0 + rr * W];
        }
#pragma unroll
        for (int rr = 0; rr < NR - 1; ++rr) {
            eyn[rr] = seyn[c0 + rr * W];
            hxo[rr] = shx[c0 + rr * W];
            hzo[rr] = shz[c0 + rr * W];
        }
#pragma unroll
        for (int r = 0; r < TY; ++r)
            hyo[r] = shy[c0 + (r + 1) * W];

        double ezi[TY], eyi[TY + 1];
#pragma unroll
        for (int r = 0; r < TY; ++r)
            ezi[r] = __shfl_down_sync(0xffffffffu, ezk[r + 1], 1);
#pragma unroll
        for (int rr = 0; rr <= TY; ++rr)
            eyi[rr] = __shfl_down_sync(0xffffffffu, eyk[rr], 1);
        if (lane == 31u) {
            /* Ey of plane kl at column i+1 lives in the stage of plane kl: its Ey box */
            const double *seyk = cur + box;
#pragma unroll
            for (int r = 0; r < TY; ++r)
                ezi[r] = sez[c0 + 1 + (r + 1) * W];
#pragma unroll
            for (int rr = 0; rr <= TY; ++rr)
                eyi[rr] = seyk[c0 + 1 + rr * W];
        }

        double hyln[TY], hzln[TY];
        if (lane == 0u) {
            double exln[TY + 1], eyln[TY], ezl[TY];
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                exln[r] = sexn[c0 - 1 + (r + 1) * W];
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                eyln[r] = seyn[c0 - 1 + (r + 1) * W];
                ezl[r] = sez[c0 - 1 + (r + 1) * W];
            }
            if (srck) {
#pragma unroll
                for (int r = 0; r < TY; ++r)
                    if (in_patch(s, i - 1, jb + r))
                        ezl[r] = s.vals[i - 1 - s.i0];
            }
            double ezo[TY];
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                ezo[r] = ezk[r + 1];
                if (srck && in_patch(s, i, jb + r))
                    ezo[r] = s.vals[i - s.i0];
            }
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                hyln[r] = yee(shy[c0 - 1 + (r + 1) * W], cH, ezo[r], ezl[r], exln[r], exl[r]);
                hzln[r] = yee(shz[c0 - 1 + (r + 1) * W], cH, exl[r + 1], exl[r], eyk[r + 1], eyl[r]);
                if (srck && in_patch(s, i - 1, jb + r))
                    hzln[r] = 0.0;
            }
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                exl[r] = exln[r];
#pragma unroll
            for (int r = 0; r < TY; ++r)
                eyl[r] = eyln[r];
        }

        if (srck) {
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
                if (in_patch(s, i, jb - 1 + rr))
                    ezk[rr] = s.vals[i - s.i0];
#pragma unroll
            for (int r = 0; r < TY; ++r)
                if (in_patch(s, i + 1, jb + r))
                    ezi[r] = s.vals[i + 1 - s.i0];
        }

        double hxn[TY + 1], hzn[TY + 1], hyn[TY];
#pragma unroll
        for (int rr = 0; rr <= TY; ++rr) {
            hxn[rr] = yee(hxo[rr], cH, eyn[rr], eyk[rr], ezk[rr + 1], ezk[rr]);
            hzn[rr] = yee(hzo[rr], cH, exk[rr + 1], exk[rr], eyi[rr], eyk[rr]);
        }
#pragma unroll
        for (int r = 0; r < TY; ++r)
            hyn[r] = yee(hyo[r], cH, ezi[r], ezk[r + 1], exn[r + 1], exk[r + 1]);
        if (srck) {
#pragma unroll
            for (int rr = 0; rr <= TY; ++rr)
                if (in_patch(s, i, jb - 1 + rr)) {
                    hxn[rr] = s.vals[s.n + i - s.i0];
                    hzn[rr] = 0.0;
                }
        }

        /* every thread has taken what it needs from stage n: hand the slot back to the producer */
        __syncthreads();
        if (leader && n + stages < nplanes)
            issue(slot_now, n + stages);

        if (kl >= kl0) {
            const bool cell = kl <= g.nk;
            const bool kin = (kl - 1 + g.kbase) >= 1 && cell;
            double *qex = b.ex + pl, *qey = b.ey + pl, *qez = b.ez + pl;
            double *qhx = b.hx + pl, *qhy = b.hy + pl, *qhz = b.hz + pl;
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                const int orow = q + r * P;
                double hyim = __shfl_up_sync(0xffffffffu, hyn[r], 1);
                double hzim = __shfl_up_sync(0xffffffffu, hzn[r + 1], 1);
                if (lane == 0u) {
                    hyim = hyln[r];
                    hzim = hzln[r];
                }
                double vex = exk[r + 1], vey = eyk[r + 1], vez = ezk[r + 1];
                if (kin) {
                    const double ux = yee(vex, cE, hzn[r + 1], hzn[r], hyn[r], hym[r]);
                    const double uy = yee(vey, cE, hxn[r + 1], hxm[r], hzn[r + 1], hzim);
                    vex = (!EDGE || up_x[r]) ? ux : vex;
                    vey = (!EDGE || up_y[r]) ? uy : vey;
                }
                if (cell) {
                    const double uz = yee(vez, cE, hyn[r], hyim, hxn[r + 1], hxn[r]);
                    vez = (!EDGE || up_z[r]) ? uz : vez;
                    if (!EDGE || st_nc[r])
                        qhx[orow] = hxn[r + 1];
                    if (!EDGE || st_cn[r])
                        qhy[orow] = hyn[r];
                    if (!EDGE || st_nn[r])
                        qez[orow] = vez;
                }
                if (!EDGE || st_cc[r])
                    qhz[orow] = hzn[r + 1];
                if (!EDGE || st_cn[r])
                    qex[orow] = vex;
                if (!EDGE || st_nc[r])
                    qey[orow] = vey;
            }
        }

#pragma unroll
        for (int rr = 0; rr < NR; ++rr)
            exk[rr] = exn[rr];
#pragma unroll
        for (int rr = 0; rr < NR - 1; ++rr)
            eyk[rr] = eyn[rr];
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            hxm[r] = hxn[r + 1];
            hym[r] = hyn[r];
        }
    }
}

/* The node-centred arrays have I+1 columns.  For the usual power-of-two I the last block in x would
 * hold the single column i = I and spend a whole chunk's time in the TMA pipeline for it, keeping an
 * SM slot busy (11 % of the blocks at 1024^3 with 128-wide tiles, 33 % at 256^3).  That column only
 * carries PEC values of Ey and Ez (copied) and the Hx update of main.c:448, so such a block does it
 * directly, one (row, plane) per thread, and leaves. */
__device__ __forceinline__ void last_column_sweep(const Geo &g, const Fld &a, const Fld &b, const double cH,
                                                  const int by0, const int BY, const int kl0, const int kl1)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const int items = BY * (kl1 - kl0);
    for (int t = tid; t < items; t += nthr) {
        const int j = by0 + t % BY, kl = kl0 + t / BY;
        if (j > g.J)
            continue;
        const long long o = g.I + (long long)g.P * (j + (long long)g.R * kl);
        const bool cell = kl <= g.nk;
        if (j < g.J) {
            const double ey = a.ey[o];
            b.ey[o] = ey; /* Ey(I, j, k) is on the wall: carried over */
            if (cell)
                b.hx[o] = yee(a.hx[o], cH, a.ey[o + g.PR], ey, a.ez[o + g.P], a.ez[o]); /* main.c:448 */
        }
        if (cell)
            b.ez[o] = a.ez[o]; /* Ez(I, j, k) is on the wall: carried over */
    }
}

/* 256-thread blocks, two per SM -- or one 512-thread block for the largest tiles */
template <int TY, int CWX, int CWY>
__global__ void __launch_bounds__((CWX * CWY > 8) ? 512 : 256, (CWX * CWY > 8) ? 1 : 2)
k_step_fused_tma(Geo g, const __grid_constant__ TmaMaps maps, Fld a, Fld b, double cH, double cE, Src s,
                 Span sp, int stages)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[kTmaMaxStages];
    double *ring = reinterpret_cast<double *>(smem_raw);

    const int bx0 = blockIdx.x * blockDim.x;
    const int by0 = blockIdx.y * blockDim.y * TY;
    const int kl0 = sp.kl_begin + blockIdx.z * sp.kchunk;
    const int kl1 = min(kl0 + sp.kchunk, sp.kl_end);

    if (bx0 == g.I) { /* block-uniform: this block holds column I only */
        last_column_sweep(g, a, b, cH, by0, (int)blockDim.y * TY, kl0, kl1);
        return;
    }

    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int st = 0; st < stages; ++st)
            tma::mbar_init(full + st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const bool interior = bx0 >= 1 && bx0 + (int)blockDim.x <= g.I && by0 >= 1 &&
                          by0 + (int)blockDim.y * TY <= g.J;
    if (interior)
        fused_tma_sweep<TY, false, CWX, CWY>(g, maps, b, cH, cE, s, stages, ring, full, bx0, by0, kl0, kl1);
    else
        fused_tma_sweep<TY, true, CWX, CWY>(g, maps, b, cH, cE, s, stages, ring, full, bx0, by0, kl0, kl1);
}

} /* namespace fdtd */
