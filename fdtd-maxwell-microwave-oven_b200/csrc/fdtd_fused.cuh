/*
 * fdtd_fused.cuh -- one launch per time step: H update and E update in a single upward sweep.
 *
 * The split kernels move 144 B per cell-update (each half-step reads the other family and
 * read-modify-writes its own).  Sweeping the planes upwards, plane k of H can be finished and
 * then immediately consumed by plane k of E while both are still in registers:
 *
 *     H_new(k) = H(k) + cH * curl E(k, k+1)          needs E(k), E(k+1), H(k)      from HBM
 *     E_new(k) = E(k) + cE * curl H_new(k, k-1)      needs H_new(k) (just computed), H_new(k-1)
 *                                                    (previous plane, registers), E(k) (registers)
 *
 * so a step reads the six arrays once and writes them once: 96 B per cell-update.  An in-place
 * sweep would race between neighbouring tiles (a tile would read E or H that its neighbour has
 * already advanced), so the state is double-buffered: every read comes from `a` (time n), every
 * write goes to `b` (time n+1), and the host swaps the two after the launch.  Elements that a
 * step does not update (tangential E on the PEC walls) are copied, so `b` is always complete.
 *
 * Work decomposition is the register-strip scheme of k_update_*_march: a thread owns the column
 * (i, jb..jb+TY-1), a warp 32 consecutive i, a block walks a chunk of planes.  What a thread
 * needs of H_new from outside its strip is recomputed instead of exchanged (nothing is read back
 * from `b`): row jb-1 of Hx/Hz by the thread itself, column i-1 of Hy/Hz by the neighbouring lane
 * (shuffle) -- lane 0 recomputes it for its warp.  A chunk that does not start at the bottom wall
 * first recomputes H_new of the plane below it.  All of this is arithmetic on operands that were
 * needed anyway or that L1/L2 still hold; HBM traffic stays at one read and one write per element.
 *
 * Instruction economy (the first version of this kernel was issue-bound at 536 instructions per
 * warp and plane, mostly predicate logic):
 *   - loads carry no predicate.  A valid output only ever depends on valid inputs (SURVEY.md
 *     Appendix A), so a load from a position outside an array's extent may return anything; the
 *     allocation has guard margins so that every such address is still inside it;
 *   - stores are predicated only in blocks that touch a wall (EDGE = true, chosen per block);
 *     interior blocks run a straight-line body;
 *   - everything that depends on the plane only (source plane, top plane, PEC planes, the
 *     "H only" prologue) is a warp-uniform branch.
 *
 * Same arithmetic as everywhere else: F = F + c * ((a - b) - (d - e)) with round-to-nearest
 * intrinsics, reference operand order (main.c:448-461, 486-499).  The source (main.c:712-753) and
 * the PEC walls are fused exactly as in the split kernels.
 */
#pragma once

#include "fdtd_types.cuh"

namespace fdtd {

template <int TY>
struct FusedCfg {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = TY >= 4 ? 2 : (TY == 3 ? 3 : 4);
};

template <int TY, bool EDGE>
__device__ __forceinline__ void fused_sweep(const Geo &g, const Fld &a, const Fld &b, const double cH,
                                            const double cE, const Src &s, const Span &sp, const int i,
                                            const int jb, const int kl0, const int kl1)
{
    constexpr int NR = TY + 2; /* rows jb-1 .. jb+TY, index rr = row - (jb-1) */
    const unsigned lane = threadIdx.x & 31u;
    const int P = g.P;

    /* store / update masks, loop invariant; all true in interior blocks */
    bool st_nc[TY], st_cn[TY], st_cc[TY], st_nn[TY], up_x[TY], up_y[TY], up_z[TY];
    if (EDGE) {
        const bool xn = i <= g.I, xc = i < g.I, xi = i >= 1 && i < g.I;
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const int j = jb + r;
            const bool jn = j <= g.J, jc = j < g.J, ji = j >= 1 && j < g.J;
            st_nc[r] = xn && jc; /* Ey, Hx */
            st_cn[r] = xc && jn; /* Ex, Hy */
            st_cc[r] = xc && jc; /* Hz */
            st_nn[r] = xn && jn; /* Ez */
            up_x[r] = xc && ji;
            up_y[r] = xi && jc;
            up_z[r] = xi && ji;
        }
    }

    /* Ex, Ey of the chunk's first plane are updated only if H_new of the plane below is known:
     * start one plane lower in "H only" mode unless that plane is outside the bottom wall. */
    const bool below = (kl0 - 1 + g.kbase) >= 1;
    const int kstart = below ? kl0 - 1 : kl0;
    const int q = i + P * jb; /* in-plane offset of (i, jb); rows are q + r*P */
    long long pl = (long long)kstart * g.PR;

    /* E of plane kstart (time n); persistent across the sweep */
    double exk[NR], eyk[NR - 1], exl[TY + 1], eyl[TY];
    {
        const double *pex = a.ex + pl, *pey = a.ey + pl;
#pragma unroll
        for (int rr = 0; rr < NR; ++rr)
            exk[rr] = __ldg(pex + (q + (rr - 1) * P));
#pragma unroll
        for (int rr = 0; rr < NR - 1; ++rr)
            eyk[rr] = __ldg(pey + (q + (rr - 1) * P));
        if (lane == 0u) {
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                exl[r] = __ldg(pex + (q - 1 + r * P));
#pragma unroll
            for (int r = 0; r < TY; ++r)
                eyl[r] = __ldg(pey + (q - 1 + r * P));
        }
        if (s.on && kstart == s.kl) { /* first set_source: Ex = 0 on the patch (main.c:749) */
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
    double hxm[TY], hym[TY]; /* H_new of the plane below, own rows */
#pragma unroll
    for (int r = 0; r < TY; ++r)
        hxm[r] = hym[r] = 0.0;

    for (int kl = kstart; kl < kl1; ++kl, pl += g.PR) {
        const bool srck = s.on && kl == s.kl; /* global plane k = 0 carries the source */
        const double *pex = a.ex + pl, *pey = a.ey + pl, *pez = a.ez + pl;
        const double *phx = a.hx + pl, *phy = a.hy + pl, *phz = a.hz + pl;

        if (sp.prefetch > 0 && kl + sp.prefetch <= g.nk + 1) { /* own rows, a few planes ahead, into L2 */
            const long long ahead = (long long)sp.prefetch * g.PR;
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                prefetch_l2(pex + ahead + (q + r * P));
                prefetch_l2(pey + ahead + (q + r * P));
                prefetch_l2(pez + ahead + (q + r * P));
                prefetch_l2(phx + ahead + (q + r * P));
                prefetch_l2(phy + ahead + (q + r * P));
                prefetch_l2(phz + ahead + (q + r * P));
            }
        }

        /* ---- loads, time n (no predicates: see the header) ---- */
        double exn[NR], eyn[NR - 1], ezk[NR], hxo[TY + 1], hyo[TY], hzo[TY + 1];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
            exn[rr] = __ldg(pex + g.PR + (q + (rr - 1) * P));
            ezk[rr] = __ldg(pez + (q + (rr - 1) * P));
        }
#pragma unroll
        for (int rr = 0; rr < NR - 1; ++rr) {
            eyn[rr] = __ldg(pey + g.PR + (q + (rr - 1) * P));
            hxo[rr] = __ldg(phx + (q + (rr - 1) * P));
            hzo[rr] = __ldg(phz + (q + (rr - 1) * P));
        }
#pragma unroll
        for (int r = 0; r < TY; ++r)
            hyo[r] = __ldg(phy + (q + r * P));

        /* i+1 neighbours of Ez (own rows) and Ey (rows -1 .. TY-1): next lane, or a direct read on
         * the warp's right edge */
        double ezi[TY], eyi[TY + 1];
#pragma unroll
        for (int r = 0; r < TY; ++r)
            ezi[r] = __shfl_down_sync(0xffffffffu, ezk[r + 1], 1);
#pragma unroll
        for (int rr = 0; rr <= TY; ++rr)
            eyi[rr] = __shfl_down_sync(0xffffffffu, eyk[rr], 1);
        if (lane == 31u) {
#pragma unroll
            for (int r = 0; r < TY; ++r)
                ezi[r] = __ldg(pez + (q + 1 + r * P));
#pragma unroll
            for (int rr = 0; rr <= TY; ++rr)
                eyi[rr] = __ldg(pey + (q + 1 + (rr - 1) * P));
        }

        /* column i-1 of Hy, Hz at time n+1/2, recomputed by lane 0 for its warp */
        double hyln[TY], hzln[TY];
        if (lane == 0u) {
            double exln[TY + 1], eyln[TY], ezl[TY];
#pragma unroll
            for (int r = 0; r <= TY; ++r)
                exln[r] = __ldg(pex + g.PR + (q - 1 + r * P));
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                eyln[r] = __ldg(pey + g.PR + (q - 1 + r * P));
                ezl[r] = __ldg(pez + (q - 1 + r * P));
            }
            if (srck) {
#pragma unroll
                for (int r = 0; r < TY; ++r)
                    if (in_patch(s, i - 1, jb + r))
                        ezl[r] = s.vals[i - 1 - s.i0];
            }
            double ezo[TY]; /* Ez(i, row r) as column i-1 sees it */
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                ezo[r] = ezk[r + 1];
                if (srck && in_patch(s, i, jb + r))
                    ezo[r] = s.vals[i - s.i0];
            }
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                hyln[r] = yee(__ldg(phy + (q - 1 + r * P)), cH, ezo[r], ezl[r], exln[r], exl[r]);
                hzln[r] = yee(__ldg(phz + (q - 1 + r * P)), cH, exl[r + 1], exl[r], eyk[r + 1], eyl[r]);
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

        if (srck) { /* first set_source: Ez on the patch (main.c:748), own column and lane 31's i+1 */
#pragma unroll
            for (int rr = 0; rr < NR; ++rr)
                if (in_patch(s, i, jb - 1 + rr))
                    ezk[rr] = s.vals[i - s.i0];
#pragma unroll
            for (int r = 0; r < TY; ++r)
                if (in_patch(s, i + 1, jb + r))
                    ezi[r] = s.vals[i + 1 - s.i0];
        }

        /* ---- H at time n+1/2 on this plane: own rows plus row jb-1 of Hx, Hz ---- */
        double hxn[TY + 1], hzn[TY + 1], hyn[TY];
#pragma unroll
        for (int rr = 0; rr <= TY; ++rr) {
            hxn[rr] = yee(hxo[rr], cH, eyn[rr], eyk[rr], ezk[rr + 1], ezk[rr]); /* main.c:448 */
            hzn[rr] = yee(hzo[rr], cH, exk[rr + 1], exk[rr], eyi[rr], eyk[rr]); /* main.c:460 */
        }
#pragma unroll
        for (int r = 0; r < TY; ++r)
            hyn[r] = yee(hyo[r], cH, ezi[r], ezk[r + 1], exn[r + 1], exk[r + 1]); /* main.c:454 */
        if (srck) { /* second set_source, main.c:750-751 */
#pragma unroll
            for (int rr = 0; rr <= TY; ++rr)
                if (in_patch(s, i, jb - 1 + rr)) {
                    hxn[rr] = s.vals[s.n + i - s.i0];
                    hzn[rr] = 0.0;
                }
        }

        if (kl >= kl0) { /* not the "H only" prologue plane */
            const bool cell = kl <= g.nk;                 /* Ez, Hx, Hy exist on this plane */
            const bool kin = (kl - 1 + g.kbase) >= 1 && cell; /* Ex, Ey are updated on this plane */
            double *qex = b.ex + pl, *qey = b.ey + pl, *qez = b.ez + pl;
            double *qhx = b.hx + pl, *qhy = b.hy + pl, *qhz = b.hz + pl;
#pragma unroll
            for (int r = 0; r < TY; ++r) {
                const int orow = q + r * P;
                /* Hy, Hz of column i-1 from the lane to the left */
                double hyim = __shfl_up_sync(0xffffffffu, hyn[r], 1);
                double hzim = __shfl_up_sync(0xffffffffu, hzn[r + 1], 1);
                if (lane == 0u) {
                    hyim = hyln[r];
                    hzim = hzln[r];
                }
                double vex = exk[r + 1], vey = eyk[r + 1], vez = ezk[r + 1];
                if (kin) {
                    const double ux = yee(vex, cE, hzn[r + 1], hzn[r], hyn[r], hym[r]);   /* main.c:486 */
                    const double uy = yee(vey, cE, hxn[r + 1], hxm[r], hzn[r + 1], hzim); /* main.c:492 */
                    vex = (!EDGE || up_x[r]) ? ux : vex;
                    vey = (!EDGE || up_y[r]) ? uy : vey;
                }
                if (cell) {
                    const double uz = yee(vez, cE, hyn[r], hyim, hxn[r + 1], hxn[r]);     /* main.c:498 */
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
                    qex[orow] = vex; /* PEC rows / planes and the source's Ex = 0 are carried over */
                if (!EDGE || st_nc[r])
                    qey[orow] = vey;
            }
        }

        /* ---- next plane ---- */
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

template <int TY>
__global__ void __launch_bounds__(FusedCfg<TY>::kThreads, FusedCfg<TY>::kMinBlocks)
k_step_fused(Geo g, Fld a, Fld b, double cH, double cE, Src s, Span sp)
{
    const int bx0 = blockIdx.x * blockDim.x;
    const int by0 = blockIdx.y * blockDim.y * TY;
    const int i = bx0 + threadIdx.x;
    const int jb = by0 + threadIdx.y * TY;
    const int kl0 = sp.kl_begin + blockIdx.z * sp.kchunk;
    const int kl1 = min(kl0 + sp.kchunk, sp.kl_end);
    if (jb > g.J)
        return; /* warp-uniform: threadIdx.y is constant inside a warp */
    /* interior block: every element of the tile exists in all six arrays and is updated */
    const bool interior = bx0 >= 1 && bx0 + (int)blockDim.x <= g.I - 0 && by0 >= 1 &&
                          by0 + (int)blockDim.y * TY <= g.J;
    if (interior)
        fused_sweep<TY, false>(g, a, b, cH, cE, s, sp, i, jb, kl0, kl1);
    else
        fused_sweep<TY, true>(g, a, b, cH, cE, s, sp, i, jb, kl0, kl1);
}

} /* namespace fdtd */
