/*
 * fdtd_step2_tma.cuh -- TWO time steps per sweep ("kernel" = 4): temporal blocking on top of the
 * TMA-staged fused sweep of fdtd_fused_tma.cuh.
 *
 * Why.  k_step_fused_tma moves every array once in and once out per time step (96 B per cell-update)
 * and runs at 0.9 of the HBM copy rate (profiles/r02_tma_default_*): only fewer bytes can make it
 * faster.  Here one sweep advances the state by two steps, so the six arrays are read and written
 * once per TWO steps: 48 B per cell-update plus the overlap of the tiles.
 *
 * How.  Sweeping planes upward, iteration k brings in plane k+1 of the old state (time n) and does
 *     A   H1(k)   = H0(k)   + cH * curl E0(k, k+1)          time n + 1/2
 *     B   E1(k)   = E0(k)   + cE * curl H1(k, k-1)          time n + 1
 *     C   H2(k-1) = H1(k-1) + cH * curl E1(k-1, k)          time n + 3/2
 *     D   E2(k-1) = E1(k-1) + cE * curl H2(k-1, k-2)        time n + 2   -> stored with H2(k-1)
 * Every thread owns one column (i) and two rows (j, j+1) of an EXTENDED tile of 32 x 2*WY sites and
 * evaluates all four stages there; what a stage needs from the neighbouring sites comes from the
 * neighbouring lanes (shuffles, x) or, across warps, from one row per warp in shared memory (y).
 * Nothing is recomputed and nothing is special at tile edges: the outermost sites of the extended
 * tile simply compute values nobody uses.  Each stage loses one ring of validity, so the tile that is
 * stored is 28 x (2*WY - 3): tiles overlap by 4 columns and 3 rows (71 % useful sites for WY = 8) --
 * the price of halving the HBM traffic.  The 28-column pitch keeps every box start even (TMA needs
 * 16-byte aligned boxes) and every stored row sector-aligned.
 *
 * A chunk of planes [a, b) starts two planes early (H1(a-2), then H1/E1(a-1), then H2(a-1)) and reads
 * up to plane b+1; with 32..64 planes per chunk that is 6..12 % extra reads.
 *
 * The arithmetic per element is the reference's (yee() in fdtd_types.cuh, operand order of
 * main.c:448-461, 486-499, no FMA), PEC walls are elements whose update is skipped at BOTH levels,
 * and the source of main.c:712-753 is substituted at both levels with the amplitudes of its own time
 * (SURVEY.md B.4): results stay bit-identical to two passes of the sequential loop body.
 */
#pragma once

#include "fdtd_fused_tma.cuh"

namespace fdtd {

constexpr int kS2TileX = 28;      /* stored columns per block */
constexpr int kS2BoxW = 34;       /* loaded columns: 32 sites + the +1 neighbour, padded to an even width */

struct Src2 {
    Src s1, s2; /* amplitudes of the first and of the second step */
};

/* WS (warp-specialised, persistent form): the TMA loads are issued by a separate producer warp
 * (step2_produce); the compute warps synchronise among themselves on a named barrier and hand ring
 * slots back through `empty` mbarriers. */
template <int WY, bool EDGE, bool WS>
__device__ __forceinline__ void step2_sweep(const Geo &g, const TmaMaps &maps, const Fld &b, const double cH,
                                            const double cE, const Src2 &src, const int stages, double *ring,
                                            double *xh, double *xe, unsigned long long *full, const int X0,
                                            const int Y0, const int kl0, const int kl1,
                                            unsigned long long *empty = nullptr, unsigned *prog = nullptr,
                                            const int zmod = 0, const int zrot_in = 0, const int zrot_out = 0)
{
    auto block_sync = [] {
        if (WS)
            asm volatile("bar.sync 1, %0;" ::"n"(32 * WY) : "memory"); /* the compute warps only */
        else
            __syncthreads();
    };
    constexpr int BYE = 2 * WY, W = kS2BoxW, HH = BYE + 1;
    const int box = tma_box_doubles(W - 4, HH - 2);
    const int stage_doubles = 6 * box;
    const int lane = threadIdx.x, w = threadIdx.y;
    const int i = X0 - 2 + lane;
    const int j0 = Y0 - 2 + 2 * w;
    const int P = g.P;
    const bool leader = lane == 0 && w == 0;
    const Src &s1 = src.s1, &s2 = src.s2;

    /* what this thread may update / store (EDGE blocks only; interior blocks have everything true
     * except the output window) */
    bool up_x[2], up_y[2], up_z[2], st_nc[2], st_cn[2], st_cc[2], st_nn[2];
    const bool outx = lane >= 2 && lane < 30;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = j0 + r, row = 2 * w + r;
        const bool out = outx && row >= 2 && row < BYE - 1;
        if (EDGE) {
            const bool xn = i >= 0 && i <= g.I, xc = i >= 0 && i < g.I, xi = i >= 1 && i < g.I;
            const bool jn = j >= 0 && j <= g.J, jc = j >= 0 && j < g.J, ji = j >= 1 && j < g.J;
            up_x[r] = xc && ji;
            up_y[r] = xi && jc;
            up_z[r] = xi && ji;
            st_nc[r] = out && xn && jc;
            st_cn[r] = out && xc && jn;
            st_cc[r] = out && xc && jc;
            st_nn[r] = out && xn && jn;
        } else {
            up_x[r] = up_y[r] = up_z[r] = true;
            st_nc[r] = st_cn[r] = st_cc[r] = st_nn[r] = out;
        }
    }

    const int g0 = 1 - g.kbase;                /* local index of the global plane k = 0 */
    const int kstart = max(kl0 - 2, g0);
    const int nplanes = kl1 - kstart + 2;      /* planes kstart .. kl1 + 1 are loaded */
    const unsigned full_bytes = 6u * (unsigned)(W * HH) * 8u;

    auto issue = [&](int slot, int n) {
        double *dst = ring + (size_t)slot * stage_doubles;
        unsigned long long *bar = full + slot;
        const int plane = kstart + n;
        tma::mbar_expect_tx(bar, full_bytes);
#pragma unroll
        for (int a = 0; a < 6; ++a)
            tma::load_box(dst + a * box, &maps.m[a], X0 - 2, Y0 - 2, zmod ? (plane + 1 + zrot_in) % zmod : plane + 1,
                          bar); /* the maps start at plane -1 */
    };
    if (!WS && leader)
        for (int n = 0; n < stages && n < nplanes; ++n)
            issue(n, n);

    const int c0 = lane + W * (2 * w);         /* (i, j0) inside a box */
    const int xrow = w * 32 + lane;            /* this thread's slot in an exchange row */
    const int xrow_dn = max(w - 1, 0) * 32 + lane, xrow_up = min(w + 1, WY - 1) * 32 + lane;
    double *xh_hx = xh, *xh_hz = xh + 32 * WY;                   /* top row of each strip: Hx, Hz */
    double *xe_ex = xe, *xe_ez = xe + 2 * 32 * WY;               /* bottom row: Ex, Ez; two planes deep */

    /* E0 of plane kstart (own sites) */
    double ex0[2], ey0[2];
    tma::mbar_wait(full + 0, 0);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        ex0[r] = ring[c0 + r * W];
        ey0[r] = ring[box + c0 + r * W];
    }
    double hx1p[2] = {0, 0}, hy1p[2] = {0, 0}, hz1p[2] = {0, 0};      /* H1(k-1) */
    double ex1p[2] = {0, 0}, ey1p[2] = {0, 0}, ez1p[2] = {0, 0};      /* E1(k-1) */
    double hx2pp[2] = {0, 0}, hy2pp[2] = {0, 0};                      /* H2(k-2) */

    long long pl = (long long)(kstart - 1) * g.PR;                     /* plane k-1, where stage D stores */
    int slot = 0;
    unsigned phase = 0;
    for (int k = kstart, n = 0; k <= kl1; ++k, ++n, pl += g.PR) {
        if (zmod) /* rolling form: plane k - 1 goes to its slot of the ring (b points at slot 1) */
            pl = (long long)((k + zrot_out) % zmod - 1) * g.PR;
        int slot1 = slot + 1;
        unsigned phase1 = phase;
        if (slot1 == stages) {
            slot1 = 0;
            phase1 ^= 1u;
        }
        const int slot_now = slot;
        slot = slot1;
        phase = phase1;
        tma::mbar_wait(full + slot1, phase1);
        const double *cur = ring + (size_t)slot_now * stage_doubles, *nxt = ring + (size_t)slot1 * stage_doubles;
        const double *sex = cur, *sey = cur + box, *sez = cur + 2 * box;
        const double *shx = cur + 3 * box, *shy = cur + 4 * box, *shz = cur + 5 * box;
        const int gk = k - 1 + g.kbase;                     /* global index of plane k */
        /* by GLOBAL plane: a halo plane of a slab is an ordinary plane of the cavity; only the cavity's top
         * node plane K has no cells */
        const bool cell1 = gk < g.K, kin1 = gk >= 1 && cell1;
        const bool cell2 = gk - 1 < g.K, kin2 = gk - 1 >= 1 && cell2;
        const bool srck1 = s1.on && k == s1.kl;             /* stage A/B work on the source plane */
        const bool srck2 = s2.on && k - 1 == s2.kl;         /* stage C/D do */

        /* ---- A: H1(k) ---- */
        double ez0[2], ez0i[2], ez0j[2], ex0j[2], ey0i[2], ex0n[2], ey0n[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            ez0[r] = sez[c0 + r * W];
            ez0i[r] = sez[c0 + 1 + r * W];
            ey0i[r] = sey[c0 + 1 + r * W];
            ex0n[r] = nxt[c0 + r * W];
            ey0n[r] = nxt[box + c0 + r * W];
        }
        ez0j[0] = ez0[1];
        ez0j[1] = sez[c0 + 2 * W];
        ex0j[0] = ex0[1];
        ex0j[1] = sex[c0 + 2 * W];
        if (srck1) { /* first set_source of step 1: Ez = amplitude, Ex = 0 on the patch (main.c:748-749) */
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int j = j0 + r;
                if (in_patch(s1, i, j)) {
                    ez0[r] = s1.vals[i - s1.i0];
                    ex0[r] = 0.0;
                }
                if (in_patch(s1, i + 1, j))
                    ez0i[r] = s1.vals[i + 1 - s1.i0];
                if (in_patch(s1, i, j + 1)) {
                    ez0j[r] = s1.vals[i - s1.i0];
                    ex0j[r] = 0.0;
                }
            }
        }
        double hx1[2], hy1[2], hz1[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            hx1[r] = yee(shx[c0 + r * W], cH, ey0n[r], ey0[r], ez0j[r], ez0[r]);
            hy1[r] = yee(shy[c0 + r * W], cH, ez0i[r], ez0[r], ex0n[r], ex0[r]);
            hz1[r] = yee(shz[c0 + r * W], cH, ex0j[r], ex0[r], ey0i[r], ey0[r]);
        }
        if (srck1) { /* second set_source of step 1 (main.c:750-751) */
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (in_patch(s1, i, j0 + r)) {
                    hx1[r] = s1.vals[s1.n + i - s1.i0];
                    hz1[r] = 0.0;
                }
        }
        xh_hx[xrow] = hx1[1];
        xh_hz[xrow] = hz1[1];
        block_sync(); /* 1: H1 rows visible; every thread is done with the stage of plane k */
        if (leader) {
            if (WS) {
                tma::mbar_arrive(empty + slot_now); /* the producer warp may refill the slot */
                if (prog)
                    atomicAdd(prog + n, 1u);        /* this block is done with plane index n */
            } else if (n + stages < nplanes) {
                issue(slot_now, n + stages);
            }
        }

        /* ---- B: E1(k) ---- */
        double ex1[2], ey1[2], ez1[2];
        {
            const double hxjm[2] = {xh_hx[xrow_dn], hx1[0]}, hzjm[2] = {xh_hz[xrow_dn], hz1[0]};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double hzim = __shfl_up_sync(0xffffffffu, hz1[r], 1);
                const double hyim = __shfl_up_sync(0xffffffffu, hy1[r], 1);
                ex1[r] = ex0[r];
                ey1[r] = ey0[r];
                ez1[r] = ez0[r];
                if (kin1) {
                    const double ux = yee(ex0[r], cE, hz1[r], hzjm[r], hy1[r], hy1p[r]);
                    const double uy = yee(ey0[r], cE, hx1[r], hx1p[r], hz1[r], hzim);
                    ex1[r] = up_x[r] ? ux : ex0[r];
                    ey1[r] = up_y[r] ? uy : ey0[r];
                }
                if (cell1) {
                    const double uz = yee(ez0[r], cE, hy1[r], hyim, hx1[r], hxjm[r]);
                    ez1[r] = up_z[r] ? uz : ez0[r];
                }
            }
        }
        if (s2.on && k == s2.kl) { /* first set_source of step 2 overwrites the patch before anything reads it */
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (in_patch(s2, i, j0 + r)) {
                    ez1[r] = s2.vals[i - s2.i0];
                    ex1[r] = 0.0;
                }
        }
        double *xe_ex_k = xe_ex + (k & 1) * 32 * WY, *xe_ez_k = xe_ez + (k & 1) * 32 * WY;
        const double *xe_ex_p = xe_ex + ((k & 1) ^ 1) * 32 * WY, *xe_ez_p = xe_ez + ((k & 1) ^ 1) * 32 * WY;
        xe_ex_k[xrow] = ex1[0];
        xe_ez_k[xrow] = ez1[0];
        block_sync(); /* 2: (orders the H1 reads above before the H2 rows below) */

        /* ---- C: H2(k-1) ---- */
        double hx2[2], hy2[2], hz2[2];
        {
            const double ezjp[2] = {ez1p[1], xe_ez_p[xrow_up]}, exjp[2] = {ex1p[1], xe_ex_p[xrow_up]};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double ezip = __shfl_down_sync(0xffffffffu, ez1p[r], 1);
                const double eyip = __shfl_down_sync(0xffffffffu, ey1p[r], 1);
                hx2[r] = yee(hx1p[r], cH, ey1[r], ey1p[r], ezjp[r], ez1p[r]);
                hy2[r] = yee(hy1p[r], cH, ezip, ez1p[r], ex1[r], ex1p[r]);
                hz2[r] = yee(hz1p[r], cH, exjp[r], ex1p[r], eyip, ey1p[r]);
            }
        }
        if (srck2) { /* second set_source of step 2 */
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (in_patch(s2, i, j0 + r)) {
                    hx2[r] = s2.vals[s2.n + i - s2.i0];
                    hz2[r] = 0.0;
                }
        }
        xh_hx[xrow] = hx2[1];
        xh_hz[xrow] = hz2[1];
        block_sync(); /* 3 */

        /* ---- D: E2(k-1), stores ---- */
        {
            const double hxjm[2] = {xh_hx[xrow_dn], hx2[0]}, hzjm[2] = {xh_hz[xrow_dn], hz2[0]};
            const bool store = k - 1 >= kl0;
            double *qex = b.ex + pl, *qey = b.ey + pl, *qez = b.ez + pl;
            double *qhx = b.hx + pl, *qhy = b.hy + pl, *qhz = b.hz + pl;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double hzim = __shfl_up_sync(0xffffffffu, hz2[r], 1);
                const double hyim = __shfl_up_sync(0xffffffffu, hy2[r], 1);
                double vex = ex1p[r], vey = ey1p[r], vez = ez1p[r];
                if (kin2) {
                    const double ux = yee(vex, cE, hz2[r], hzjm[r], hy2[r], hy2pp[r]);
                    const double uy = yee(vey, cE, hx2[r], hx2pp[r], hz2[r], hzim);
                    vex = up_x[r] ? ux : vex;
                    vey = up_y[r] ? uy : vey;
                }
                if (cell2) {
                    const double uz = yee(vez, cE, hy2[r], hyim, hx2[r], hxjm[r]);
                    vez = up_z[r] ? uz : vez;
                }
                if (store) {
                    const int o = i + P * (j0 + r);
                    if (cell2) {
                        if (st_nc[r]) qhx[o] = hx2[r];
                        if (st_cn[r]) qhy[o] = hy2[r];
                        if (st_nn[r]) qez[o] = vez;
                    }
                    if (st_cc[r]) qhz[o] = hz2[r];
                    if (st_cn[r]) qex[o] = vex;
                    if (st_nc[r]) qey[o] = vey;
                }
            }
        }

        /* rotate */
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            hx2pp[r] = hx2[r];
            hy2pp[r] = hy2[r];
            hx1p[r] = hx1[r];
            hy1p[r] = hy1[r];
            hz1p[r] = hz1[r];
            ex1p[r] = ex1[r];
            ey1p[r] = ey1[r];
            ez1p[r] = ez1[r];
            ex0[r] = ex0n[r];
            ey0[r] = ey0n[r];
        }
    }
}

/* block = 32 x WY threads, two resident blocks per SM */
template <int WY>
__global__ void __launch_bounds__(32 * WY, (WY > 8) ? 1 : 2)
k_step2_tma(Geo g, const __grid_constant__ TmaMaps maps, Fld b, double cH, double cE, Src2 src, Span sp, int stages)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[kTmaMaxStages];
    __shared__ double xh[2 * 32 * WY];
    __shared__ double xe[4 * 32 * WY];
    double *ring = reinterpret_cast<double *>(smem_raw);

    constexpr int BYE = 2 * WY;
    /* Which tile?  Blocks start in the order of their linear index, a few hundred at a time, and sweep
     * upward at the same pace: two blocks share the halo of their tiles through L2 only if they start
     * close together.  Row-major order puts the tile above a whole row of tiles (37 at 1024^3) later,
     * too far for L2; so tiles are numbered column by column inside bands of sp.band tile rows --
     * then both the tile above and the tile to the right start within a few blocks. */
    int tx = blockIdx.x, ty = blockIdx.y;
    if (sp.band > 1) {
        const int nx = gridDim.x, ny = gridDim.y;
        const int lin = blockIdx.x + nx * blockIdx.y;
        const int band = lin / (sp.band * nx), rem = lin - band * sp.band * nx;
        const int rows = min(sp.band, ny - band * sp.band);
        tx = rem / rows;
        ty = band * sp.band + rem - tx * rows;
    }
    const int X0 = tx * kS2TileX;
    const int Y0 = ty * (BYE - 3);
    const int kl0 = sp.kl_begin + blockIdx.z * sp.kchunk;
    const int kl1 = min(kl0 + sp.kchunk, sp.kl_end);
    if (X0 > g.I || Y0 > g.J)
        return; /* padding of a grid rounded up to whole clusters */

    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int st = 0; st < stages; ++st)
            tma::mbar_init(full + st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    /* every site of the extended tile strictly inside the walls: no predicate can be false */
    const bool interior = X0 - 2 >= 1 && X0 + 29 < g.I && Y0 - 2 >= 1 && Y0 - 2 + BYE - 1 < g.J;
    if (interior)
        step2_sweep<WY, false, false>(g, maps, b, cH, cE, src, stages, ring, xh, xe, full, X0, Y0, kl0, kl1, nullptr,
                                      nullptr, sp.zmod, sp.zrot_in, sp.zrot_out);
    else
        step2_sweep<WY, true, false>(g, maps, b, cH, cE, src, stages, ring, xh, xe, full, X0, Y0, kl0, kl1, nullptr,
                                     nullptr, sp.zmod, sp.zrot_in, sp.zrot_out);
}

/* The producer warp of the persistent form: lane 0 issues the six boxes of every plane of the sweep,
 * `stages` planes ahead of the compute warps.  Before re-using a ring slot it waits for the compute warps
 * to hand it back (`empty`); before asking for plane m it waits until every block of the round is done
 * with plane m - stages - window (flow control: blocks that sweep within a few planes of each other find
 * the halo sectors they share in L2 -- ncu: reads 1.06 x the arrays instead of 1.4 x). */
template <int WY>
__device__ __forceinline__ void step2_produce(const Geo &g, const TmaMaps &maps, const int stages, double *ring,
                                              unsigned long long *full, unsigned long long *empty, const int X0,
                                              const int Y0, const int kl0, const int kl1, const unsigned *prog,
                                              const unsigned nblocks, const int window)
{
    constexpr int BYE = 2 * WY, W = kS2BoxW, HH = BYE + 1;
    const int box = tma_box_doubles(W - 4, HH - 2);
    const int stage_doubles = 6 * box;
    const int g0 = 1 - g.kbase;
    const int kstart = max(kl0 - 2, g0);
    const int nplanes = kl1 - kstart + 2;
    const unsigned full_bytes = 6u * (unsigned)(W * HH) * 8u;
    /* lane 0 does the work; the other lanes wait for it here, so that the warp reaches the block-wide
     * barrier of the round loop as one (bar.sync is a per-warp instruction) */
    if (threadIdx.x == 0) {
        int slot = 0;
        unsigned parity = 1; /* of the wait on `empty` for the current pass over the ring (first pass: none) */
        for (int m = 0; m < nplanes; ++m) {
            if (m >= stages)
                tma::mbar_wait(empty + slot, parity);
            const int idx = m - stages - window;
            if (prog && idx >= 0) {
                const volatile unsigned *gate = prog + idx;
                while (*gate < nblocks)
                    __nanosleep(64);
            }
            double *dst = ring + (size_t)slot * stage_doubles;
            tma::mbar_expect_tx(full + slot, full_bytes);
#pragma unroll
            for (int a = 0; a < 6; ++a)
                tma::load_box(dst + a * box, &maps.m[a], X0 - 2, Y0 - 2, kstart + m + 1, full + slot);
            if (++slot == stages) {
                slot = 0;
                parity ^= 1u;
            }
        }
    }
    __syncwarp();
}

/* Persistent, cooperative, warp-specialised form: one block of WY compute warps + 1 producer warp per
 * resident slot of the GPU; the blocks take the tiles of the cavity round by round (a round = gridDim.x
 * consecutive tiles, whole rows of tiles) and sweep the WHOLE plane range of the launch -- no chunks, so
 * no run-in planes are read twice.  Needs a cooperative launch: every block of a round must be resident,
 * or the flow-control gate would never open. */
template <int WY>
__global__ void __launch_bounds__(32 * (WY + 1), (WY > 8) ? 1 : 2)
k_step2_tma_ws(Geo g, const __grid_constant__ TmaMaps maps, Fld b, double cH, double cE, Src2 src, Span sp, int stages)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[kTmaMaxStages];
    __shared__ __align__(8) unsigned long long empty[kTmaMaxStages];
    __shared__ double xh[2 * 32 * WY];
    __shared__ double xe[4 * 32 * WY];
    double *ring = reinterpret_cast<double *>(smem_raw);

    constexpr int BYE = 2 * WY;
    const int tiles = sp.tiles_x * sp.tiles_y;
    const int nb = (int)gridDim.x;
    const bool producer = threadIdx.y == WY;
    for (int round = 0; round * nb < tiles; ++round) {
        const int tile = round * nb + (int)blockIdx.x;
        if (tile >= tiles)
            break;
        const unsigned in_round = (unsigned)min(nb, tiles - round * nb);
        const int ty = tile / sp.tiles_x, tx = tile - ty * sp.tiles_x;
        const int X0 = tx * kS2TileX, Y0 = ty * (BYE - 3);
        __syncthreads(); /* the previous round is over for every thread, producer included */
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            for (int st = 0; st < stages; ++st) {
                if (round > 0) {
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tma::smem_u32(full + st)) : "memory");
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tma::smem_u32(empty + st)) : "memory");
                }
                tma::mbar_init(full + st, 1);
                tma::mbar_init(empty + st, 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        unsigned *prog = sp.progress + (size_t)round * sp.progress_stride;
        if (producer) {
            step2_produce<WY>(g, maps, stages, ring, full, empty, X0, Y0, sp.kl_begin, sp.kl_end, prog, in_round, sp.window);
            continue;
        }
        const bool interior = X0 - 2 >= 1 && X0 + 29 < g.I && Y0 - 2 >= 1 && Y0 - 2 + BYE - 1 < g.J;
        if (interior)
            step2_sweep<WY, false, true>(g, maps, b, cH, cE, src, stages, ring, xh, xe, full, X0, Y0, sp.kl_begin, sp.kl_end,
                                         empty, prog);
        else
            step2_sweep<WY, true, true>(g, maps, b, cH, cE, src, stages, ring, xh, xe, full, X0, Y0, sp.kl_begin, sp.kl_end,
                                        empty, prog);
    }
}

} /* namespace fdtd */
