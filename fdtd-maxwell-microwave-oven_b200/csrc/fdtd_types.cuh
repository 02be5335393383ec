/*
 * fdtd_types.cuh -- geometry, argument structs and the arithmetic helper shared by every kernel
 * of the FDTD hot path (sm_100a).
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
 *        kl = k - k0 + 1.  Plane 0 receives the lower neighbour's halo (Hx, Hy; for the fused step
 *        also Ex, Ey, Ez), plane nk + 1 the upper neighbour's Ex/Ey halo -- or, on the last slab, it
 *        IS the global node plane K.
 *   Padding (columns beyond an array's extent, the unused row/plane) is zero and never stored to.
 *
 * Arithmetic.  Every update is evaluated with explicit round-to-nearest intrinsics in the
 * reference's operand order (main.c:448-461, 486-499):
 *     F = F + c * ((a - b) - (d - e))
 * so no FMA can be formed whatever the compiler flags; the build also passes -fmad=false.
 */
#pragma once

#include <cuda.h>
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
 * on == 0 in validation mode and on slabs that never touch the global plane k = 0; kl = local index of
 * that plane: 1 on the slab that owns it, 0 on a slab that starts at k = 1 (the fused step recomputes
 * H of the plane below its slab, which then is the source plane). */
struct Src {
    int on;
    int kl;
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

struct Span {
    int kl_begin, kl_end; /* local planes [kl_begin, kl_end) handled by this launch */
    int kchunk;           /* planes per block */
    int prefetch;         /* planes ahead to pull into L2 (0 = off) */
    int band;             /* two-step kernel: tiles are handed out column by column inside bands of this many tile rows */
    /* persistent form of the two-step kernel (k_step2_tma_ws): tiles per row / column of the cavity, the
     * planes-completed counters (one row of counters per round of tiles) and how many planes a block may
     * run ahead of the slowest block of its round */
    int tiles_x, tiles_y;
    unsigned *progress;
    int progress_stride;
    int window;
    /* rolling (in-place) form of the two-step kernel: the arrays are rings of zmod plane slots; local plane
     * kl is read from slot (kl + 1 + zrot_in) mod zmod and written to slot (kl + 1 + zrot_out) mod zmod.
     * zmod == 0: the usual two buffer sets, plane kl at slot kl + 1 of each. */
    int zmod, zrot_in, zrot_out;
};

/* non-blocking hint: bring the line holding p into L2 (no register, no scoreboard entry) */
__device__ __forceinline__ void prefetch_l2(const double *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

/* resident blocks per SM the register budget is capped for: 256 threads x 2 blocks x 128 registers
 * fill the register file for TY = 4; shorter strips need fewer registers and fit more blocks */
struct DenseView {
    int w, h, np;
    long long kd0;
};

/* value = 2 * u - 1, u = top 53 bits of splitmix64(seed ^ array<<58 ^ dense index) / 2^53 */
/* tensor maps of the buffer set the TMA-staged fused step reads (fdtd_fused_tma.cuh) */
struct TmaMaps {
    CUtensorMap m[6]; /* ex ey ez hx hy hz */
};

/* shared-memory footprint of one box, padded so that every box starts on a 128-byte boundary */
__host__ __device__ inline int tma_box_doubles(int bx, int by)
{
    const int n = (bx + 4) * (by + 2);
    return (n + 15) / 16 * 16;
}

constexpr int kTmaMaxStages = 8;

} /* namespace fdtd */
