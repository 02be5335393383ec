/*
 * fdtd_ctx.cu -- device context, transfers, kernel launches, halo exchange and stepping: the core of
 * libfdtd_b200.so's C ABI (include/fdtd_b200.h).  Dumps: fdtd_dump.cu; diagnostics: fdtd_diag.cu;
 * all slabs from one thread: fdtd_group.cu; shared declarations: fdtd_ctx.hpp.
 *
 * One context = one z-slab of the cavity resident in the HBM of one B200: the six field arrays
 * in the pitched layout described in fdtd_types.cuh, a compute stream, a halo stream with its
 * NCCL communicator, and a dump stream with pinned staging.  The host control thread only
 * queues work; nothing on the stepping path synchronises with the device.
 */
#include "fdtd_ctx.hpp"
#include "fdtd_update.cuh"
#include <cctype>
#include <sys/syscall.h>
#include <unistd.h>
#include "fdtd_fused.cuh"
#include "fdtd_fused_tma.cuh"
#include "fdtd_step2_tma.cuh"

using namespace fdtdi;

namespace fdtdi {

int check_ctx(const fdtd_ctx *c, const char *who)
{
    if (!c) {
        fdtd_set_error("%s: context is NULL", who);
        return FDTD_E_ARG;
    }
    if (c->opt_kernel != 4 && c->opt_wx * c->opt_wy > (c->opt_kernel == 3 ? 16 : 8)) { /* a block is at most 256 threads (512 for the TMA kernel) */
        fdtd_set_error("%s: options warps_x (%ld) * warps_y (%ld) must be <= %d for kernel %ld", who, c->opt_wx, c->opt_wy,
                       c->opt_kernel == 3 ? 16 : 8, c->opt_kernel);
        return FDTD_E_ARG;
    }
    return FDTD_OK;
}

int check_solo(const fdtd_ctx *c, const char *who)
{
    FDTD_TRY(check_ctx(c, who));
    if (c->in_group && c->nranks > 1) {
        fdtd_set_error("%s: this context is a slab of an fdtd_group; use the fdtd_group_* call", who);
        return FDTD_E_STATE;
    }
    return FDTD_OK;
}

int use_device(const fdtd_ctx *c)
{
    CUDA_TRY(cudaSetDevice(c->device));
    return FDTD_OK;
}

double *field_ptr(const fdtd_ctx *c, int idx)
{
    return c->base + (size_t)idx * c->array_elems;
}


DenseShape dense_shape(const fdtd_params &p, int idx)
{
    const size_t I = p.maxi, J = p.maxj, K = p.maxk;
    switch (idx) {
    case 0: return {I, J + 1, K + 1, true};      /* Ex */
    case 1: return {I + 1, J, K + 1, true};      /* Ey */
    case 2: return {I + 1, J + 1, K, false};     /* Ez */
    case 3: return {I + 1, J, K, false};         /* Hx */
    case 4: return {I, J + 1, K, false};         /* Hy */
    default: return {I, J, K + 1, true};         /* Hz */
    }
}

/* copy local planes [kl0, kl1) (clipped to the planes this slab owns of array idx) between the dense
 * host array -- host_first_owned_plane addresses local plane 1 -- and the pitched device array */
int copy_planes(fdtd_ctx *c, int idx, double *host_first_owned_plane, int kl0, int kl1, bool to_device, cudaStream_t st)
{
    const DenseShape s = dense_shape(c->p, idx);
    const int last = c->g.nk + ((s.node_planes && c->g.top) ? 1 : 0); /* last owned local plane */
    kl0 = std::max(kl0, 1);
    kl1 = std::min(kl1, last + 1);
    if (kl1 <= kl0 || s.w == 0 || s.h == 0)
        return FDTD_OK;
    double *dev = field_ptr(c, idx) + (size_t)c->g.PR * (size_t)kl0;
    double *hst = host_first_owned_plane + (size_t)(kl0 - 1) * s.w * s.h;
    cudaMemcpy3DParms m;
    memset(&m, 0, sizeof m);
    cudaPitchedPtr hp = make_cudaPitchedPtr(hst, s.w * sizeof(double), s.w * sizeof(double), s.h);
    cudaPitchedPtr dp = make_cudaPitchedPtr(dev, (size_t)c->g.P * sizeof(double),
                                            (size_t)c->g.P * sizeof(double), (size_t)c->g.R);
    m.srcPtr = to_device ? hp : dp;
    m.dstPtr = to_device ? dp : hp;
    m.extent = make_cudaExtent(s.w * sizeof(double), s.h, (size_t)(kl1 - kl0));
    m.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    CUDA_TRY(cudaMemcpy3DAsync(&m, st));
    return FDTD_OK;
}

/* all the planes this slab owns */
int copy_field(fdtd_ctx *c, int idx, double *host_first_owned_plane, bool to_device)
{
    return copy_planes(c, idx, host_first_owned_plane, 1, c->g.nk + 2, to_device, c->s_main);
}

Src make_src(const fdtd_ctx *c, const double *row)
{
    Src s;
    s.on = c->src_staged ? 1 : 0;
    s.kl = c->src_plane;
    s.i0 = (int)c->plan.i0;
    s.i1 = (int)c->plan.i1;
    s.j0 = (int)c->plan.j0;
    s.j1 = (int)c->plan.j1;
    s.n = c->src_n;
    s.vals = row;
    return s;
}

Src no_src()
{
    Src s;
    memset(&s, 0, sizeof s);
    return s;
}

/* ---- launches ------------------------------------------------------------------------------ */

template <int TY>
void launch_h_march_t(const fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    int wx = (int)c->opt_wx, wy = (int)c->opt_wy;
    if (wx * wy > 8) { /* the block shape belongs to another kernel (TMA kernels go up to 16 warps) */
        wx = 2;
        wy = 4;
    }
    Span sp{kl_begin, kl_end, (int)c->opt_kchunk, (int)c->opt_prefetch};
    dim3 block(32 * wx, wy);
    dim3 grid((c->g.I + 1 + block.x - 1) / block.x, (c->g.J + 1 + wy * TY - 1) / (wy * TY),
              (kl_end - kl_begin + sp.kchunk - 1) / sp.kchunk);
    k_update_h_march<TY><<<grid, block, 0, st>>>(c->g, c->f, c->ch, s, sp);
    ++c->launches;
}

template <int TY>
void launch_e_march_t(const fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    int wx = (int)c->opt_wx, wy = (int)c->opt_wy;
    if (wx * wy > 8) { /* the block shape belongs to another kernel (TMA kernels go up to 16 warps) */
        wx = 2;
        wy = 4;
    }
    Span sp{kl_begin, kl_end, (int)c->opt_kchunk, (int)c->opt_prefetch};
    dim3 block(32 * wx, wy);
    dim3 grid((c->g.I + 1 + block.x - 1) / block.x, (c->g.J + 1 + wy * TY - 1) / (wy * TY),
              (kl_end - kl_begin + sp.kchunk - 1) / sp.kchunk);
    k_update_e_march<TY><<<grid, block, 0, st>>>(c->g, c->f, c->ce, s, sp);
    ++c->launches;
}

/* H update of the local planes [kl_begin, kl_end) */
void launch_h(const fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    if (kl_end <= kl_begin)
        return;
    if (c->opt_kernel == 0) {
        dim3 block(64, 4);
        dim3 grid((c->g.I + 1 + 63) / 64, (c->g.J + 1 + 3) / 4, kl_end - kl_begin);
        Geo g = c->g;
        Fld f = c->f;
        /* the cell kernel numbers planes from blockIdx.z + 1: shift the base pointers instead */
        const long long shift = (long long)(kl_begin - 1) * g.PR;
        f.ex += shift; f.ey += shift; f.ez += shift; f.hx += shift; f.hy += shift; f.hz += shift;
        g.nk -= (kl_begin - 1);
        g.kbase += (kl_begin - 1);
        k_update_h_cell<<<grid, block, 0, st>>>(g, f, c->ch);
        ++c->launches;
        return;
    }
    switch (c->opt_strip) {
    case 1: launch_h_march_t<1>(c, s, kl_begin, kl_end, st); break;
    case 4: launch_h_march_t<4>(c, s, kl_begin, kl_end, st); break;
    default: launch_h_march_t<2>(c, s, kl_begin, kl_end, st); break;
    }
}

void launch_e(const fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    if (kl_end <= kl_begin)
        return;
    if (c->opt_kernel == 0) {
        dim3 block(64, 4);
        dim3 grid((c->g.I + 1 + 63) / 64, (c->g.J + 1 + 3) / 4, kl_end - kl_begin);
        Geo g = c->g;
        Fld f = c->f;
        const long long shift = (long long)(kl_begin - 1) * g.PR;
        f.ex += shift; f.ey += shift; f.ez += shift; f.hx += shift; f.hy += shift; f.hz += shift;
        g.nk -= (kl_begin - 1);
        g.kbase += (kl_begin - 1);
        k_update_e_cell<<<grid, block, 0, st>>>(g, f, c->ce);
        ++c->launches;
        return;
    }
    switch (c->opt_strip) {
    case 1: launch_e_march_t<1>(c, s, kl_begin, kl_end, st); break;
    case 4: launch_e_march_t<4>(c, s, kl_begin, kl_end, st); break;
    default: launch_e_march_t<2>(c, s, kl_begin, kl_end, st); break;
    }
}

/* Guard margins of a state allocation, in doubles.  The fused kernel loads without predicates
 * (fdtd_fused.cuh): a thread of an edge block may read up to one row and one column before the
 * first array and up to two planes and a few rows past the last one.  Those values are never
 * used for a stored result; the margins only keep the addresses inside the allocation. */
/* (the front margin also holds the extra plane below array 0, see create_impl) */
size_t margin_front(const fdtd_ctx *c) { return (size_t)c->g.P + 64 + (size_t)c->g.PR; }
/* back: the unpredicated kernel (k_step_fused) reads, from an edge block of the last plane, up to
 * one plane plus a block's rows (at most 8 warps x 4 rows) plus a block's columns (256) further */
size_t margin_back(const fdtd_ctx *c) { return 2 * (size_t)c->g.PR + 40 * (size_t)c->g.P + 2048; }
size_t margin_front_of(const fdtd_ctx *c) { return margin_front(c); }

/* one zero-filled set of six arrays (padding must be zero and stays zero) */
cudaError_t alloc_state(fdtd_ctx *c, double **raw, double **base)
{
    const size_t n = margin_front(c) + 6 * c->array_elems + margin_back(c);
    cudaError_t e = cudaMalloc((void **)raw, n * sizeof(double));
    if (e != cudaSuccess) {
        *raw = nullptr;
        return e;
    }
    *base = *raw + margin_front(c);
    return cudaMemsetAsync(*raw, 0, n * sizeof(double), c->s_main);
}

/* second buffer set of the fused step */
int ensure_pong(fdtd_ctx *c)
{
    if (c->base2)
        return FDTD_OK;
    cudaError_t e = alloc_state(c, &c->raw2, &c->base2);
    if (e != cudaSuccess) {
        c->base2 = nullptr;
        fdtd_set_error("fused step needs a second copy of the state (%zu bytes): %s",
                       6 * c->array_elems * sizeof(double), cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FDTD_E_NOMEM : FDTD_E_CUDA;
    }
    double **p2[6] = {&c->f2.ex, &c->f2.ey, &c->f2.ez, &c->f2.hx, &c->f2.hy, &c->f2.hz};
    for (int a = 0; a < 6; ++a)
        *p2[a] = c->base2 + (size_t)a * c->array_elems;
    return FDTD_OK;
}

void swap_buffers(fdtd_ctx *c)
{
    c->flip ^= 1;
    std::swap(c->raw, c->raw2);
    std::swap(c->base, c->base2);
    std::swap(c->f, c->f2);
}

template <int TY>
void launch_fused_t(const fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    /* 128 threads per block; a chunk shorter than 2 planes would put the source plane into a
     * chunk's "H only" prologue for no gain */
    int wx = (int)c->opt_wx, wy = (int)c->opt_wy;
    while (wx * wy > 4) {
        if (wy > 1) wy /= 2; else wx /= 2;
    }
    Span sp{kl_begin, kl_end, (int)std::max(c->opt_kchunk, 2L), (int)c->opt_prefetch};
    dim3 block(32 * wx, wy);
    dim3 grid((c->g.I + 1 + block.x - 1) / block.x, (c->g.J + 1 + wy * TY - 1) / (wy * TY),
              (kl_end - kl_begin + sp.kchunk - 1) / sp.kchunk);
    k_step_fused<TY><<<grid, block, 0, st>>>(c->g, c->f, c->f2, c->ch, c->ce, s, sp);
    ++c->launches;
}

/* cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda) */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int ring_slots(const fdtd_ctx *c) { return c->g.planes + 2 + kRollGap; }

/* mode 0: planes 0 .. nk+1;  1 ("wide"): planes -1 .. nk+2;  2: the whole ring of slots (current set only) */
int encode_maps(fdtd_ctx *c, int bx, int by, int mode = 0)
{
    const bool wide = mode >= 1;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            fdtd_set_error("cuTensorMapEncodeTiled is not available in this driver");
            return FDTD_E_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    double *sets[2] = {c->base, c->base2};
    const bool same_shape = c->tma_bx == bx && c->tma_by == by && c->tma_promo == c->opt_l2promo && c->tma_mode == mode;
    if (same_shape && c->tma_base[0] == sets[0] && c->tma_base[1] == sets[1])
        return FDTD_OK;
    if (same_shape && c->tma_base[0] == sets[1] && c->tma_base[1] == sets[0]) {
        std::swap(c->tma_maps[0], c->tma_maps[1]); /* the two states swapped roles */
        std::swap(c->tma_base[0], c->tma_base[1]);
        return FDTD_OK;
    }
    /* wide: the tensor starts at the spare plane below plane 0, so local plane kl is z = kl + 1 */
    const cuuint64_t dims[3] = {(cuuint64_t)c->g.P, (cuuint64_t)c->g.R,
                                (cuuint64_t)(mode == 2 ? ring_slots(c) : c->g.planes + (wide ? 2 : 0))};
    const size_t shift = wide ? (size_t)c->g.PR : 0;
    const cuuint64_t strides[2] = {(cuuint64_t)c->g.P * 8, (cuuint64_t)c->g.PR * 8};
    const cuuint32_t box[3] = {(cuuint32_t)(bx + 4), (cuuint32_t)(by + 2), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapL2promotion promo = c->opt_l2promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : c->opt_l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : c->opt_l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                               : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    for (int set = 0; set < 2; ++set)
        for (int a = 0; a < 6 && sets[set]; ++a) {
            CUresult r = encode(&c->tma_maps[set].m[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                                sets[set] + (size_t)a * c->array_elems - shift, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                fdtd_set_error("cuTensorMapEncodeTiled failed with CUresult %d (box %d x %d)", (int)r, bx + 2, by + 2);
                return FDTD_E_CUDA;
            }
        }
    c->tma_bx = bx;
    c->tma_by = by;
    c->tma_promo = (int)c->opt_l2promo;
    c->tma_mode = mode;
    c->tma_base[0] = sets[0];
    c->tma_base[1] = sets[1];
    return FDTD_OK;
}

template <int TY, int CWX, int CWY>
int launch_fused_tma_t(fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    const int wx = (int)c->opt_wx, wy = (int)c->opt_wy;
    const int bx = 32 * wx, by = wy * TY;
    FDTD_TRY(encode_maps(c, bx, by));
    const int stages = (int)c->opt_stages;
    const size_t smem = (size_t)stages * 6 * tma_box_doubles(bx, by) * sizeof(double);
    if (smem > 226 * 1024) {
        fdtd_set_error("TMA ring of %d stages x %d x %d tile needs %zu bytes of shared memory", stages, bx, by, smem);
        return FDTD_E_ARG;
    }
    /* the opt-in to more than 48 KB of dynamic shared memory is per device: remember it per context */
    const void *fn = (const void *)k_step_fused_tma<TY, CWX, CWY>;
    bool configured = false;
    for (int k = 0; k < c->n_smem_optin; ++k)
        configured = configured || c->smem_optin[k] == fn;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_step_fused_tma<TY, CWX, CWY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      226 * 1024));
        if (c->n_smem_optin < 16)
            c->smem_optin[c->n_smem_optin++] = fn;
    }
    Span sp{kl_begin, kl_end, (int)std::max(c->opt_kchunk, 2L), 0};
    dim3 block(bx, wy);
    dim3 grid((c->g.I + 1 + bx - 1) / bx, (c->g.J + 1 + by - 1) / by, (kl_end - kl_begin + sp.kchunk - 1) / sp.kchunk);
    k_step_fused_tma<TY, CWX, CWY><<<grid, block, smem, st>>>(c->g, c->tma_maps[0], c->f, c->f2, c->ch, c->ce, s, sp, stages);
    ++c->launches;
    return FDTD_OK;
}

/* the launch shapes that won the sweeps get a kernel with the shape baked in; any other shape runs
 * the generic instantiation */
int launch_fused_tma(fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    if (c->opt_kernel == 4) { /* a single step on a context that otherwise takes two per sweep: the default shape */
        const long keep_wx = c->opt_wx, keep_wy = c->opt_wy, keep_st = c->opt_stages, keep_kc = c->opt_kchunk;
        c->opt_wx = 1; c->opt_wy = 8; c->opt_stages = 4; c->opt_kchunk = 32;
        const int rc = launch_fused_tma_t<1, 1, 8>(c, s, kl_begin, kl_end, st);
        c->opt_wx = keep_wx; c->opt_wy = keep_wy; c->opt_stages = keep_st; c->opt_kchunk = keep_kc;
        return rc;
    }
    const long ty = c->opt_strip == 1 ? 1 : 2, wx = c->opt_wx, wy = c->opt_wy;
    if (ty == 2 && wx == 4 && wy == 2) return launch_fused_tma_t<2, 4, 2>(c, s, kl_begin, kl_end, st);
    if (ty == 2 && wx == 2 && wy == 2) return launch_fused_tma_t<2, 2, 2>(c, s, kl_begin, kl_end, st);
    if (ty == 2 && wx == 2 && wy == 4) return launch_fused_tma_t<2, 2, 4>(c, s, kl_begin, kl_end, st);
    if (ty == 2 && wx == 4 && wy == 4) return launch_fused_tma_t<2, 4, 4>(c, s, kl_begin, kl_end, st); /* 512 threads */
    if (ty == 2 && wx == 2 && wy == 8) return launch_fused_tma_t<2, 2, 8>(c, s, kl_begin, kl_end, st); /* 512 threads */
    if (wx * wy > 8) {
        fdtd_set_error("TMA kernel: %ld x %ld warps has no 512-thread instantiation (use 4x4 or 2x8 with strip 2)", wx, wy);
        return FDTD_E_ARG;
    }
    if (ty == 1 && wx == 1 && wy == 8) return launch_fused_tma_t<1, 1, 8>(c, s, kl_begin, kl_end, st);
    if (ty == 2) return launch_fused_tma_t<2, 0, 0>(c, s, kl_begin, kl_end, st);
    return launch_fused_tma_t<1, 0, 0>(c, s, kl_begin, kl_end, st);
}

/* one whole step (H then E) of the local planes [kl_begin, kl_end): reads c->f, writes c->f2 */
void launch_fused(fdtd_ctx *c, const Src &s, int kl_begin, int kl_end, cudaStream_t st)
{
    if (kl_end <= kl_begin)
        return;
    if (c->opt_kernel >= 3) {
        const int rc = launch_fused_tma(c, s, kl_begin, kl_end, st);
        if (rc != FDTD_OK)
            c->launch_error = rc;
        return;
    }
    switch (c->opt_strip) {
    case 1: launch_fused_t<1>(c, s, kl_begin, kl_end, st); break;
    case 3: launch_fused_t<3>(c, s, kl_begin, kl_end, st); break;
    case 4: launch_fused_t<4>(c, s, kl_begin, kl_end, st); break;
    default: launch_fused_t<2>(c, s, kl_begin, kl_end, st); break;
    }
}

/* two whole steps of the local planes [kl_begin, kl_end) in one sweep: reads c->f, writes c->f2 */
template <int WY>
int launch_step2_t(fdtd_ctx *c, const Src &s1, const Src &s2, int kl_begin, int kl_end, cudaStream_t st)
{
    constexpr int BYE = 2 * WY;
    FDTD_TRY(encode_maps(c, kS2BoxW - 4, BYE - 1, c->rolling ? 2 : 1));
    /* as many stages as asked for, as far as the shared memory of one SM goes (at least two) */
    const size_t stage_bytes = (size_t)6 * tma_box_doubles(kS2BoxW - 4, BYE - 1) * sizeof(double);
    const int stages = (int)std::max(2L, std::min(c->opt_stages, (long)(200 * 1024 / stage_bytes)));
    const size_t smem = (size_t)stages * stage_bytes;
    const void *fn = (const void *)k_step2_tma<WY>;
    bool configured = false;
    for (int k = 0; k < c->n_smem_optin; ++k)
        configured = configured || c->smem_optin[k] == fn;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_step2_tma<WY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (c->n_smem_optin < 16)
            c->smem_optin[c->n_smem_optin++] = fn;
    }
    Span sp{kl_begin, kl_end, (int)std::max(c->opt_kchunk, 2L), 0, (int)c->opt_band};
    dim3 block(32, WY);
    dim3 grid((c->g.I + 1 + kS2TileX - 1) / kS2TileX, (c->g.J + 1 + BYE - 4) / (BYE - 3),
              (kl_end - kl_begin + sp.kchunk - 1) / sp.kchunk);
    Src2 src{s1, s2};
    if (c->rolling) {
        /* In place on the ring: plane k of the new state goes D slots below where plane k of the old one
         * sits.  A chunk only ever overwrites slots whose planes the chunks before it have finished with
         * (D >= planes per chunk + its run-in), so the chunks run one after another, bottom to top. */
        const int Z = ring_slots(c);
        const int kc = (int)std::min((long)sp.kchunk, (long)kRollGap - 4), D = kc + 4;
        sp.kchunk = kc;
        sp.zmod = Z;
        sp.zrot_in = c->roll_rot;
        sp.zrot_out = ((c->roll_rot - D) % Z + Z) % Z;
        grid.z = 1;
        for (int a = kl_begin; a < kl_end; a += kc) {
            sp.kl_begin = a;
            sp.kl_end = std::min(a + kc, kl_end);
            k_step2_tma<WY><<<grid, block, smem, st>>>(c->g, c->tma_maps[0], c->f, c->ch, c->ce, src, sp, stages);
            ++c->launches;
        }
        return FDTD_OK;
    }
    const unsigned cx = (unsigned)c->opt_cluster_x, cy = (unsigned)c->opt_cluster_y;
    if (cx * cy > 8) {
        fdtd_set_error("two-step kernel: cluster_x (%u) * cluster_y (%u) must be <= 8", cx, cy);
        return FDTD_E_ARG;
    }
    if (cx * cy > 1) {
        /* Thread-block clusters: the blocks of a cluster are co-scheduled on one GPC and start together,
         * so neighbouring tiles sweep z in step and L2 serves the halo they share.  The grid is padded to
         * whole clusters; blocks whose tile lies outside the cavity leave at once. */
        grid.x = (grid.x + cx - 1) / cx * cx;
        grid.y = (grid.y + cy - 1) / cy * cy;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cx;
        attr.val.clusterDim.y = cy;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, k_step2_tma<WY>, c->g, c->tma_maps[0], c->f2, c->ch, c->ce, src, sp, stages));
    } else {
        k_step2_tma<WY><<<grid, block, smem, st>>>(c->g, c->tma_maps[0], c->f2, c->ch, c->ce, src, sp, stages);
    }
    ++c->launches;
    return FDTD_OK;
}

/* the persistent, warp-specialised form (k_step2_tma_ws): WYC compute warps + one producer warp per
 * block, as many blocks as the GPU holds at once, rounds of tiles, planes-completed counters */
template <int WYC>
int launch_step2_ws_t(fdtd_ctx *c, const Src &s1, const Src &s2, int kl_begin, int kl_end, cudaStream_t st)
{
    constexpr int BYE = 2 * WYC;
    FDTD_TRY(encode_maps(c, kS2BoxW - 4, BYE - 1, true));
    const size_t stage_bytes = (size_t)6 * tma_box_doubles(kS2BoxW - 4, BYE - 1) * sizeof(double);
    const int stages = (int)std::max(2L, std::min(c->opt_stages, (long)(200 * 1024 / stage_bytes)));
    const size_t smem = (size_t)stages * stage_bytes;
    const void *fn = (const void *)k_step2_tma_ws<WYC>;
    bool configured = false;
    for (int k = 0; k < c->n_smem_optin; ++k)
        configured = configured || c->smem_optin[k] == fn;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_step2_tma_ws<WYC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (c->n_smem_optin < 16)
            c->smem_optin[c->n_smem_optin++] = fn;
    }
    dim3 block(32, WYC + 1);
    const int tiles_x = (c->g.I + 1 + kS2TileX - 1) / kS2TileX, tiles_y = (c->g.J + 1 + BYE - 4) / (BYE - 3);
    int per_sm = 0, sms = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_step2_tma_ws<WYC>, 32 * (WYC + 1), smem));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int tiles = tiles_x * tiles_y;
    int nb = std::min(tiles, std::max(1, per_sm * sms));
    if (nb > tiles_x) /* whole rows of tiles per round: the x overlap stays inside a round */
        nb = nb / tiles_x * tiles_x;
    const int rounds = (tiles + nb - 1) / nb;
    const int planes = kl_end - kl_begin + 6;
    const size_t need = (size_t)rounds * planes;
    if (c->progress_elems < need) {
        if (c->progress_dev) cudaFree(c->progress_dev);
        c->progress_dev = nullptr;
        c->progress_elems = 0;
        CUDA_TRY(cudaMalloc((void **)&c->progress_dev, need * sizeof(unsigned)));
        c->progress_elems = need;
    }
    CUDA_TRY(cudaMemsetAsync(c->progress_dev, 0, need * sizeof(unsigned), st));
    Span sp{kl_begin, kl_end, kl_end - kl_begin, 0, 1};
    sp.tiles_x = tiles_x;
    sp.tiles_y = tiles_y;
    sp.progress = c->progress_dev;
    sp.progress_stride = planes;
    sp.window = (int)c->opt_window;
    Geo g = c->g;
    TmaMaps maps = c->tma_maps[0];
    Fld out = c->f2;
    double ch = c->ch, ce = c->ce;
    Src2 src{s1, s2};
    int stg = stages;
    void *args[] = {&g, &maps, &out, &ch, &ce, &src, &sp, &stg};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(nb), block, args, smem, st));
    ++c->launches;
    return FDTD_OK;
}

int launch_step2(fdtd_ctx *c, const Src &s1, const Src &s2, int kl_begin, int kl_end, cudaStream_t st)
{
    if (kl_end <= kl_begin)
        return FDTD_OK;
    int rc;
    if (c->opt_persistent && !c->rolling)
        rc = c->opt_wy > 8 ? launch_step2_ws_t<15>(c, s1, s2, kl_begin, kl_end, st)
                           : launch_step2_ws_t<7>(c, s1, s2, kl_begin, kl_end, st);
    else if (c->opt_wy >= 16)
        rc = launch_step2_t<16>(c, s1, s2, kl_begin, kl_end, st);
    else if (c->opt_wy >= 12)
        rc = launch_step2_t<12>(c, s1, s2, kl_begin, kl_end, st);
    else
        rc = launch_step2_t<8>(c, s1, s2, kl_begin, kl_end, st);
    return rc;
}

void launch_set_source(const fdtd_ctx *c, const double *row_dev, cudaStream_t st)
{
    const Src s = make_src(c, row_dev);
    dim3 block(32, 8);
    dim3 grid((s.i1 - s.i0 + 31) / 32, (s.j1 - s.j0 + 7) / 8);
    k_set_source<<<grid, block, 0, st>>>(c->g, c->f, s);
    ++c->launches;
}

/* ---- the rolling window (in-place two-step sweeps) ----------------------------------------------- */

static int roll_shift(const fdtd_ctx *c)
{
    return (int)std::min(std::max(c->opt_kchunk, 2L), (long)kRollGap - 4) + 4;
}

/* after a sweep: the state sits roll_shift() slots lower in every ring of the run (mine and, in
 * lockstep, my neighbours') */
static void roll_advance(fdtd_ctx *c)
{
    const int D = roll_shift(c), Z = ring_slots(c);
    c->roll_rot = ((c->roll_rot - D) % Z + Z) % Z;
    if (c->roll_peer_z_lo)
        c->roll_peer_rot_lo = ((c->roll_peer_rot_lo - D) % c->roll_peer_z_lo + c->roll_peer_z_lo) % c->roll_peer_z_lo;
    if (c->roll_peer_z_hi)
        c->roll_peer_rot_hi = ((c->roll_peer_rot_hi - D) % c->roll_peer_z_hi + c->roll_peer_z_hi) % c->roll_peer_z_hi;
}

/* Rotate every ring back so that local plane -1 is in slot 0 again -- the layout everything outside
 * the stepping loop expects.  Each slot moves once (cycle by cycle, one spare plane). */
int roll_canonicalise(fdtd_ctx *c)
{
    if (!c->rolling || c->roll_rot == 0)
        return FDTD_OK;
    FDTD_TRY(wait_halos(c)); /* halo planes still arriving are part of the rings */
    const int Z = ring_slots(c), r = c->roll_rot;
    const size_t PR = (size_t)c->g.PR, bytes = PR * sizeof(double);
    if (!c->roll_tmp)
        CUDA_TRY(cudaMalloc((void **)&c->roll_tmp, bytes));
    int g = Z, t = r;
    while (t) {
        const int u = g % t;
        g = t;
        t = u;
    }
    for (int a = 0; a < 6; ++a) {
        double *slot0 = field_ptr(c, a) - PR;
        for (int c0 = 0; c0 < g; ++c0) {
            CUDA_TRY(cudaMemcpyAsync(c->roll_tmp, slot0 + (size_t)c0 * PR, bytes, cudaMemcpyDeviceToDevice, c->s_main));
            int j = c0;
            for (;;) {
                const int src = (j + r) % Z;
                if (src == c0)
                    break;
                CUDA_TRY(cudaMemcpyAsync(slot0 + (size_t)j * PR, slot0 + (size_t)src * PR, bytes, cudaMemcpyDeviceToDevice,
                                         c->s_main));
                j = src;
            }
            CUDA_TRY(cudaMemcpyAsync(slot0 + (size_t)j * PR, c->roll_tmp, bytes, cudaMemcpyDeviceToDevice, c->s_main));
        }
    }
    c->roll_rot = c->roll_peer_rot_lo = c->roll_peer_rot_hi = 0; /* every slab of the run does this at the same point */
    if (c->nranks > 1 && c->transport != TR_NCCL) {
        /* a neighbour that pushes into (or pulls from) this slab must find the canonical layout: mark the
         * point for the event transport; the flag transport orders through its acknowledgements */
        CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
    }
    return FDTD_OK;
}

static void announce_rolling(fdtd_ctx *c)
{
    if (!c->roll_announced && c->rank == 0 && !c->opt_rolling)
        fprintf(stderr, "[fdtd_b200] a second copy of the state (%.1f GB per slab) does not fit in HBM on every slab: the "
                        "two-step kernel works in place on a rolling window of plane slots (slower by its chunk-by-chunk "
                        "launches, same results)\n", 6e-9 * (double)c->array_elems * sizeof(double));
    c->roll_announced = true;
    c->rolling = true;
}

/* A time step is made of segments -- one for the fused kernels (H and E in one sweep), two for the
 * split kernels (H half-step, E half-step).  seg_launch queues a segment's kernels on the compute
 * stream: wait for the halos of the previous segment, boundary plane(s) first, the ev_bnd event,
 * then the interior planes; exchange_many(seg_xchg(seg), on_comm = true) then moves the boundary
 * planes on the halo streams while the interior planes run (fdtd_halo.cu).  The two are separate so
 * that one host thread can drive every slab of a group: all launches, then all transfers. */
int seg_launch(fdtd_ctx *c, const Src &s, Segment seg, const Src *second)
{
    FDTD_TRY(use_device(c));
    const int nk = c->g.nk;
    const int h_end = nk + c->g.top + 1; /* exclusive */
    const bool multi = c->nranks > 1;
    const bool sends_up = c->rank + 1 < c->nranks, sends_down = c->rank > 0;
    if (multi)
        FDTD_TRY(wait_halos(c));
    if (seg != SEG_STEP2)
        c->wide_halo_valid = false; /* a single step refreshes one halo plane each way only */
    if (seg == SEG_STEP2) {
        /* two steps in one sweep: reads c->f (with two halo planes each way), writes c->f2, swap.
         * The two planes at either end travel: they go first. */
        if (c->rolling) {
            /* in place on the ring, chunk after chunk from the bottom: nothing to reorder, the planes
             * that travel up are final only at the end of the sweep */
            FDTD_TRY(launch_step2(c, s, *second, 1, h_end, c->s_main));
            if (multi)
                CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
            roll_advance(c);
        } else if (!multi) {
            FDTD_TRY(launch_step2(c, s, *second, 1, h_end, c->s_main));
        } else if (nk < 6) {
            FDTD_TRY(launch_step2(c, s, *second, 1, h_end, c->s_main));
            CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
        } else {
            /* The chunk at either end of the slab goes first -- whole chunks, so that no plane is loaded
             * more often than in a single launch (a chunk brings its own two planes of run-in anyway);
             * the halo planes travel while the chunks in between run. */
            const int bc = (int)std::min((long)std::max(2, nk / 4), std::max(c->opt_kchunk, 2L));
            int lo = 1, hi = h_end;
            if (sends_up) {
                FDTD_TRY(launch_step2(c, s, *second, h_end - bc, h_end, c->s_main));
                hi = h_end - bc;
            }
            if (sends_down) {
                FDTD_TRY(launch_step2(c, s, *second, 1, 1 + bc, c->s_main));
                lo = 1 + bc;
            }
            CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
            FDTD_TRY(launch_step2(c, s, *second, lo, hi, c->s_main));
        }
        if (!c->rolling)
            swap_buffers(c);
        c->e_halo_valid = c->h_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = true; /* after the exchange */
    } else if (seg == SEG_FUSED) {
        /* reads c->f, writes c->f2, then the two swap */
        if (!multi) {
            launch_fused(c, s, 1, h_end, c->s_main);
            swap_buffers(c);
        } else {
            if (nk < 3) {
                launch_fused(c, s, 1, h_end, c->s_main);
                CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
            } else {
                int lo = 1, hi = h_end;
                if (sends_up) { /* top owned plane first: 5 arrays of it travel up */
                    launch_fused(c, s, nk, nk + 1, c->s_main);
                    hi = nk;
                }
                if (sends_down) { /* first owned plane: Ex, Ey travel down */
                    launch_fused(c, s, 1, 2, c->s_main);
                    lo = 2;
                }
                CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
                launch_fused(c, s, lo, hi, c->s_main);
            }
            swap_buffers(c); /* the exchange works on the new state */
        }
        if (c->launch_error != FDTD_OK) {
            const int rc = c->launch_error;
            c->launch_error = FDTD_OK;
            return rc;
        }
    } else if (seg == SEG_H) {
        if (c->opt_kernel == 0 && c->src_here)
            launch_set_source(c, s.vals, c->s_main);
        if (!multi) {
            launch_h(c, s, 1, h_end, c->s_main);
        } else if (sends_up) {
            launch_h(c, s, nk, nk + 1, c->s_main); /* boundary plane first */
            CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
            launch_h(c, s, 1, nk, c->s_main);
        } else {
            launch_h(c, s, 1, h_end, c->s_main);
            CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
        }
    } else {
        if (c->opt_kernel == 0 && c->src_here)
            launch_set_source(c, s.vals, c->s_main);
        if (!multi) {
            launch_e(c, s, 1, nk + 1, c->s_main);
        } else if (sends_down) {
            launch_e(c, s, 1, 2, c->s_main);
            CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
            launch_e(c, s, 2, nk + 1, c->s_main);
        } else {
            launch_e(c, s, 1, nk + 1, c->s_main);
            CUDA_TRY(cudaEventRecord(c->ev_bnd, c->s_main));
        }
        c->low_e_halo_valid = false; /* the split kernels refresh only Hx, Hy of plane 0 */
    }
    CUDA_TRY(cudaGetLastError());
    return FDTD_OK;
}

Xchg seg_xchg(Segment seg)
{
    Xchg x{};
    x.h = seg == SEG_FUSED || seg == SEG_H || seg == SEG_STEP2;
    x.h_with_e = seg == SEG_FUSED;
    x.e = seg == SEG_FUSED || seg == SEG_E || seg == SEG_STEP2;
    x.e_with_hz = false;
    x.wide = seg == SEG_STEP2;
    return x;
}

bool step2_usable(const fdtd_ctx *c)
{
    /* a slab must own the two planes it sends each way */
    return c->nranks == 1 || c->p.maxk / (size_t)c->nranks >= 2;
}

/* One pass of the loop body main.c:770-779 for one context (one process per GPU). */
int queue_step(fdtd_ctx *c, const Src &s, cudaEvent_t ev_h_begin, cudaEvent_t ev_mid, cudaEvent_t ev_e_end)
{
    if (ev_h_begin)
        CUDA_TRY(cudaEventRecord(ev_h_begin, c->s_main));
    if (c->opt_kernel >= 2) {
        FDTD_TRY(seg_launch(c, s, SEG_FUSED));
        FDTD_TRY(exchange_many(&c, 1, seg_xchg(SEG_FUSED), true));
        if (ev_mid)
            CUDA_TRY(cudaEventRecord(ev_mid, c->s_main));
    } else {
        FDTD_TRY(seg_launch(c, s, SEG_H));
        FDTD_TRY(exchange_many(&c, 1, seg_xchg(SEG_H), true));
        if (ev_mid)
            CUDA_TRY(cudaEventRecord(ev_mid, c->s_main));
        FDTD_TRY(seg_launch(c, s, SEG_E));
        FDTD_TRY(exchange_many(&c, 1, seg_xchg(SEG_E), true));
    }
    if (ev_e_end)
        CUDA_TRY(cudaEventRecord(ev_e_end, c->s_main));
    return FDTD_OK;
}

/* Fill and upload the source rows of steps [0, count) starting at time t (ring slot = step index
 * modulo kSrcRing).  Returns the time counter after `count` steps through *t_io. */
int stage_source_rows(fdtd_ctx *c, size_t count, double *t_io)
{
    const size_t row = 2 * (size_t)c->src_n;
    CUDA_TRY(cudaEventSynchronize(c->ev_src)); /* previous upload has left the pinned buffer */
    double t = *t_io;
    for (size_t s = 0; s < count; ++s, t += c->p.time_step) {
        if (c->src_staged)
            FDTD_TRY(fdtd_source_values(&c->p, &c->plan, t, c->src_host + s * row,
                                        c->src_host + s * row + c->src_n));
    }
    if (c->src_staged && row > 0) {
        CUDA_TRY(cudaMemcpyAsync(c->src_dev, c->src_host, count * row * sizeof(double),
                                 cudaMemcpyHostToDevice, c->s_main));
        CUDA_TRY(cudaEventRecord(c->ev_src, c->s_main));
    }
    *t_io = t;
    return FDTD_OK;
}

/* The fused step needs the state twice in HBM; when that does not fit and the caller did not ask
 * for a particular kernel, the in-place split kernels (144 B per cell-update instead of 96) take
 * over -- still on the GPU, and never silently: a line on stderr, and the options "fallback" and
 * "kernel" report it. */
void fall_back_to_split(fdtd_ctx *c)
{
    cudaGetLastError();
    if (!c->fallback && c->rank == 0)
        fprintf(stderr, "[fdtd_b200] the fused step needs a second copy of the state (%.1f GB per slab) which does not "
                        "fit in HBM on every slab: using the in-place split kernels (kernel=1) instead\n",
                6e-9 * (double)c->array_elems * sizeof(double));
    c->fallback = 1;
    c->opt_kernel = 1;
    c->opt_strip = 2;
    c->opt_kchunk = 8;
    c->opt_wx = 2;
    c->opt_wy = 2;
    if (c->raw2 && !c->n_ipc) {
        cudaFree(c->raw2);
        c->raw2 = c->base2 = nullptr;
    }
}

/* Decide, identically on every slab, whether the fused kernels can run.  A single context decides
 * alone; slabs of one process per GPU agreed when they were wired (fdtd_ctx_comm_init /
 * fdtd_ctx_peer_connect); a group decides in group_run. */
int settle_kernel(fdtd_ctx *c)
{
    if (c->opt_kernel < 2)
        return FDTD_OK;
    if (c->opt_kernel == 4 && c->opt_rolling && step2_usable(c)) {
        announce_rolling(c);
        return FDTD_OK;
    }
    c->rolling = false;
    if (c->nranks == 1) {
        const int rc = ensure_pong(c);
        if (rc == FDTD_E_NOMEM && c->opt_kernel == 4 && step2_usable(c)) {
            cudaGetLastError();
            announce_rolling(c);
            return FDTD_OK;
        }
        if (rc == FDTD_E_NOMEM && c->kernel_auto) {
            fall_back_to_split(c);
            return FDTD_OK;
        }
        return rc;
    }
    if (!c->wired) {
        fdtd_set_error("multi-rank context is not wired to its neighbours: call fdtd_ctx_comm_init or "
                       "fdtd_ctx_peer_connect first");
        return FDTD_E_STATE;
    }
    if (c->fused_ok)
        return ensure_pong(c); /* already there */
    if (c->opt_kernel == 4 && step2_usable(c)) { /* every rank sees the same fused_ok: the same decision */
        announce_rolling(c);
        return FDTD_OK;
    }
    if (c->kernel_auto) {
        fall_back_to_split(c);
        return FDTD_OK;
    }
    fdtd_set_error("the fused kernels (kernel >= 2) need a second copy of the state on every slab; it was not "
                   "available when the slabs were wired (select the kernel before wiring, or use kernel 0/1)");
    return FDTD_E_NOMEM;
}

int run_impl(fdtd_ctx *c, size_t steps, double *time_counter, float *total_ms, float *h_ms, float *e_ms)
{
    FDTD_TRY(use_device(c));
    FDTD_TRY(settle_kernel(c));
    if (!(c->opt_kernel == 4 && step2_usable(c) && steps >= 2)) /* (the two-step kernel refreshes its wider halos itself) */
        FDTD_TRY(refresh_halos_many(&c, 1, c->opt_kernel >= 2));
    const bool timed = total_ms != nullptr;
    const bool per_kernel = timed && (h_ms || e_ms);
    const size_t max_kernel_events = 2048;
    std::vector<cudaEvent_t> evs;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    if (timed) {
        CUDA_TRY(cudaEventCreate(&ev_begin));
        CUDA_TRY(cudaEventCreate(&ev_end));
        if (per_kernel) {
            const size_t n = std::min(steps, max_kernel_events);
            evs.resize(3 * n);
            for (auto &e : evs)
                CUDA_TRY(cudaEventCreate(&e));
        }
    }
    double t = *time_counter;
    const size_t row = 2 * (size_t)c->src_n;
    bool first = true;
    for (size_t done = 0; done < steps;) {
        const size_t chunk = std::min(steps - done, (size_t)kSrcRing);
        double t_chunk = t;
        FDTD_TRY(stage_source_rows(c, chunk, &t_chunk));
        if (first && timed) {
            CUDA_TRY(cudaStreamSynchronize(c->s_main));
            CUDA_TRY(cudaEventRecord(ev_begin, c->s_main));
        }
        first = false;
        for (size_t s = 0; s < chunk; ++s) {
            const Src src = make_src(c, c->src_dev + s * row);
            const size_t gs = done + s;
            const bool ev = per_kernel && gs * 3 + 2 < evs.size();
            if (c->opt_kernel == 4 && s + 1 < chunk && step2_usable(c)) {
                /* two steps in one sweep */
                const Src src2 = make_src(c, c->src_dev + (s + 1) * row);
                const bool ev2 = per_kernel && (gs + 1) * 3 + 2 < evs.size();
                FDTD_TRY(refresh_halos_many(&c, 1, true, true));
                if (ev)
                    CUDA_TRY(cudaEventRecord(evs[3 * gs], c->s_main));
                FDTD_TRY(seg_launch(c, src, SEG_STEP2, &src2));
                FDTD_TRY(exchange_many(&c, 1, seg_xchg(SEG_STEP2), true));
                CUDA_TRY(cudaGetLastError());
                if (ev) { /* the sweep is booked on the first of its two steps */
                    CUDA_TRY(cudaEventRecord(evs[3 * gs + 1], c->s_main));
                    CUDA_TRY(cudaEventRecord(evs[3 * gs + 2], c->s_main));
                }
                if (ev2)
                    for (int e = 0; e < 3; ++e)
                        CUDA_TRY(cudaEventRecord(evs[3 * (gs + 1) + e], c->s_main));
                ++s;
                continue;
            }
            if (c->rolling) {
                /* no second buffer set: the odd step runs with the in-place split kernels on the canonical layout */
                FDTD_TRY(roll_canonicalise(c));
                const long keep[5] = {c->opt_kernel, c->opt_strip, c->opt_kchunk, c->opt_wx, c->opt_wy};
                c->opt_kernel = 1; c->opt_strip = 2; c->opt_kchunk = 8; c->opt_wx = 2; c->opt_wy = 2;
                int rc1 = refresh_halos_many(&c, 1, false);
                if (rc1 == FDTD_OK)
                    rc1 = ev ? queue_step(c, src, evs[3 * gs], evs[3 * gs + 1], evs[3 * gs + 2])
                             : queue_step(c, src, nullptr, nullptr, nullptr);
                c->opt_kernel = keep[0]; c->opt_strip = keep[1]; c->opt_kchunk = keep[2]; c->opt_wx = keep[3]; c->opt_wy = keep[4];
                FDTD_TRY(rc1);
                continue;
            }
            if (c->opt_kernel == 4)
                FDTD_TRY(refresh_halos_many(&c, 1, true)); /* a single step after pairs, or slabs too thin for pairs */
            if (ev)
                FDTD_TRY(queue_step(c, src, evs[3 * gs], evs[3 * gs + 1], evs[3 * gs + 2]));
            else
                FDTD_TRY(queue_step(c, src, nullptr, nullptr, nullptr));
        }
        t = t_chunk;
        done += chunk;
    }
    *time_counter = t;
    FDTD_TRY(roll_canonicalise(c)); /* (rolling form only) everything else expects the canonical layout */
    if (timed) {
        CUDA_TRY(cudaEventRecord(ev_end, c->s_main));
        if (c->nranks > 1)
            CUDA_TRY(cudaStreamSynchronize(c->s_comm));
        CUDA_TRY(cudaStreamSynchronize(c->s_main));
        CUDA_TRY(cudaEventElapsedTime(total_ms, ev_begin, ev_end));
        if (per_kernel) {
            float hs = 0.f, es = 0.f;
            const size_t n = evs.size() / 3;
            for (size_t s = 0; s < n; ++s) {
                float a = 0.f, b = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&a, evs[3 * s], evs[3 * s + 1]));
                CUDA_TRY(cudaEventElapsedTime(&b, evs[3 * s + 1], evs[3 * s + 2]));
                hs += a;
                es += b;
            }
            /* scale to all steps when only the first max_kernel_events were instrumented */
            const float scale = n ? (float)steps / (float)n : 0.f;
            if (h_ms) *h_ms = hs * scale;
            if (e_ms) *e_ms = es * scale;
            for (auto &e : evs)
                cudaEventDestroy(e);
        }
        cudaEventDestroy(ev_begin);
        cudaEventDestroy(ev_end);
    }
    return FDTD_OK;
}

int create_impl(const fdtd_params *p, int device, int rank, int nranks, fdtd_ctx **out)
{
    if (!p || !out) {
        fdtd_set_error("fdtd_ctx_create: NULL argument");
        return FDTD_E_ARG;
    }
    if (p->maxi < 1 || p->maxj < 1 || p->maxk < 1 || p->maxi > 60000 || p->maxj > 60000 ||
        p->maxk > (size_t)1 << 30) {
        fdtd_set_error("fdtd_ctx_create: unsupported grid %zu x %zu x %zu", p->maxi, p->maxj, p->maxk);
        return FDTD_E_ARG;
    }
    if (p->mode != 0 && p->mode != 1) {
        fdtd_set_error("fdtd_ctx_create: mode must be 0 (validation) or 1 (computation), got %d", p->mode);
        return FDTD_E_ARG;
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        fdtd_set_error("fdtd_ctx_create: device %d not available (%d visible)", device, ndev);
        return FDTD_E_CUDA;
    }
    fdtd_ctx *c = new (std::nothrow) fdtd_ctx();
    if (!c) {
        fdtd_set_error("fdtd_ctx_create: out of host memory");
        return FDTD_E_NOMEM;
    }
    memset(c, 0, sizeof *c);
    c->p = *p;
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    int rc = fdtd_slab_range(p->maxk, rank, nranks, &c->k0, &c->k1);
    if (rc != FDTD_OK) {
        delete c;
        return rc;
    }
    if (c->k1 == c->k0) {
        fdtd_set_error("fdtd_ctx_create: rank %d of %d owns no plane of a %zu-plane cavity", rank, nranks, p->maxk);
        delete c;
        return FDTD_E_ARG;
    }
    Geo &g = c->g;
    g.I = (int)p->maxi;
    g.J = (int)p->maxj;
    g.K = (int)p->maxk;
    g.P = (int)(((p->maxi + 1) + 15) / 16 * 16);
    g.R = g.J + 1;
    g.PR = (long long)g.P * g.R;
    g.nk = (int)(c->k1 - c->k0);
    g.kbase = (int)c->k0;
    g.top = (rank == nranks - 1) ? 1 : 0;
    g.planes = g.nk + 2;
    /* every array has one spare plane below local plane 0 and one above plane nk + 1: the second halo
     * plane of the two-steps-per-sweep kernel (local planes -1 and nk + 2).  Everything else addresses
     * planes 0 .. nk + 1 and never sees them. */
    c->array_elems = (size_t)g.PR * (size_t)(g.planes + 2 + kRollGap);
    /* launch-shape limits: the per-cell kernels put one plane per blockIdx.z, and the fused kernels
     * index inside a plane with 32-bit offsets */
    if (g.planes > 65535 || g.PR >= (1LL << 31) - 4 * (long long)g.P) {
        fdtd_set_error("fdtd_ctx_create: slab of %d planes of %d x %d is outside the supported launch shapes "
                       "(at most 65533 planes per slab, plane below 2^31 elements)", g.nk, g.I, g.J);
        delete c;
        return FDTD_E_ARG;
    }
    c->ch = fdtd_factor_h(p);
    c->ce = fdtd_factor_e(p);
    /* default: two time steps per sweep over the TMA ring (fdtd_step2_tma.cuh) with the launch shape that
     * won the sweeps on a B200 (profiles/r02_sweep_step2_*): 28 x 13 stored sites per block of 32 x 8
     * threads, 32 planes per block, 3 stages in flight.  An odd step, and slabs thinner than two
     * planes, take the single-step TMA sweep (32 x 8 tile, 4 stages). */
    c->opt_kernel = 4;
    c->kernel_auto = true;
    c->opt_strip = 1;
    c->opt_kchunk = 32;
    c->opt_wx = 1;
    c->opt_wy = 8;
    c->opt_stages = 3;
    c->opt_prefetch = 3;
    c->opt_band = 1;
    c->opt_l2promo = 3;
    c->opt_cluster_x = c->opt_cluster_y = 1;
    c->opt_persistent = 0;
    c->opt_window = 4;
    c->opt_host_chunk = 0; /* automatic */
    c->opt_host_pipeline = 1;

    rc = fdtd_source_plan_make(p, &c->plan);
    if (rc != FDTD_OK) {
        delete c;
        return rc;
    }
    c->src_n = 0;
    c->src_here = false;
    c->src_staged = false;
    c->src_plane = 0;
    if (p->mode == 1) {
        /* the reference writes the patch without a bounds check (main.c:745-752); outside the
         * grid that is undefined behaviour there and an error here */
        if (c->plan.i0 < 0 || c->plan.j0 < 0 || c->plan.i1 > (long)p->maxi || c->plan.j1 > (long)p->maxj ||
            c->plan.i1 <= c->plan.i0 || c->plan.j1 <= c->plan.j0) {
            fdtd_set_error("fdtd_ctx_create: source patch i[%ld,%ld) j[%ld,%ld) does not fit the %zu x %zu grid",
                           c->plan.i0, c->plan.i1, c->plan.j0, c->plan.j1, p->maxi, p->maxj);
            delete c;
            return FDTD_E_ARG;
        }
        c->src_n = (int)(c->plan.i1 - c->plan.i0);
        c->src_here = (c->k0 == 0);
        c->src_staged = c->k0 <= 2; /* k0 = 1, 2: the plane is a halo plane that gets H recomputed here */
        c->src_plane = 1 - (int)c->k0;
    }

#define CREATE_TRY(expr)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fdtd_set_error("%s: %s", #expr, cudaGetErrorString(e_));                           \
            fdtd_ctx_destroy(c);                                                               \
            return e_ == cudaErrorMemoryAllocation ? FDTD_E_NOMEM : FDTD_E_CUDA;               \
        }                                                                                      \
    } while (0)

    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking));
    CREATE_TRY(alloc_state(c, &c->raw, &c->base)); /* initialize_fields(), main.c:294-364: all zero */
    c->f.ex = field_ptr(c, 0);
    c->f.ey = field_ptr(c, 1);
    c->f.ez = field_ptr(c, 2);
    c->f.hx = field_ptr(c, 3);
    c->f.hy = field_ptr(c, 4);
    c->f.hz = field_ptr(c, 5);
    int lo = 0, hi = 0;
    CREATE_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CREATE_TRY(cudaStreamCreateWithPriority(&c->s_comm, cudaStreamNonBlocking, hi));
    CREATE_TRY(cudaStreamCreateWithPriority(&c->s_dump, cudaStreamNonBlocking, lo));
    CREATE_TRY(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&c->ev_sent, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&c->ev_hhalo, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&c->ev_ehalo, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&c->ev_src, cudaEventDisableTiming));
    const size_t src_row = 2 * (size_t)std::max(c->src_n, 1);
    CREATE_TRY(cudaMalloc((void **)&c->src_dev, kSrcRing * src_row * sizeof(double)));
    CREATE_TRY(cudaMalloc((void **)&c->src_one_dev, src_row * sizeof(double)));
    CREATE_TRY(cudaHostAlloc((void **)&c->src_host, kSrcRing * src_row * sizeof(double), cudaHostAllocDefault));
    CREATE_TRY(cudaEventRecord(c->ev_src, c->s_main));
    CREATE_TRY(cudaStreamSynchronize(c->s_main));
#undef CREATE_TRY
    c->e_halo_valid = c->h_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = true; /* all zero: halos agree */
    *out = c;
    return FDTD_OK;
}


} /* namespace fdtdi */


/* ============================================ C ABI ========================================= */
extern "C" {

int fdtd_ctx_create(const fdtd_params *p, int device, fdtd_ctx **out)
{
    return create_impl(p, device, 0, 1, out);
}

int fdtd_ctx_create_slab(const fdtd_params *p, int device, int rank, int nranks, fdtd_ctx **out)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        fdtd_set_error("fdtd_ctx_create_slab: bad rank %d of %d", rank, nranks);
        return FDTD_E_ARG;
    }
    return create_impl(p, device, rank, nranks, out);
}

int fdtd_ctx_destroy(fdtd_ctx *c)
{
    if (!c)
        return FDTD_OK;
    cudaSetDevice(c->device);
    if (c->s_main && c->wired) wait_halos(c);
    if (c->s_main) cudaStreamSynchronize(c->s_main);
    if (c->s_comm) cudaStreamSynchronize(c->s_comm);
    if (c->s_dump) cudaStreamSynchronize(c->s_dump);
    pipe_destroy(c);
    halo_destroy(c);
    if (c->raw) cudaFree(c->raw);
    if (c->raw2) cudaFree(c->raw2);
    if (c->src_dev) cudaFree(c->src_dev);
    if (c->src_one_dev) cudaFree(c->src_one_dev);
    if (c->src_host) cudaFreeHost(c->src_host);
    if (c->agg_dev) cudaFree(c->agg_dev);
    if (c->progress_dev) cudaFree(c->progress_dev);
    if (c->roll_tmp) cudaFree(c->roll_tmp);
    cudaEvent_t evs[] = {c->ev_bnd, c->ev_sent, c->ev_hhalo, c->ev_ehalo, c->ev_src};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (c->s_main) cudaStreamDestroy(c->s_main);
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    if (c->s_dump) cudaStreamDestroy(c->s_dump);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    delete c;
    return FDTD_OK;
}

int fdtd_ctx_set_option(fdtd_ctx *c, const char *key, long value)
{
    if (!c) {
        fdtd_set_error("fdtd_ctx_set_option: context is NULL");
        return FDTD_E_ARG;
    }
    if (!key) {
        fdtd_set_error("fdtd_ctx_set_option: NULL key");
        return FDTD_E_ARG;
    }
    if (!strcmp(key, "kernel") && value >= 0 && value <= 4) {
        if (value == 4 && c->opt_kernel != 4) { /* its own launch defaults (profiles/r02_sweep_step2_*) */
            c->opt_wy = 8;
            c->opt_stages = 3;
            c->opt_kchunk = 32;
        }
        if (value == 3 && c->opt_kernel == 4) { /* and the single-step TMA sweep's */
            c->opt_strip = 1;
            c->opt_wx = 1;
            c->opt_wy = 8;
            c->opt_stages = 4;
            c->opt_kchunk = 32;
        }
        c->opt_kernel = value;
        c->kernel_auto = false;
    }
    else if (!strcmp(key, "stages") && value >= 2 && value <= kTmaMaxStages) c->opt_stages = value;
    else if (!strcmp(key, "strip") && value >= 1 && value <= 4) c->opt_strip = value;
    else if (!strcmp(key, "kchunk") && value >= 1 && value <= 1 << 20) c->opt_kchunk = value;
    else if (!strcmp(key, "warps_x") && value >= 1 && value <= 8) c->opt_wx = value;
    else if (!strcmp(key, "warps_y") && value >= 1 && value <= (c->opt_kernel == 4 ? 16 : 8)) c->opt_wy = value;
    else if (!strcmp(key, "prefetch") && value >= 0 && value <= 64) c->opt_prefetch = value;
    else if (!strcmp(key, "band") && value >= 1 && value <= 1024) c->opt_band = value;
    else if (!strcmp(key, "l2promo") && value >= 0 && value <= 3) c->opt_l2promo = value;
    else if (!strcmp(key, "persistent") && value >= 0 && value <= 1) c->opt_persistent = value;
    else if (!strcmp(key, "rolling") && value >= 0 && value <= 1) c->opt_rolling = value;
    else if (!strcmp(key, "window") && value >= 1 && value <= 64) c->opt_window = value;
    else if (!strcmp(key, "cluster_x") && value >= 1 && value <= 8) c->opt_cluster_x = value;
    else if (!strcmp(key, "cluster_y") && value >= 1 && value <= 8) c->opt_cluster_y = value;
    else if (!strcmp(key, "host_chunk") && value >= 0 && value <= 1 << 20) c->opt_host_chunk = value;
    else if (!strcmp(key, "host_pipeline") && value >= 0 && value <= 1) c->opt_host_pipeline = value;
    else {
        fdtd_set_error("fdtd_ctx_set_option: unknown key or bad value: %s = %ld", key, value);
        return FDTD_E_ARG;
    }
    return FDTD_OK;
}

int fdtd_ctx_get_option(fdtd_ctx *c, const char *key, long *value)
{
    FDTD_TRY(check_ctx(c, "fdtd_ctx_get_option"));
    if (!key || !value) {
        fdtd_set_error("fdtd_ctx_get_option: NULL argument");
        return FDTD_E_ARG;
    }
    if (!strcmp(key, "kernel")) *value = c->opt_kernel;
    else if (!strcmp(key, "strip")) *value = c->opt_strip;
    else if (!strcmp(key, "kchunk")) *value = c->opt_kchunk;
    else if (!strcmp(key, "warps_x")) *value = c->opt_wx;
    else if (!strcmp(key, "warps_y")) *value = c->opt_wy;
    else if (!strcmp(key, "prefetch")) *value = c->opt_prefetch;
    else if (!strcmp(key, "stages")) *value = c->opt_stages;
    else if (!strcmp(key, "band")) *value = c->opt_band;
    else if (!strcmp(key, "l2promo")) *value = c->opt_l2promo;
    else if (!strcmp(key, "persistent")) *value = c->opt_persistent;
    else if (!strcmp(key, "rolling")) *value = c->rolling ? 1 : c->opt_rolling;
    else if (!strcmp(key, "window")) *value = c->opt_window;
    else if (!strcmp(key, "cluster_x")) *value = c->opt_cluster_x;
    else if (!strcmp(key, "cluster_y")) *value = c->opt_cluster_y;
    else if (!strcmp(key, "host_chunk")) *value = c->opt_host_chunk;
    else if (!strcmp(key, "host_pipeline")) *value = c->opt_host_pipeline;
    else if (!strcmp(key, "k0")) *value = (long)c->k0;
    else if (!strcmp(key, "k1")) *value = (long)c->k1;
    else if (!strcmp(key, "launches")) *value = c->launches;
    else if (!strcmp(key, "fallback")) *value = c->fallback;
    else if (!strcmp(key, "transport")) *value = c->transport;
    else if (!strcmp(key, "fused_ok")) *value = c->nranks == 1 ? (c->base2 != nullptr) : (c->fused_ok ? 1 : 0);
    else {
        fdtd_set_error("fdtd_ctx_get_option: unknown key %s", key);
        return FDTD_E_ARG;
    }
    return FDTD_OK;
}

int fdtd_ctx_info(fdtd_ctx *c, size_t *hbm_bytes, size_t *pitch, size_t *rows, size_t *planes)
{
    FDTD_TRY(check_ctx(c, "fdtd_ctx_info"));
    if (hbm_bytes) *hbm_bytes = 6 * c->array_elems * sizeof(double) + c->agg_elems * sizeof(double);
    if (pitch) *pitch = (size_t)c->g.P;
    if (rows) *rows = (size_t)c->g.R;
    if (planes) *planes = (size_t)c->g.planes;
    return FDTD_OK;
}

static size_t slab_offset(const fdtd_ctx *c, int idx, bool whole_cavity)
{
    const DenseShape s = dense_shape(c->p, idx);
    return whole_cavity ? c->k0 * s.w * s.h : 0;
}

static int upload_impl(fdtd_ctx *c, const fdtd_fields *host, bool whole_cavity);
static int download_impl(fdtd_ctx *c, const fdtd_fields *host, bool whole_cavity);

int fdtd_upload(fdtd_ctx *c, const fdtd_fields *host) { return upload_impl(c, host, true); }
int fdtd_download(fdtd_ctx *c, const fdtd_fields *host) { return download_impl(c, host, true); }
int fdtd_upload_slab(fdtd_ctx *c, const fdtd_fields *host) { return upload_impl(c, host, false); }
int fdtd_download_slab(fdtd_ctx *c, const fdtd_fields *host) { return download_impl(c, host, false); }

static int upload_impl(fdtd_ctx *c, const fdtd_fields *host, bool whole_cavity)
{
    FDTD_TRY(check_ctx(c, "fdtd_upload"));
    if (!host || !host->Ex || !host->Ey || !host->Ez || !host->Hx || !host->Hy || !host->Hz) {
        fdtd_set_error("fdtd_upload: NULL field pointer");
        return FDTD_E_ARG;
    }
    FDTD_TRY(use_device(c));
    FDTD_TRY(wait_halos(c)); /* a neighbour may still be reading the planes this overwrites */
    double *h[6] = {host->Ex, host->Ey, host->Ez, host->Hx, host->Hy, host->Hz};
    for (int a = 0; a < 6; ++a)
        FDTD_TRY(copy_field(c, a, h[a] + slab_offset(c, a, whole_cavity), true));
    c->e_halo_valid = c->h_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = (c->nranks == 1);
    return FDTD_OK;
}

static int download_impl(fdtd_ctx *c, const fdtd_fields *host, bool whole_cavity)
{
    FDTD_TRY(check_ctx(c, "fdtd_download"));
    if (!host || !host->Ex || !host->Ey || !host->Ez || !host->Hx || !host->Hy || !host->Hz) {
        fdtd_set_error("fdtd_download: NULL field pointer");
        return FDTD_E_ARG;
    }
    FDTD_TRY(use_device(c));
    double *h[6] = {host->Ex, host->Ey, host->Ez, host->Hx, host->Hy, host->Hz};
    for (int a = 0; a < 6; ++a)
        FDTD_TRY(copy_field(c, a, h[a] + slab_offset(c, a, whole_cavity), false));
    CUDA_TRY(cudaStreamSynchronize(c->s_main));
    return FDTD_OK;
}

int fdtd_set_initial_conditions(fdtd_ctx *c)
{
    FDTD_TRY(check_ctx(c, "fdtd_set_initial_conditions"));
    FDTD_TRY(use_device(c));
    /* glibc sin on the host, like the reference (main.c:422-423), for the planes this slab owns,
     * then one upload of Ey */
    const DenseShape s = dense_shape(c->p, 1);
    const size_t nplanes = (size_t)c->g.nk + (c->g.top ? 1 : 0);
    double *ey = (double *)malloc(s.w * s.h * nplanes * sizeof(double));
    if (!ey) {
        fdtd_set_error("fdtd_set_initial_conditions: out of host memory");
        return FDTD_E_NOMEM;
    }
    int rc = fdtd_initial_conditions_planes(&c->p, c->k0, nplanes, ey);
    if (rc == FDTD_OK)
        rc = wait_halos(c);
    if (rc == FDTD_OK)
        rc = copy_field(c, 1, ey, true);
    if (rc == FDTD_OK && cudaStreamSynchronize(c->s_main) != cudaSuccess) {
        fdtd_set_error("fdtd_set_initial_conditions: upload failed");
        rc = FDTD_E_CUDA;
    }
    free(ey);
    c->e_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = (c->nranks == 1);
    return rc;
}

int fdtd_set_source(fdtd_ctx *c, double t)
{
    FDTD_TRY(check_solo(c, "fdtd_set_source"));
    FDTD_TRY(use_device(c));
    if (c->p.mode != 1 || c->src_n == 0) {
        fdtd_set_error("fdtd_set_source: context has no source (validation mode)");
        return FDTD_E_STATE;
    }
    if (c->nranks > 1)
        c->h_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = false; /* plane k = 0 may be a plane that travels (same on all ranks) */
    if (!c->src_here)
        return FDTD_OK; /* the patch lives on the slab that holds k = 0 */
    FDTD_TRY(wait_halos(c));
    std::vector<double> row(2 * (size_t)c->src_n);
    FDTD_TRY(fdtd_source_values(&c->p, &c->plan, t, row.data(), row.data() + c->src_n));
    CUDA_TRY(cudaMemcpyAsync(c->src_one_dev, row.data(), row.size() * sizeof(double),
                             cudaMemcpyHostToDevice, c->s_main));
    CUDA_TRY(cudaStreamSynchronize(c->s_main)); /* `row` is pageable and about to go away */
    launch_set_source(c, c->src_one_dev, c->s_main);
    CUDA_TRY(cudaGetLastError());
    return FDTD_OK;
}

int fdtd_update_H_field(fdtd_ctx *c)
{
    FDTD_TRY(check_solo(c, "fdtd_update_H_field"));
    FDTD_TRY(use_device(c));
    FDTD_TRY(refresh_halos_many(&c, 1, false));
    FDTD_TRY(wait_halos(c));
    launch_h(c, no_src(), 1, c->g.nk + c->g.top + 1, c->s_main);
    CUDA_TRY(cudaGetLastError());
    c->h_halo_valid = c->wide_halo_valid = (c->nranks == 1);
    return FDTD_OK;
}

int fdtd_update_E_field(fdtd_ctx *c)
{
    FDTD_TRY(check_solo(c, "fdtd_update_E_field"));
    FDTD_TRY(use_device(c));
    FDTD_TRY(refresh_halos_many(&c, 1, false));
    FDTD_TRY(wait_halos(c));
    launch_e(c, no_src(), 1, c->g.nk + 1, c->s_main);
    CUDA_TRY(cudaGetLastError());
    c->e_halo_valid = c->low_e_halo_valid = c->wide_halo_valid = (c->nranks == 1);
    return FDTD_OK;
}

int fdtd_run(fdtd_ctx *c, size_t steps, double *time_counter)
{
    FDTD_TRY(check_solo(c, "fdtd_run"));
    if (!time_counter) {
        fdtd_set_error("fdtd_run: time_counter is NULL");
        return FDTD_E_ARG;
    }
    return run_impl(c, steps, time_counter, nullptr, nullptr, nullptr);
}

int fdtd_run_timed(fdtd_ctx *c, size_t steps, double *time_counter, float *total_ms, float *h_ms, float *e_ms)
{
    FDTD_TRY(check_solo(c, "fdtd_run_timed"));
    if (!time_counter || !total_ms) {
        fdtd_set_error("fdtd_run_timed: NULL argument");
        return FDTD_E_ARG;
    }
    return run_impl(c, steps, time_counter, total_ms, h_ms, e_ms);
}

int fdtd_sync(fdtd_ctx *c)
{
    FDTD_TRY(check_ctx(c, "fdtd_sync"));
    FDTD_TRY(use_device(c));
    FDTD_TRY(wait_halos(c)); /* also what the neighbours still push into this slab's halo planes */
    CUDA_TRY(cudaStreamSynchronize(c->s_main));
    CUDA_TRY(cudaStreamSynchronize(c->s_comm));
    CUDA_TRY(cudaStreamSynchronize(c->s_dump));
    return FDTD_OK;
}
/* NUMA placement of pinned host buffers.  On a two-socket 8-GPU box every rank streams ~50 GB each way
 * through its buffers; if they all sit on one socket, the GPUs of the other one pull them across the
 * socket interconnect.  node >= 0: the pages come from that node; node == -2: interleaved over all nodes
 * (for boxes that do not tell which node a GPU hangs on); anything else: the default policy. */
static long set_mempolicy_raw(int mode, const unsigned long *mask, unsigned long maxnode)
{
#ifdef SYS_set_mempolicy
    return syscall(SYS_set_mempolicy, mode, mask, maxnode);
#else
    (void)mode; (void)mask; (void)maxnode;
    return -1;
#endif
}

static int numa_node_count()
{
    int n = 0;
    for (int k = 0; k < 64; ++k) {
        char path[64];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d", k);
        if (access(path, F_OK) == 0)
            n = k + 1;
    }
    return n;
}

static int numa_node_of_device(int device)
{
    char bus[32] = {0}, path[128];
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *q = bus; *q; ++q)
        *q = (char)tolower((unsigned char)*q);
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f)
        return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1)
        node = -1;
    fclose(f);
    return node;
}

static int host_alloc_policy(size_t bytes, void **out, int node)
{
    if (!out) {
        fdtd_set_error("fdtd_host_alloc: NULL argument");
        return FDTD_E_ARG;
    }
    const int nodes = numa_node_count();
    bool policy_set = false;
    if (nodes > 1 && nodes <= 64) {
        unsigned long mask = 0;
        int mode = 0;
        if (node >= 0 && node < nodes) {
            mask = 1ul << node;
            mode = 1; /* MPOL_PREFERRED */
        } else if (node == -2) {
            mask = nodes == 64 ? ~0ul : (1ul << nodes) - 1;
            mode = 3; /* MPOL_INTERLEAVE */
        }
        if (mask)
            policy_set = set_mempolicy_raw(mode, &mask, 65) == 0;
    }
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (policy_set)
        set_mempolicy_raw(0 /* MPOL_DEFAULT */, nullptr, 0);
    if (e != cudaSuccess) {
        fdtd_set_error("cudaHostAlloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        *out = nullptr;
        return e == cudaErrorMemoryAllocation ? FDTD_E_NOMEM : FDTD_E_CUDA;
    }
    return FDTD_OK;
}

int fdtd_host_alloc(size_t bytes, void **out)
{
    return host_alloc_policy(bytes, out, -1);
}

int fdtd_host_alloc_near(int device, size_t bytes, void **out)
{
    /* FDTD_B200_HOST_NUMA: "near" (default: the node the GPU hangs on, if the box tells), "interleave", "off" */
    const char *env = getenv("FDTD_B200_HOST_NUMA");
    int node = -1;
    if (env && !strcmp(env, "interleave"))
        node = -2;
    else if (!(env && !strcmp(env, "off")))
        node = numa_node_of_device(device);
    return host_alloc_policy(bytes, out, node);
}

int fdtd_host_numa_info(int device, int *nodes, int *node_of_device)
{
    if (nodes) *nodes = numa_node_count();
    if (node_of_device) *node_of_device = numa_node_of_device(device);
    return FDTD_OK;
}

int fdtd_host_free(void *ptr)
{
    if (ptr)
        CUDA_TRY(cudaFreeHost(ptr));
    return FDTD_OK;
}

} /* extern "C" */
