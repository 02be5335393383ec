"""numpy restatement of the FUSED single-sweep step (csrc/fdtd_fused.cuh / fdtd_fused_tma.cuh), for the
CPU test-suite: same plane-by-plane schedule as the kernels -- H of plane k is finished and at once
consumed by E of plane k, H_new of the plane below is carried along (or recomputed at the start of
a chunk), reads come from state `a` only, writes go to state `b` only, un-updated elements are
copied, source and PEC are applied by substitution exactly as SURVEY.md B.4 prescribes.

Everything is elementwise IEEE double in the reference's operand order
(F = F + c * ((a - b) - (d - e)), no FMA in numpy), so the result must equal the oracle's bit for bit.
Arrays are the reference's dense ones, lower-case keys, shapes (planes, rows, columns).

`klo` / `khi` restrict the sweep to a slab's cell planes [klo, khi) of a cavity whose arrays are
passed whole (planes outside [klo-1, khi] are never touched), which is how the slab tests use it.
"""
import numpy as np


def yee(f, c, a, b, d, e):
    return f + c * ((a - b) - (d - e))


def h_plane(a, k, nz, ch, src):
    """H at time n+1/2 on plane k from state a (time n): (hx, hy, hz); hx, hy are None for k == nz."""
    ex_k, ey_k = a["ex"][k], a["ey"][k]
    srck = src is not None and k == 0
    if srck:   # first set_source: Ex = 0, Ez = amplitude on the patch (main.c:748-749)
        i0, i1, j0, j1, ez_vals, _ = src
        ex_k = ex_k.copy()
        ex_k[j0:j1, i0:i1] = 0.0
    hz = yee(a["hz"][k], ch, ex_k[1:, :], ex_k[:-1, :], ey_k[:, 1:], ey_k[:, :-1])      # main.c:460
    hx = hy = None
    if k < nz:
        ez_k = a["ez"][k]
        if srck:
            ez_k = ez_k.copy()
            ez_k[j0:j1, i0:i1] = ez_vals[None, :]
        hx = yee(a["hx"][k], ch, a["ey"][k + 1], ey_k, ez_k[1:, :], ez_k[:-1, :])       # main.c:448
        hy = yee(a["hy"][k], ch, ez_k[:, 1:], ez_k[:, :-1], a["ex"][k + 1], ex_k)       # main.c:454
    if srck:   # second set_source overwrites Hx, Hz on the patch (main.c:750-751)
        _, _, _, _, _, hx_vals = src
        hz[j0:j1, i0:i1] = 0.0
        if hx is not None:
            hx[j0:j1, i0:i1] = hx_vals[None, :]
    return hx, hy, hz


def fused_step(a, b, dims, ch, ce, src=None, kchunk=8, klo=0, khi=None):
    """One time step: reads `a`, writes `b` (dicts of the six dense arrays); cell planes [klo, khi)
    plus, when khi == nz, the top node plane nz."""
    nx, ny, nz = dims
    khi = nz if khi is None else khi
    top = khi == nz
    for c0 in range(klo, khi + (1 if top else 0), kchunk):
        c1 = min(c0 + kchunk, khi + (1 if top else 0))
        hx_m = hy_m = None
        if c0 >= 1:                       # "H only" prologue: H_new of the plane below the chunk
            hx_m, hy_m, _ = h_plane(a, c0 - 1, nz, ch, src)
        for k in range(c0, c1):
            hx, hy, hz = h_plane(a, k, nz, ch, src)
            srck = src is not None and k == 0
            b["hz"][k] = hz
            ex_new, ey_new = a["ex"][k].copy(), a["ey"][k].copy()
            if srck:
                i0, i1, j0, j1, ez_vals, _ = src
                ex_new[j0:j1, i0:i1] = 0.0                                              # main.c:749
            if k < nz:
                b["hx"][k], b["hy"][k] = hx, hy
                if k >= 1:                # PEC: Ex, Ey untouched on k = 0 and k = nz
                    ex_new[1:ny, :] = yee(ex_new[1:ny, :], ce, hz[1:ny, :], hz[0:ny - 1, :],
                                          hy[1:ny, :], hy_m[1:ny, :])                   # main.c:486
                    ey_new[:, 1:nx] = yee(ey_new[:, 1:nx], ce, hx[:, 1:nx], hx_m[:, 1:nx],
                                          hz[:, 1:nx], hz[:, 0:nx - 1])                 # main.c:492
                ez_new = a["ez"][k].copy()
                if srck:
                    ez_new[j0:j1, i0:i1] = ez_vals[None, :]                             # main.c:748
                ez_new[1:ny, 1:nx] = yee(ez_new[1:ny, 1:nx], ce, hy[1:ny, 1:nx], hy[1:ny, 0:nx - 1],
                                         hx[1:ny, 1:nx], hx[0:ny - 1, 1:nx])            # main.c:498
                b["ez"][k] = ez_new
            b["ex"][k], b["ey"][k] = ex_new, ey_new
            hx_m, hy_m = hx, hy
