"""numpy restatement of the TWO-STEPS-PER-SWEEP schedule (csrc/fdtd_step2_tma.cuh), for the CPU test-suite.

It follows the kernel block by block: an extended tile of 32 x 2*WY sites is loaded plane by plane
(zero outside the arrays, like TMA), every site evaluates the four stages
    A  H1(k)   B  E1(k)   C  H2(k-1)   D  E2(k-1)
with its neighbours' values taken from the tile, and only the inner 28 x (2*WY - 3) sites are stored.
Whatever a site at the rim of the extended tile would read from beyond the tile is NaN here (the
kernel reads leftovers there): if the validity bookkeeping were wrong, NaN would reach a stored value.
PEC walls, the source at both time levels, the two-plane run-in of a chunk and the "cell plane by
GLOBAL index" rule for slabs are the kernel's.  Elementwise IEEE double in the reference's operand
order, so the result must equal the oracle's bit for bit.

Arrays are the reference's dense ones (lower-case keys, shape (planes, rows, columns)); `klo`/`khi`
restrict the sweep to a slab's cell planes of a cavity whose arrays are passed whole.
"""
import numpy as np

TILE_X, EXT_X = 28, 32


def yee(f, c, a, b, d, e):
    return f + c * ((a - b) - (d - e))


def _box(arr, k, x0, y0, w, h):
    """h x w window of plane k starting at (x0, y0); zero wherever the array has no element"""
    out = np.zeros((h, w))
    if k < 0 or k >= arr.shape[0]:
        return out
    ys, xs = max(y0, 0), max(x0, 0)
    ye, xe = min(y0 + h, arr.shape[1]), min(x0 + w, arr.shape[2])
    if ye > ys and xe > xs:
        out[ys - y0:ye - y0, xs - x0:xe - x0] = arr[k, ys:ye, xs:xe]
    return out


def _xm(v):   # value of the site at i - 1 (lane 0 has no such lane)
    return np.concatenate([np.full((v.shape[0], 1), np.nan), v[:, :-1]], axis=1)


def _xp(v):
    return np.concatenate([v[:, 1:], np.full((v.shape[0], 1), np.nan)], axis=1)


def _ym(v):   # value of the site at j - 1 (the lowest row of the tile has none)
    return np.concatenate([np.full((1, v.shape[1]), np.nan), v[:-1, :]], axis=0)


def _yp(v):
    return np.concatenate([v[1:, :], np.full((1, v.shape[1]), np.nan)], axis=0)


def step2_sweep(a, b, dims, ch, ce, src1=None, src2=None, wy=8, kchunk=4, klo=0, khi=None):
    """Two time steps: reads `a`, writes `b`; cell planes [klo, khi) plus, when khi == nz, node plane nz.
    src = (i0, i1, j0, j1, ez_vals, hx_vals) of the first / second step, or None."""
    nx, ny, nz = dims
    khi = nz if khi is None else khi
    bye = 2 * wy
    end = khi + (1 if khi == nz else 0)
    for c0 in range(klo, end, kchunk):
        c1 = min(c0 + kchunk, end)
        for Y0 in range(0, ny + 1, bye - 3):
            for X0 in range(0, nx + 1, TILE_X):
                _tile(a, b, dims, ch, ce, src1, src2, bye, X0, Y0, c0, c1)


def _tile(a, b, dims, ch, ce, src1, src2, bye, X0, Y0, c0, c1):
    nx, ny, nz = dims
    x0, y0 = X0 - 2, Y0 - 2
    i = x0 + np.arange(EXT_X)[None, :]
    j = y0 + np.arange(bye)[:, None]
    xn, xc, xi = (i >= 0) & (i <= nx), (i >= 0) & (i < nx), (i >= 1) & (i < nx)
    jn, jc, ji = (j >= 0) & (j <= ny), (j >= 0) & (j < ny), (j >= 1) & (j < ny)
    up_x, up_y, up_z = xc & ji, xi & jc, xi & ji
    out = (np.arange(EXT_X)[None, :] >= 2) & (np.arange(EXT_X)[None, :] < 30) & \
          (np.arange(bye)[:, None] >= 2) & (np.arange(bye)[:, None] < bye - 1)
    st = {"hx": out & xn & jc, "ey": out & xn & jc, "hy": out & xc & jn, "ex": out & xc & jn,
          "hz": out & xc & jc, "ez": out & xn & jn}

    def patch(src, di=0, dj=0):
        i0, i1, j0, j1 = src[:4]
        return (i + di >= i0) & (i + di < i1) & (j + dj >= j0) & (j + dj < j1)

    def amp(src, which, di=0):   # amplitude by column, wherever the column is inside the patch
        vals = src[4] if which == "ez" else src[5]
        col = np.clip(i + di - src[0], 0, len(vals) - 1)
        return np.broadcast_to(vals[col], (bye, EXT_X))

    def win(name, k, dx=0, dy=0):
        return _box(a[name], k, x0 + dx, y0 + dy, EXT_X, bye)

    kstart = max(c0 - 2, 0)
    ex0, ey0 = win("ex", kstart), win("ey", kstart)
    zero = np.zeros((bye, EXT_X))
    hx1p, hy1p, hz1p = zero, zero, zero
    ex1p, ey1p, ez1p = zero, zero, zero
    hx2pp, hy2pp = zero, zero
    for k in range(kstart, c1 + 1):
        cell1, kin1 = k < nz, 1 <= k < nz
        cell2, kin2 = k - 1 < nz, 1 <= k - 1 < nz
        # ---- A: H1(k)
        ez0, ez0i, ez0j = win("ez", k), win("ez", k, dx=1), win("ez", k, dy=1)
        ex0j, ey0i = win("ex", k, dy=1), win("ey", k, dx=1)
        ex0n, ey0n = win("ex", k + 1), win("ey", k + 1)
        if src1 is not None and k == 0:
            ez0 = np.where(patch(src1), amp(src1, "ez"), ez0)
            ex0 = np.where(patch(src1), 0.0, ex0)
            ez0i = np.where(patch(src1, di=1), amp(src1, "ez", di=1), ez0i)
            ez0j = np.where(patch(src1, dj=1), amp(src1, "ez"), ez0j)
            ex0j = np.where(patch(src1, dj=1), 0.0, ex0j)
        hx1 = yee(win("hx", k), ch, ey0n, ey0, ez0j, ez0)
        hy1 = yee(win("hy", k), ch, ez0i, ez0, ex0n, ex0)
        hz1 = yee(win("hz", k), ch, ex0j, ex0, ey0i, ey0)
        if src1 is not None and k == 0:
            hx1 = np.where(patch(src1), amp(src1, "hx"), hx1)
            hz1 = np.where(patch(src1), 0.0, hz1)
        # ---- B: E1(k)
        ex1, ey1, ez1 = ex0, ey0, ez0
        if kin1:
            ex1 = np.where(up_x, yee(ex0, ce, hz1, _ym(hz1), hy1, hy1p), ex0)
            ey1 = np.where(up_y, yee(ey0, ce, hx1, hx1p, hz1, _xm(hz1)), ey0)
        if cell1:
            ez1 = np.where(up_z, yee(ez0, ce, hy1, _xm(hy1), hx1, _ym(hx1)), ez0)
        if src2 is not None and k == 0:
            ez1 = np.where(patch(src2), amp(src2, "ez"), ez1)
            ex1 = np.where(patch(src2), 0.0, ex1)
        # ---- C: H2(k-1)
        hx2 = yee(hx1p, ch, ey1, ey1p, _yp(ez1p), ez1p)
        hy2 = yee(hy1p, ch, _xp(ez1p), ez1p, ex1, ex1p)
        hz2 = yee(hz1p, ch, _yp(ex1p), ex1p, _xp(ey1p), ey1p)
        if src2 is not None and k - 1 == 0:
            hx2 = np.where(patch(src2), amp(src2, "hx"), hx2)
            hz2 = np.where(patch(src2), 0.0, hz2)
        # ---- D: E2(k-1), stores
        vex, vey, vez = ex1p, ey1p, ez1p
        if kin2:
            vex = np.where(up_x, yee(ex1p, ce, hz2, _ym(hz2), hy2, hy2pp), ex1p)
            vey = np.where(up_y, yee(ey1p, ce, hx2, hx2pp, hz2, _xm(hz2)), ey1p)
        if cell2:
            vez = np.where(up_z, yee(ez1p, ce, hy2, _xm(hy2), hx2, _ym(hx2)), ez1p)
        if k - 1 >= c0:
            stores = [("hz", hz2), ("ex", vex), ("ey", vey)]
            if cell2:
                stores += [("hx", hx2), ("hy", hy2), ("ez", vez)]
            for name, val in stores:
                rows, cols = np.nonzero(st[name])
                b[name][k - 1, rows + y0, cols + x0] = val[rows, cols]
        hx2pp, hy2pp = hx2, hy2
        hx1p, hy1p, hz1p = hx1, hy1, hz1
        ex1p, ey1p, ez1p = ex1, ey1, ez1
        ex0, ey0 = ex0n, ey0n
