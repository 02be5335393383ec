"""The rolling window of the in-place two-step sweep (csrc/fdtd_ctx.cu, launch_step2_t / roll_*), as a
model on plane LABELS: every array is a ring of nk + 4 + gap plane slots; a sweep reads local plane k of
time level n from slot (k + 1 + rot) mod Z and writes plane k of level n + 1 to slot (k + 1 + rot - D) mod
Z, chunk after chunk from the bottom, D = planes per chunk + 4.  Within a chunk the blocks run in any
order and at any relative pace (here: the adversarial extremes), so a write may only hit a slot whose
old content no block of the same or a later chunk still needs.  The model checks that every read finds the
label it expects, over several sweeps (the ring wraps), with halo planes rewritten between sweeps, and
that rotating the ring back restores the canonical layout."""
import itertools

import pytest

GAP = 40


def sweep(ring, rot, nk, top, kc, level, order):
    """one sweep of a slab with nk cell planes: planes 1 .. nk + top are written; reads reach two planes
    below a chunk and one above.  `order`: how the blocks of a chunk interleave their plane iterations."""
    Z = len(ring)
    D = kc + 4
    slot = lambda k, r: (k + 1 + r) % Z
    end = nk + top + 1
    for a in range(1, end, kc):
        b = min(a + kc, end)
        kstart = max(a - 2, -1)
        # two blocks of the same chunk: a leader and a laggard.  order = "lockstep", "lead" (the leader finishes
        # the whole chunk before the laggard starts) -- the laggard must still find every plane it reads.
        # (on the top slab the sweep also touches plane nk + 3, which does not exist: the kernel never uses it)
        reads = [(k, slot(k, rot)) for k in range(kstart, min(b + 2, nk + 3))]
        writes = [(k, slot(k, rot - D)) for k in range(a, b)]
        if order == "lead":
            snapshot = list(ring)
            for k, s in reads:
                assert ring[s] == (k, level), f"leader reads plane {k}: slot holds {ring[s]}"
            for k, s in writes:
                ring[s] = (k, level + 1)
            for k, s in reads:                      # the laggard comes after all of the leader's writes
                assert ring[s] == (k, level) or ring[s] == snapshot[s] == (k, level), \
                    f"chunk [{a},{b}): plane {k} of level {level} was overwritten by a block of the same chunk"
        else:
            for k, s in reads:
                assert ring[s] == (k, level), f"plane {k}: slot {s} holds {ring[s]}"
            for k, s in writes:
                ring[s] = (k, level + 1)
    return ((rot - D) % Z + Z) % Z


@pytest.mark.parametrize("nk,top,kc,order", [c for c in itertools.product((2, 5, 36, 41, 128, 300), (0, 1), (2, 3, 32, 36),
                                                                          ("lockstep", "lead"))])
def test_in_place_sweeps_never_lose_a_plane(nk, top, kc, order):
    Z = nk + 4 + GAP
    ring = [None] * Z
    for k in range(-1, nk + 3):                      # canonical layout: local plane k in slot k + 1
        ring[k + 1] = (k, 0)
    rot = 0
    for level in range(7):                           # enough sweeps for the ring to wrap several times
        rot_new = sweep(ring, rot, nk, top, kc, level, order)
        # what the sweep does not write is rewritten at the new rotation before the next one reads it:
        # the halo planes (by the exchange) and, on a slab that is not the top one, planes nk+1, nk+2
        written = set(range(1, nk + top + 1))
        for k in range(-1, nk + 3):
            if k not in written:
                ring[(k + 1 + rot_new) % Z] = (k, level + 1)
        rot = rot_new
        for k in range(-1, nk + 3):
            assert ring[(k + 1 + rot) % Z] == (k, level + 1), (k, level)
    # rotate back (every slot moves once, cycle by cycle, as roll_canonicalise does)
    r, g = rot, 0
    a, b = Z, r
    while b:
        a, b = b, a % b
    g = a
    if r:
        for c0 in range(g):
            tmp, j = ring[c0], c0
            while True:
                src = (j + r) % Z
                if src == c0:
                    break
                ring[j] = ring[src]
                j = src
            ring[j] = tmp
    for k in range(-1, nk + 3):
        assert ring[k + 1] == (k, 7)
