"""The schedule of fdtd_run_hosted (csrc/fdtd_hosted.cu) checked on the CPU, without any arithmetic:
the order in which a slab queues chunk uploads, sweeps, halo pushes and downloads is restated here
(same wave / sweep / chunk formulas, same sequence numbers) and executed by a little interpreter
that runs every rank's in-order stream until it blocks.  Checked: nobody deadlocks; a sweep only
starts when the chunks it reads (its own and both neighbours, one sweep older) are there, when the
buffer it overwrites has no reader left, and -- next to an interface -- when the neighbour's halo of
the right sweep has arrived and has not been overwritten by a newer one; pushes never overwrite a
halo that is still unread; every chunk is downloaded after its last sweep."""
import itertools

import pytest


def schedule(rank, nranks, M, S):
    """the op list of one rank, in stream order (wavefront_slab)"""
    up_first = rank % 2 == 0
    has_lo, has_hi = rank > 0, rank + 1 < nranks
    chunk_of = (lambda q: q) if up_first else (lambda q: M - 1 - q)
    ops = []
    pushed0 = {"top": False, "bottom": False}

    def touches(q):
        k = chunk_of(q)
        return has_hi and k == M - 1, has_lo and k == 0

    for u in range(1, M - 1 + S + 1):
        landed = min(u, M - 1)
        if u <= M:
            ops.append(("wait_upload", landed))
        for q in range(landed + 1):
            top, bottom = touches(q)
            if top and not pushed0["top"]:
                ops.append(("push", "top", 0))
                pushed0["top"] = True
            if bottom and not pushed0["bottom"]:
                ops.append(("push", "bottom", 0))
                pushed0["bottom"] = True
        for s in range(max(1, u - M + 1), min(S, u) + 1):
            q = u - s
            top, bottom = touches(q)
            if top:
                ops.append(("wait_halo", "top", s - 1))
            if bottom:
                ops.append(("wait_halo", "bottom", s - 1))
            ops.append(("sweep", s, chunk_of(q)))
            if top:
                ops.append(("push", "top", s))
            if bottom:
                ops.append(("push", "bottom", s))
            if s == S:
                ops.append(("download", chunk_of(q)))
    return ops, [chunk_of(q) for q in range(M)]


def run(nranks, M, S):
    ops, upload_order, pc = {}, {}, {}
    for r in range(nranks):
        ops[r], upload_order[r] = schedule(r, nranks, M, S)
        pc[r] = 0
    done = {r: {} for r in range(nranks)}            # (chunk) -> last sweep completed; 0 = uploaded
    halo = {r: {"top": {}, "bottom": {}} for r in range(nranks)}   # side -> buffer parity -> sweep number of the content
    consumed = {r: {"top": -1, "bottom": -1} for r in range(nranks)}   # highest halo number this rank has used, per side
    downloaded = {r: set() for r in range(nranks)}
    progress = True
    while progress:
        progress = False
        for r in range(nranks):
            while pc[r] < len(ops[r]):
                op = ops[r][pc[r]]
                if op[0] == "wait_upload":
                    for k in upload_order[r][:op[1] + 1]:          # the copy stream is in order
                        done[r].setdefault(k, 0)
                elif op[0] == "wait_halo":
                    side, p = op[1], op[2]
                    if halo[r][side].get(p & 1, -1) < p:
                        break                                       # blocked: the neighbour has not pushed it yet
                    assert halo[r][side][p & 1] == p, f"rank {r}: halo {side} of sweep {p} was overwritten"
                elif op[0] == "sweep":
                    s, k = op[1], op[2]
                    for kk in (k - 1, k, k + 1):
                        if 0 <= kk < M:
                            assert done[r].get(kk, -1) >= s - 1, f"rank {r} sweep {s} chunk {k}: chunk {kk} not at {s - 1}"
                            assert done[r].get(kk, -1) <= s if kk != k else done[r][kk] == s - 1, \
                                f"rank {r} sweep {s} chunk {k}: chunk {kk} already overwritten (at {done[r][kk]})"
                    # the buffer this sweep writes (that of sweep s - 2) must have no reader left among my chunks
                    for kk in (k - 1, k + 1):
                        if 0 <= kk < M and s >= 2:
                            assert done[r].get(kk, -1) >= s - 1, f"rank {r}: chunk {kk} still reads what sweep {s} of {k} overwrites"
                    done[r][k] = s
                    for side, edge in (("top", M - 1), ("bottom", 0)):
                        if k == edge and (r + 1 < nranks if side == "top" else r > 0):
                            consumed[r][side] = s - 1
                elif op[0] == "push":
                    side, p = op[1], op[2]
                    nb, nb_side = (r + 1, "bottom") if side == "top" else (r - 1, "top")
                    # the slot (parity p) still holds push p - 2: the neighbour must have used it (ack >= p - 2)
                    if p >= 2 and consumed[nb][nb_side] < p - 2:
                        break                                       # blocked on the acknowledgement
                    edge = M - 1 if side == "top" else 0
                    assert done[r].get(edge, -1) == p, f"rank {r}: pushes sweep {p} of a chunk that is at {done[r].get(edge)}"
                    halo[nb][nb_side][p & 1] = p
                elif op[0] == "download":
                    assert done[r][op[1]] == S
                    downloaded[r].add(op[1])
                pc[r] += 1
                progress = True
    for r in range(nranks):
        assert pc[r] == len(ops[r]), f"deadlock: rank {r} stuck at {ops[r][pc[r]]} (M={M}, S={S}, n={nranks})"
        assert downloaded[r] == set(range(M))
        assert all(done[r][k] == S for k in range(M))


@pytest.mark.parametrize("nranks,M,S", [c for c in itertools.product((1, 2, 3, 4, 5, 8), (1, 2, 3, 7, 16), (1, 2, 3, 10, 32))])
def test_meshed_wavefront_never_blocks_and_respects_every_dependency(nranks, M, S):
    run(nranks, M, S)
