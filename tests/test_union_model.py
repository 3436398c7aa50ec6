"""CPU model of the invariants kmd_union.cuh relies on (the kernels themselves are held to the oracle on the GPU,
tests/test_gpu_merge.py): a numpy restatement of one level of the union tree — merge-path splits of the tiles and of
the threads, ties put run A first, a B key is dropped when the A key right before it (the halo at a tile's left edge)
equals it — of the ticket order of the tiles, and of the row search bounds of the scatter."""
import numpy as np

TILE, ITEMS = 16, 4
THREADS = TILE // ITEMS


def merge_path(diag, la, lb, a, b):
    lo = diag - lb if diag > lb else 0
    hi = diag if diag < la else la
    while lo < hi:
        mid = (lo + hi) >> 1
        if a(mid) <= b(diag - 1 - mid):
            lo = mid + 1
        else:
            hi = mid
    return lo


def union_level(A, B):
    """level_partition_kernel + union_level_kernel for one pair, tile by tile and thread by thread"""
    la, lb = len(A), len(B)
    total = la + lb
    out = []
    for lt in range((total + TILE - 1) // TILE + 2):     # the host's tile count is an upper bound: extra tiles are empty
        d0 = lt * TILE
        if d0 >= total:
            continue
        d1 = min(d0 + TILE, total)
        ai0 = merge_path(d0, la, lb, lambda i: A[i], lambda i: B[i])
        ai1 = la if d1 == total else merge_path(d1, la, lb, lambda i: A[i], lambda i: B[i])
        na, nb = ai1 - ai0, (d1 - ai1) - (d0 - ai0)
        halo = ai0 > 0
        s = [A[ai0 - 1] if halo else None] + [A[ai0 + i] for i in range(na)] + [B[d0 - ai0 + i] for i in range(nb)]
        nt = na + nb
        for tid in range(THREADS):
            t0 = min(tid * ITEMS, nt)
            t1 = min(t0 + ITEMS, nt)
            ia = merge_path(t0, na, nb, lambda i: s[1 + i], lambda i: s[1 + na + i])
            ib = t0 - ia
            pa_ok = ia > 0 or halo
            pa = s[ia] if pa_ok else None
            for _ in range(t1 - t0):
                take_a = ib >= nb or (ia < na and s[1 + ia] <= s[1 + na + ib])
                if take_a:
                    out.append(s[1 + ia]); pa, pa_ok = s[1 + ia], True; ia += 1
                else:
                    kb = s[1 + na + ib]
                    if not (pa_ok and pa == kb):
                        out.append(kb)
                    ib += 1
    return np.array(out, dtype=np.uint64)


def test_one_level_is_the_union_of_two_distinct_sorted_runs():
    rng = np.random.default_rng(1)
    for _ in range(1500):
        univ = np.unique(rng.integers(0, 60, size=rng.integers(0, 80)).astype(np.uint64))
        A = univ[rng.random(len(univ)) < rng.random()]
        B = univ[rng.random(len(univ)) < rng.random()]
        assert np.array_equal(union_level(A, B), np.union1d(A, B))


def test_ticket_order_is_round_robin_and_keeps_a_runs_tiles_in_order():
    """tile_order_kernel: ticket of tile (pair q, index lt) = sum_q' min(tiles[q'], lt) + #(q' < q with tiles[q'] > lt)"""
    rng = np.random.default_rng(2)
    for _ in range(500):
        cnt = rng.integers(0, 7, size=rng.integers(1, 9))
        pt = np.concatenate([[0], np.cumsum(cnt)])
        order = np.full(pt[-1], -1)
        for q, c in enumerate(cnt):
            for lt in range(c):
                t = int(np.minimum(cnt, lt).sum() + (cnt[:q] > lt).sum())
                assert order[t] == -1
                order[t] = pt[q] + lt
        assert (order >= 0).all()
        ticket = np.argsort(order)          # tile -> ticket
        for q, c in enumerate(cnt):         # the look-back waits for earlier tiles of the same run: they hold lower tickets
            assert (np.diff(ticket[pt[q]:pt[q] + c]) > 0).all()
        want = [pt[q] + lt for lt in range(int(cnt.max(initial=0))) for q in range(len(cnt)) if cnt[q] > lt]
        assert list(order) == want


def test_row_search_bounds_of_the_scatter():
    """a stream's records inside a row tile: record j of [a, e) has j - a of the stream's keys before it and e - 1 - j
    after it, all of them rows of the tile — the binary search runs on [j - a, rt - (e - j)]"""
    rng = np.random.default_rng(3)
    RT = 8
    for _ in range(1500):
        U = np.unique(rng.integers(0, 100, size=rng.integers(1, 60)).astype(np.uint64))
        S = U[rng.random(len(U)) < rng.random()]
        tiles = (len(U) + RT - 1) // RT
        split = [int(np.searchsorted(S, U[b * RT], side="left")) for b in range(tiles)] + [len(S)]
        for b in range(tiles):
            r0, rt = b * RT, min(RT, len(U) - b * RT)
            a, e = split[b], split[b + 1]
            for j in range(a, e):
                lo, hi = j - a, rt - (e - j)
                assert lo <= hi < rt
                while lo < hi:
                    mid = (lo + hi) >> 1
                    if S[j] <= U[r0 + mid]:
                        hi = mid
                    else:
                        lo = mid + 1
                assert U[r0 + lo] == S[j]
