"""A CPU model of the search's three stages -- per-lane approximate lists with shared thresholds and "dropped" bounds (the
scan), certificate + exact re-scoring of the band candidates (the merge kernel), exact fallback scan of the ambiguous
lanes (the fixup kernel) -- driven through the host-compiled list rule of the kernel (`sa_debug_list_insert`, the very
source lines of csrc/sa_scan.cuh).  It checks the claim in csrc/sa_aux.cuh: the answer equals the exact top-k for ANY
approximation error up to eps, however the near-ties are placed -- including the one-tile crowd that the round-1 design
(fixed 2*kKL re-scoring, no certificate) provably lost."""
import ctypes as C

import numpy as np
import pytest

KL = 16
SEL_MAX = 128


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def lane_list(lib, approx, rows, floor=None):
    approx = np.ascontiguousarray(approx, np.float32)
    rows = np.ascontiguousarray(rows, np.int32)
    out_s = np.empty(KL, np.float32)
    out_r = np.empty(KL, np.int32)
    drop = np.empty(1, np.float32)
    f = None if floor is None else ptr(np.ascontiguousarray(floor, np.float32))
    assert lib.sa_debug_list_insert(ptr(approx), ptr(rows), len(approx), KL, f, ptr(out_s), ptr(out_r), ptr(drop)) == 0
    return out_s, out_r, float(drop[0])


def search_model(lib, exact, approx, eps, n_lanes, k, share=True, legacy=False):
    """exact / approx: per-row scores (float64 / float32) with |approx - exact| <= eps.  Rows are dealt to lanes tile by
    tile (256 rows per tile, tile t -> lane t % n_lanes) as in the scan; lanes run one after the other here, each seeing
    the bound the previous ones published (any interleaving is allowed by the kernel; this is one of them).
    Returns (rows of the answer, number of lanes sent to the fallback scan)."""
    n = len(exact)
    tiles = np.arange(n) // 256
    lists, drops = [], []
    shared = -np.inf
    for lane in range(n_lanes):
        rows = np.flatnonzero(tiles % n_lanes == lane).astype(np.int32)
        floor = None
        if share and np.isfinite(shared) and len(rows):
            floor = np.full(len(rows), -np.inf, np.float32)
            floor[0] = shared                                  # visible from this lane's first chunk on
        s, r, d = lane_list(lib, approx[rows], rows, floor)
        if share and np.isfinite(s[KL - 1]):
            shared = max(shared, float(s[KL - 1]))             # atomicMax of the lane's KL-th best
        lists.append((s, r))
        drops.append(d)
    cs = np.concatenate([s for s, _ in lists])
    cr = np.concatenate([r for _, r in lists])
    ok = cr >= 0
    cs, cr = cs[ok], cr[ok]
    if legacy:                                                 # round 1: 2*KL best of the union, no certificate
        sel = cr[np.lexsort((cr, -cs))[: 2 * KL]]
        return sel[np.lexsort((sel, -exact[sel]))[:k]], 0
    # ---- merge kernel: A_k, band, ambiguous lanes, band candidates of U
    order = np.lexsort((cr, -cs))
    a_k = cs[order[k - 1]] if len(order) >= k else None
    band = -np.inf if a_k is None else np.float32(a_k) - np.float32(2 * eps)
    amb = [d > -np.inf and d >= band for d in drops]
    sel = cr[cs >= band]
    if len(sel) > SEL_MAX:                                      # more than the merge kernel re-scores: every lane is rescanned
        amb = [True] * n_lanes
        sel = sel[:SEL_MAX]
    best = set(sel.tolist())
    # ---- fixup kernel: exact rescan of the ambiguous lanes, prefiltered by an approximate score with error <= eps
    for lane in np.flatnonzero(amb):
        rows = np.flatnonzero(tiles % n_lanes == lane)
        aprime = exact[rows]                                   # the CUDA-core dot: error far inside eps (taken as 0 here)
        best.update(rows[aprime >= band].tolist())
    cand = np.fromiter(best, dtype=np.int64)
    return cand[np.lexsort((cand, -exact[cand]))[:k]], int(np.sum(amb))


def exact_topk(exact, k):
    return np.lexsort((np.arange(len(exact)), -exact))[:k]


def test_model_is_exact_under_fp32_scale_noise(lib):
    g = np.random.default_rng(0)
    n_fix = 0
    for trial in range(20):
        n, n_lanes, k = int(g.integers(3000, 40000)), int(g.choice([1, 4, 18, 37, 148])), int(g.choice([1, 3, 10, 12]))
        exact = g.standard_normal(n) * 0.03
        eps = 2e-4
        approx = (exact + g.uniform(-eps, eps, n) * 0.01).astype(np.float32)          # the scan's real error is ~1 % of eps
        for share in (True, False):
            got, fixed = search_model(lib, exact, approx, eps, n_lanes, k, share)
            assert (got == exact_topk(exact, k)).all(), (trial, n, n_lanes, k, share)
            n_fix += fixed
    assert n_fix <= 4          # on iid data the certificate holds almost always: the fallback is the exception


@pytest.mark.parametrize("err", [0.01, 1.0])
def test_model_is_exact_for_any_error_up_to_eps(lib, err):
    """The certificate must hold for the WORST approximation the bound allows, not just the typical one."""
    g = np.random.default_rng(5)
    for trial in range(12):
        n, n_lanes, k = int(g.integers(2000, 20000)), int(g.choice([1, 3, 18, 64])), int(g.choice([1, 5, 10, 12]))
        exact = g.standard_normal(n) * 0.03
        eps = 2e-4
        approx = (exact + err * g.choice([-eps, eps], n) * 0.999).astype(np.float32)
        got, _ = search_model(lib, exact, approx, eps, n_lanes, k)
        assert (got == exact_topk(exact, k)).all(), (trial, n, n_lanes, k)


def test_crowds_wherever_they_sit(lib):
    g = np.random.default_rng(1)
    n, k = 30000, 10
    exact = g.standard_normal(n) * 0.03
    # 24 rows within 1e-7 of each other at the very top, scattered over the corpus (hence over the lanes)
    crowd = g.choice(n, 24, replace=False)
    exact[crowd] = 0.5 + g.uniform(0, 1e-7, 24)
    approx = exact.astype(np.float32)
    approx[crowd] = np.float32(0.5) + g.permutation(24).astype(np.float32) * np.float32(6e-8)   # fp32 order is noise
    assert not (np.lexsort((np.arange(n), -approx.astype(np.float64)))[:k] == exact_topk(exact, k)).all()  # fp32 alone fails
    for n_lanes in (4, 18, 148):
        got, _ = search_model(lib, exact, approx, 2e-4, n_lanes, k)
        assert (got == exact_topk(exact, k)).all()
    # the case round 1 lost: the whole crowd of 24 in ONE tile (one lane) with the approximate order reversed -- that lane
    # keeps only its 16 best by approximate score and cuts true top-10 rows.  The old selection is wrong ...
    exact2 = g.standard_normal(n) * 0.03
    approx2 = exact2.astype(np.float32)
    rows = np.arange(24)
    exact2[rows] = 0.5 + np.arange(24) * 1e-9                  # true order: row 23 best ... row 0 worst
    approx2[rows] = np.float32(0.5) + (23 - np.arange(24)).astype(np.float32) * np.float32(6e-8)   # approx order reversed
    old, _ = search_model(lib, exact2, approx2, 2e-4, 18, k, legacy=True)
    assert not (old == exact_topk(exact2, k)).all()
    # ... and the certificate sees that lane's dropped bound inside the band, rescans it exactly, and is right
    for n_lanes in (1, 18, 148):
        got, fixed = search_model(lib, exact2, approx2, 2e-4, n_lanes, k)
        assert (got == exact_topk(exact2, k)).all() and fixed >= 1
    # 300 exact duplicates of the best row: more band candidates than the merge kernel re-scores -> every lane rescanned
    exact3 = g.standard_normal(n) * 0.03
    dup = g.choice(n, 300, replace=False)
    exact3[dup] = 0.4
    got, fixed = search_model(lib, exact3, exact3.astype(np.float32), 2e-4, 18, k)
    assert (got == np.sort(dup)[:k]).all() and fixed == 18      # ties resolve to the lowest rows


def window_floors(lib, approx, n_lanes):
    """Per lane, the floor each of its rows sees under the window bound, for a lock-step interleaving (every lane finishes
    tile t before anybody starts tile t+1): lane L's bound for its tile t is the (KL/2)-th largest, over the 16 lanes
    starting at L, of their second-best score after their tile t-1 -- computed by the kernel's own network
    (sa_debug_window_bound over sa_debug_float_keys keys).  The second bests used are the TRUE ones of each lane's prefix,
    at least as large as what a lane that already drops rows would publish: a tighter bound than the kernel's, still valid
    (the rows exist), and so a harder case for the certificate."""
    n = len(approx)
    tiles = np.arange(n) // 256
    lane_rows = [np.flatnonzero(tiles % n_lanes == lane).astype(np.int32) for lane in range(n_lanes)]
    n_t = max((len(r) + 255) // 256 for r in lane_rows)
    sb = np.full((n_lanes, n_t), -np.inf, np.float32)          # second best after tile t (lane-local tile index)
    for lane, rows in enumerate(lane_rows):
        for t in range(n_t):
            pre = approx[rows[: (t + 1) * 256]]
            if len(pre) >= 2:
                sb[lane, t] = np.partition(pre, -2)[-2]
    floors = []
    for lane, rows in enumerate(lane_rows):
        f = np.full(len(rows), -np.inf, np.float32)
        for t in range(1, (len(rows) + 255) // 256):
            win = np.ascontiguousarray(sb[[(lane + i) % n_lanes for i in range(16)], t - 1])
            key = np.zeros(16, np.uint32); back = np.empty(16, np.float32); below = np.empty(16, np.float32)
            fin = np.isfinite(win)
            if fin.any():
                k2 = np.empty(int(fin.sum()), np.uint32)
                assert lib.sa_debug_float_keys(ptr(np.ascontiguousarray(win[fin])), int(fin.sum()), ptr(k2), ptr(back), ptr(below)) == 0
                key[fin] = k2                                   # lanes with fewer than two rows so far: key 0 = unpublished
            out = np.empty(1, np.uint32)
            assert lib.sa_debug_window_bound(ptr(key), 1, KL, ptr(out), None) == 0
            if out[0] != 0:
                f[t * 256] = win[fin][k2 == out[0]][0]
        floors.append(f)
    return lane_rows, floors


@pytest.mark.parametrize("n_lanes", [16, 18, 37, 148])
def test_model_is_exact_with_the_window_bound(lib, n_lanes):
    """The scan with the window bound in force (every lane's threshold floored by kKL/2 lanes' second bests), then the same
    merge / certificate / fallback: exact for iid data, for error at the edge of eps, and for crowds of near-ties --
    scattered (most lanes' second best EQUALS the score to keep) and packed into one tile."""
    g = np.random.default_rng(100 + n_lanes)
    eps, k = 2e-4, 10
    cases = []
    for trial in range(4):
        n = int(g.integers(n_lanes * 256 * 2, n_lanes * 256 * 6))
        exact = g.standard_normal(n) * 0.03
        cases.append((exact, (exact + g.choice([-eps, eps], n) * (0.999 if trial % 2 else 0.01)).astype(np.float32)))
    n = n_lanes * 256 * 4
    exact = g.standard_normal(n) * 0.03
    exact[g.choice(n, 200, replace=False)] = 0.4               # duplicates everywhere
    cases.append((exact, exact.astype(np.float32)))
    exact = g.standard_normal(n) * 0.03
    exact[5000:5024] = 0.5 + np.arange(24) * 1e-9              # the one-tile crowd, approximate order reversed
    approx = exact.astype(np.float32)
    approx[5000:5024] = np.float32(0.5) + (23 - np.arange(24)).astype(np.float32) * np.float32(6e-8)
    cases.append((exact, approx))
    tight = 0
    for ci, (exact, approx) in enumerate(cases):
        n = len(exact)
        tiles = np.arange(n) // 256
        lane_rows, floors = window_floors(lib, approx, n_lanes)
        lists, drops = [], []
        for rows, f in zip(lane_rows, floors):
            s, r, d = lane_list(lib, approx[rows], rows, f)
            lists.append((s, r)); drops.append(d)
            tight += int(np.isfinite(f).any())
        cs = np.concatenate([s for s, _ in lists]); cr = np.concatenate([r for _, r in lists])
        ok = cr >= 0
        cs, cr = cs[ok], cr[ok]
        # the bound is valid: the union still holds the top-KL by approximate score (ties aside, at least their scores)
        top = np.sort(approx)[::-1][:KL]
        assert (np.sort(cs)[::-1][:KL] == top).all(), ci
        order = np.lexsort((cr, -cs))
        band = np.float32(cs[order[k - 1]]) - np.float32(2 * eps)
        amb = [d > -np.inf and d >= band for d in drops]
        sel = cr[cs >= band]
        if len(sel) > SEL_MAX:
            amb, sel = [True] * n_lanes, sel[:SEL_MAX]
        best = set(sel.tolist())
        for lane in np.flatnonzero(amb):
            rows = np.flatnonzero(tiles % n_lanes == lane)
            best.update(rows[exact[rows] >= band].tolist())
        cand = np.fromiter(best, dtype=np.int64)
        got = cand[np.lexsort((cand, -exact[cand]))[:k]]
        assert (got == exact_topk(exact, k)).all(), (ci, n_lanes)
    assert tight > 0                                            # the window bound really was in force
