"""A CPU model of the two-stage selection (per-lane approximate lists with shared thresholds -> union -> exact
rescoring), driven through the host-compiled list rule of the kernel (`sa_debug_list_insert`).  It checks the guarantee
DESIGN.md section 4.2 states -- the answer is exact unless more than kKL - k rows of one lane sit within the scan's
rounding error of the k-th score -- and shows that the bound is tight (a constructed violation does lose a row)."""
import ctypes as C

import numpy as np

KL = 16


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def lane_list(lib, approx, rows, floor=None):
    approx = np.ascontiguousarray(approx, np.float32)
    rows = np.ascontiguousarray(rows, np.int32)
    out_s = np.empty(KL, np.float32)
    out_r = np.empty(KL, np.int32)
    f = None if floor is None else ptr(np.ascontiguousarray(floor, np.float32))
    assert lib.sa_debug_list_insert(ptr(approx), ptr(rows), len(approx), KL, f, ptr(out_s), ptr(out_r)) == 0
    return out_s, out_r


def two_stage(lib, exact, approx, n_lanes, k, share=True):
    """exact / approx: per-row scores (float64 / float32).  Rows are dealt to lanes tile by tile (256 rows per tile, tile t
    -> lane t % n_lanes) as in the scan; lanes run one after the other here, each seeing the bound the previous ones
    published (any interleaving is allowed by the kernel; this is one of them)."""
    n = len(exact)
    tiles = np.arange(n) // 256
    cands_s, cands_r = [], []
    shared = -np.inf
    for lane in range(n_lanes):
        rows = np.flatnonzero(tiles % n_lanes == lane).astype(np.int32)
        floor = None
        if share and np.isfinite(shared) and len(rows):
            floor = np.full(len(rows), -np.inf, np.float32)
            floor[0] = shared                                  # visible from this lane's first value on
        s, r = lane_list(lib, approx[rows], rows, floor)
        if share and np.isfinite(s[KL - 1]):
            shared = max(shared, float(s[KL - 1]))             # atomicMax of the lane's KL-th best
        cands_s.append(s)
        cands_r.append(r)
    cs, cr = np.concatenate(cands_s), np.concatenate(cands_r)
    ok = cr >= 0
    cs, cr = cs[ok], cr[ok]
    top = np.lexsort((cr, -cs))[: 2 * KL]                      # merge: 2*KL best of the union by (approx desc, row asc)
    sel = cr[top]
    order = np.lexsort((sel, -exact[sel]))[:k]                 # exact rescoring, final order (exact desc, row asc)
    return sel[order]


def exact_topk(exact, k):
    return np.lexsort((np.arange(len(exact)), -exact))[:k]


def test_model_is_exact_under_fp32_scale_noise(lib):
    g = np.random.default_rng(0)
    for trial in range(20):
        n, n_lanes, k = int(g.integers(3000, 40000)), int(g.choice([1, 4, 18, 37, 148])), int(g.choice([1, 3, 10, 12]))
        exact = g.standard_normal(n) * 0.03
        approx = (exact * (1 + g.uniform(-1e-6, 1e-6, n))).astype(np.float32)       # the scan's rounding error
        for share in (True, False):
            got = two_stage(lib, exact, approx, n_lanes, k, share)
            assert (got == exact_topk(exact, k)).all(), (trial, n, n_lanes, k, share)


def test_model_survives_a_crowd_spread_over_lanes_and_shows_the_tight_bound(lib):
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
        assert (two_stage(lib, exact, approx, n_lanes, k) == exact_topk(exact, k)).all()
    # the documented limit: put MORE than kKL - k = 6 such rows beyond rank k into ONE tile (one lane): with the crowd of
    # 24 all in tile 0, that lane keeps only its 16 best by approximate score, and a true top-10 row can be among the 8 cut
    exact2 = g.standard_normal(n) * 0.03
    approx2 = exact2.astype(np.float32)
    rows = np.arange(24)
    exact2[rows] = 0.5 + np.arange(24) * 1e-9                  # true order: row 23 best ... row 0 worst
    approx2[rows] = np.float32(0.5) + (23 - np.arange(24)).astype(np.float32) * np.float32(6e-8)   # approx order reversed
    got = two_stage(lib, exact2, approx2, 18, k)
    assert not (got == exact_topk(exact2, k)).all()            # lost: the guarantee is tight, as DESIGN.md says
    assert set(got.tolist()).issubset(set(rows.tolist()))      # ... and what is returned is still from the crowd
