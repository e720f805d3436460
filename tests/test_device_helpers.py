"""The kernels' pure helper functions, compiled for the host and called through the C ABI (no GPU): the same source
lines the device executes.  Covers the numeric rules parity rests on: fp32 -> bf16 rounding equals the oracle's, the
order-preserving keys are monotone, and the epilogue's list rule returns exactly the top-KL by (score desc, row asc)."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import bruteforce as bf


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_bf16_rounding_equals_the_oracle(lib):
    g = np.random.default_rng(0)
    x = np.concatenate([g.standard_normal(200_000).astype(np.float32) * np.float32(10.0) ** g.integers(-30, 30, 200_000).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.005859375, 1.0078125, np.inf, -np.inf, 3.3895314e38, 1e-45, -1e-45],
                                 dtype=np.float32)]).astype(np.float32)
    bits = np.empty(len(x), np.uint16)
    back = np.empty(len(x), np.float32)
    assert lib.sa_debug_bf16_round(ptr(x), len(x), ptr(bits), ptr(back)) == 0
    assert (bits == bf.f32_to_bf16_bits(x)).all()
    assert (back.view(np.uint32) == bf.bf16_bits_to_f32(bits).view(np.uint32)).all()
    nan = np.array([np.nan], np.float32)
    assert lib.sa_debug_bf16_round(ptr(nan), 1, ptr(bits), ptr(back)) == 0 and np.isnan(back[0])


def test_float_keys_are_order_preserving_and_invertible(lib):
    g = np.random.default_rng(1)
    x = np.concatenate([g.standard_normal(50_000).astype(np.float32) * np.float32(10.0) ** g.integers(-20, 20, 50_000).astype(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, -3.4e38], np.float32)]).astype(np.float32)
    key = np.empty(len(x), np.uint32); back = np.empty(len(x), np.float32); below = np.empty(len(x), np.float32)
    assert lib.sa_debug_float_keys(ptr(x), len(x), ptr(key), ptr(back), ptr(below)) == 0
    assert (back.view(np.uint32) == x.view(np.uint32)).all() and (key > 0).all()
    order = np.argsort(x, kind="stable")
    xs, ks = x[order], key[order].astype(np.uint64)
    assert ((np.diff(ks.astype(np.int64)) > 0) == (np.diff(xs) > 0))[np.diff(xs) != 0].all()     # x < y  <=>  key(x) < key(y)
    fin = np.isfinite(x)
    assert (below[fin] < x[fin]).all()
    assert (np.nextafter(below[fin & (x != 0)], np.float32(np.inf)) == x[fin & (x != 0)]).all()   # nothing in between


def test_merge_keys_order_by_score_then_lower_row(lib):
    g = np.random.default_rng(2)
    n = 20_000
    s = (np.round(g.standard_normal(n), 2) + 0.0).astype(np.float32)   # many exact score ties (and no -0.0: keys order bits)
    r = g.permutation(n).astype(np.int32)
    key = np.empty(n, np.uint64); rb = np.empty(n, np.int32)
    assert lib.sa_debug_merge_keys(ptr(s), ptr(r), n, ptr(key), ptr(rb)) == 0
    assert (rb == r).all()
    by_key = np.argsort(key)[::-1]
    ref = np.lexsort((r, -s))
    assert (by_key == ref).all()
    empty = np.array([-np.inf], np.float32); er = np.array([-1], np.int32)
    assert lib.sa_debug_merge_keys(ptr(empty), ptr(er), 1, ptr(key), ptr(rb)) == 0 and rb[0] == -1 and 0 < key[0] < key[1:].min()


def run_list(lib, s, rows, kl, floor=None, want_drop=False):
    s = np.ascontiguousarray(s, np.float32); rows = np.ascontiguousarray(rows, np.int32)
    out_s = np.empty(kl, np.float32); out_r = np.empty(kl, np.int32); drop = np.empty(1, np.float32)
    f = None if floor is None else ptr(np.ascontiguousarray(floor, np.float32))
    assert lib.sa_debug_list_insert(ptr(s), ptr(rows), len(s), kl, f, ptr(out_s), ptr(out_r), ptr(drop)) == 0
    return (out_s, out_r, drop[0]) if want_drop else (out_s, out_r)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.one_of(st.floats(width=32, allow_nan=False, allow_infinity=False, min_value=-4, max_value=4),
                          st.sampled_from([0.5, 0.25, -1.0, float("nan")])), min_size=0, max_size=120),
       st.sampled_from([16, 32]))
def test_list_rule_keeps_the_top_kl_with_ties_to_the_lower_row(lib, scores, kl):
    s = np.asarray(scores, np.float32)
    rows = np.arange(100, 100 + len(s), dtype=np.int32)                # rows arrive in ascending order, as in the scan
    got_s, got_r, drop = run_list(lib, s, rows, kl, want_drop=True)
    ok = ~np.isnan(s)                                                  # NaN (masked rows) never enters
    order = np.lexsort((rows[ok], -s[ok]))[:kl]
    exp_s = np.full(kl, -np.inf, np.float32); exp_r = np.full(kl, -1, np.int32)
    exp_s[:len(order)] = s[ok][order]; exp_r[:len(order)] = rows[ok][order]
    assert (got_r == exp_r).all() and (got_s.view(np.uint32) == exp_s.view(np.uint32)).all()
    # the "dropped" bound of the exactness certificate: the largest score seen and not held (-inf if nothing was dropped)
    rest = np.sort(s[ok])[::-1][kl:]
    assert drop == (rest[0] if len(rest) else -np.inf)


def test_shared_floor_drops_only_what_cannot_matter(lib):
    """A bound x published by another lane (KL rows there score >= x) may drop rows scoring < x but must keep ties."""
    g = np.random.default_rng(3)
    s = np.round(g.standard_normal(400), 1).astype(np.float32)
    rows = np.arange(400, dtype=np.int32)
    x = np.float32(1.0)
    floor = np.full(400, -np.inf, np.float32); floor[50] = x           # visible from the 32-value chunk holding value 50
    got_s, got_r, drop = run_list(lib, s, rows, 16, floor, want_drop=True)
    full_s, full_r = run_list(lib, s, rows, 16)
    keep = full_s >= x                                                 # everything at or above the bound is untouched
    assert (got_r[keep] == full_r[keep]).all() and (got_s[keep] == full_s[keep]).all()
    # exact semantics: rows seen after the bound became visible are admitted iff they score >= x (ties included)
    admit = (rows < 32) | (s >= x)
    exp_s, exp_r = run_list(lib, s[admit], rows[admit], 16)
    assert (got_r == exp_r).all() and (got_s == exp_s).all()
    # everything the list does not hold -- rejected by the floor included -- is covered by the dropped bound
    held = set(got_r.tolist())
    assert drop >= max(s[r] for r in range(400) if r not in held)
    assert ((s == x) & (rows >= 50)).sum() > 0                         # the data really contains ties with the bound


def test_window_bound_is_an_order_statistic_of_the_lanes_second_bests(lib):
    """The epilogue's 16-input network sorts any 16 keys, and the bound it returns is the (KL/2)-th largest -- with
    unpublished lanes (key 0) counting as 'nothing there yet'."""
    g = np.random.default_rng(5)
    w = g.integers(0, 2 ** 32, size=(4000, 16), dtype=np.uint64).astype(np.uint32)
    w[:500] = g.integers(0, 4, size=(500, 16)).astype(np.uint32)       # heavy ties and zeros
    w[500:1000, ::2] = 0                                               # half of the lanes have published nothing
    w[1000] = 0
    w[1001] = 0xFFFFFFFF
    for kl in (16, 32):
        out = np.empty(len(w), np.uint32); srt = np.empty_like(w)
        assert lib.sa_debug_window_bound(ptr(w), len(w), kl, ptr(out), ptr(srt)) == 0
        ref = -np.sort(-w.astype(np.int64), axis=1)
        assert (srt.astype(np.int64) == ref).all()
        assert (out.astype(np.int64) == ref[:, kl // 2 - 1]).all()
    assert out[1000] == 0 and (out[500:1000] == 0).all()               # KL = 32 needs all 16 lanes


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31), st.sampled_from([16, 32]), st.integers(1, 6))
def test_window_bound_never_exceeds_the_global_kl_th_best(lib, seed, kl, tiles):
    """Validity of the bound on simulated lanes: 16 lanes with disjoint rows, each holding the list rule's result over
    its rows so far; the bound built from their second bests never exceeds the KL-th best score over all their rows, so
    dropping rows below it cannot lose a member of the global top-KL (ties are admitted by the floor's >= rule)."""
    g = np.random.default_rng(seed)
    lanes = [np.round(g.standard_normal(tiles * 64), 2).astype(np.float32) for _ in range(16)]   # coarse: many ties
    second = np.array([np.sort(x)[::-1][1] for x in lanes], np.float32)
    key = np.empty(16, np.uint32); back = np.empty(16, np.float32); below = np.empty(16, np.float32)
    assert lib.sa_debug_float_keys(ptr(second), 16, ptr(key), ptr(back), ptr(below)) == 0
    out = np.empty(1, np.uint32)
    assert lib.sa_debug_window_bound(ptr(key), 1, kl, ptr(out), None) == 0
    bound = second[key == out[0]][0]
    allv = np.sort(np.concatenate(lanes))[::-1]
    assert bound <= allv[kl - 1]
    assert (allv >= bound).sum() >= kl


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2 ** 31), st.sampled_from([16, 32]), st.integers(1, 5))
def test_list_rule_under_arbitrary_shared_bounds(lib, seed, kl, n_bounds):
    """Any sequence of shared bounds (tightening or not, arriving at any chunk) leaves the invariants the exactness
    certificate rests on: (i) every value the list does not hold is <= the dropped bound; (ii) a value that is at or above
    every bound visible when it arrived, and belongs to the top-KL of all values, is in the list; (iii) the list is sorted
    (score desc, row asc) and holds no value below the tightest bound visible when that value arrived."""
    g = np.random.default_rng(seed)
    n = int(g.integers(1, 400))
    s = np.round(g.standard_normal(n), 1).astype(np.float32)               # coarse grid: plenty of ties
    s[g.random(n) < 0.05] = np.nan                                          # masked rows
    rows = np.arange(n, dtype=np.int32)
    floor = np.full(n, -np.inf, np.float32)
    for _ in range(n_bounds):
        floor[int(g.integers(0, n))] = np.float32(np.round(g.normal(0.5, 0.8), 1))
    got_s, got_r, drop = run_list(lib, s, rows, kl, floor, want_drop=True)
    # the bound in force for value i: the max of the bounds that became visible at or before the start of its chunk
    vis = np.full(n, -np.inf, np.float32)
    cur = -np.inf
    for c0 in range(0, n, 32):
        for i in range(c0, min(n, c0 + 32)):          # a bound attached to value i is applied when value i is staged ...
            if floor[i] > -np.inf:
                cur = max(cur, float(floor[i]))
            vis[i] = cur                               # ... i.e. before the chunk is processed, but in staging order
    # the hook applies floors while staging the chunk, so within a chunk every bound of that chunk is visible to all of it
    for c0 in range(0, n, 32):
        vis[c0:c0 + 32] = vis[min(n, c0 + 32) - 1]
    held = {int(r) for r in got_r if r >= 0}
    ok = ~np.isnan(s)
    rest = [float(s[i]) for i in range(n) if ok[i] and i not in held]
    assert all(v <= drop for v in rest), (max(rest) if rest else None, drop)                          # (i)
    order = np.lexsort((rows[ok], -s[ok]))
    top = set(rows[ok][order[:kl]].tolist())
    for i in top:
        if s[i] >= vis[i]:
            assert i in held, (i, float(s[i]), float(vis[i]))                                      # (ii)
    hs = [(float(x), int(r)) for x, r in zip(got_s, got_r) if r >= 0]
    assert hs == sorted(hs, key=lambda t: (-t[0], t[1]))                                             # (iii)
    assert all(float(s[r]) >= vis[r] for _, r in hs)
