"""CPU tests of the oracle itself (oracle/bruteforce.py): the fast variant and the sgemm baseline are pinned
against the float64 definition, conversions against torch, tie / zero-row / short-corpus rules, shard merge."""
import numpy as np
import pytest
import torch

from oracle import bruteforce as bf


def test_bf16_roundtrip_matches_torch():
    g = np.random.default_rng(0)
    x = np.concatenate([g.standard_normal(4096).astype(np.float32) * 10.0 ** g.integers(-20, 20, 4096),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.005859375, np.inf, -np.inf, 3.3895314e38],
                                 dtype=np.float32)]).astype(np.float32)
    mine = bf.f32_to_bf16_bits(x)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert (mine == ref).all()
    back = bf.bf16_bits_to_f32(mine)
    assert (bf.f32_to_bf16_bits(back) == mine).all()


def test_synth_is_deterministic_and_unnormalised():
    a = bf.synth_rows(1234, 3, 100, 64)
    b = bf.synth_rows(1234, 3, 100, 64)
    assert (a == b).all()
    n = np.linalg.norm(bf.bf16_bits_to_f32(a).astype(np.float64), axis=1)
    assert n.std() / n.mean() > 0.2


@pytest.mark.parametrize("n,dim,nq,k", [(5000, 256, 33, 10), (1000, 1536, 8, 5), (70000, 128, 16, 10)])
def test_fast_equals_definition(n, dim, nq, k):
    c = bf.synth_rows(7, 0, n, dim)
    q = bf.synth_queries(9, nq, dim, c)
    s, i = bf.cosine_topk_f64(q, c, k)
    chunks = [(lo, c[lo:lo + 4096]) for lo in range(0, n, 4096)]
    s2, i2 = bf.cosine_topk_fast(q, chunks, k)
    assert (i == i2).all()
    assert np.abs(s - s2).max() < 1e-14
    s3, i3 = bf.cosine_topk_sgemm(q, chunks, k)
    rep = bf.compare_topk(i3, s3, i, s, tie_tol=2e-6)
    assert rep["recall"] > 0.99 and rep["max_abs_dscore"] < 1e-5


def test_planted_queries_have_known_top1():
    c = bf.synth_rows(1234, 0, 3000, 256)
    q = bf.synth_queries(4321, 40, 256, c)
    _, i = bf.cosine_topk_f64(q, c, 3)
    for r in range(1, 40, 2):
        assert i[r, 0] == bf.planted_row(r, 3000)


def test_ties_zero_rows_and_short_corpus():
    dim = 64
    c = bf.synth_rows(5, 0, 20, dim)
    c[7] = c[3]            # exact duplicate: tie -> lower row first
    c[11] = 0              # all-zero row: never returned
    q = c[3:4].copy()
    s, i = bf.cosine_topk_f64(q, c, 25)
    assert i[0, 0] == 3 and i[0, 1] == 7 and s[0, 0] == s[0, 1]
    assert 11 not in i[0]
    assert (i[0, 19:] == -1).all() and np.isneginf(s[0, 19:]).all()   # 19 eligible rows, k = 25
    # scale invariance of cosine: doubling a row (exact in bf16) leaves its score unchanged
    c2 = c.copy()
    c2[5] = bf.f32_to_bf16_bits(bf.bf16_bits_to_f32(c[5]) * 2)
    s2, i2 = bf.cosine_topk_f64(q, c2, 25)
    assert (i2 == i).all() and np.abs(s2 - s)[np.isfinite(s)].max() < 1e-15
    # zero query: score 0 everywhere, lowest rows win
    z = np.zeros((1, dim), dtype=np.uint16)
    s3, i3 = bf.cosine_topk_f64(z, c, 4)
    assert i3[0].tolist() == [0, 1, 2, 3] and (s3 == 0).all()


def test_shard_merge_equals_global():
    n, dim, nq, k = 6000, 128, 12, 10
    c = bf.synth_rows(3, 0, n, dim)
    c[4500] = c[100]
    q = bf.synth_queries(4, nq, dim, c)
    s, i = bf.cosine_topk_f64(q, c, k)
    cuts = [0, 1500, 3000, 4600, n]
    ss, ii = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        x, y = bf.cosine_topk_f64(q, c[a:b], k)
        ss.append(x)
        ii.append(y)
    ms, mi = bf.merge_shard_topk(ss, ii, cuts[:-1], k)
    assert (mi == i).all() and np.abs(ms - s).max() < 1e-14


def test_oracle_agrees_with_independent_cosine_implementations():
    """The reference holds no golden vector for this arithmetic (parity unpinned), so the definition is at least
    cross-checked against two independent, widely used implementations of cosine similarity -- scipy's
    ``cdist(metric="cosine")`` and scikit-learn's ``cosine_similarity`` -- on the same bf16-rounded vectors, and against
    torch's float64 ``topk`` for the ranking."""
    from scipy.spatial.distance import cdist
    from sklearn.metrics.pairwise import cosine_similarity
    n, dim, nq, k = 3000, 1536, 12, 10
    c = bf.synth_rows(1234, 0, n, dim)
    q = bf.synth_queries(4321, nq, dim, c)
    s, i = bf.cosine_topk_f64(q, c, k)
    c64 = bf.bf16_bits_to_f32(c).astype(np.float64)
    q64 = bf.bf16_bits_to_f32(q).astype(np.float64)
    sim_scipy = 1.0 - cdist(q64, c64, metric="cosine")
    sim_sklearn = cosine_similarity(q64, c64)
    for sim in (sim_scipy, sim_sklearn):
        assert np.abs(np.take_along_axis(sim, i, axis=1) - s).max() < 1e-12
        order = np.argsort(-sim, axis=1, kind="stable")[:, :k]          # stable: ties keep the lower row first
        assert (order == i).all()
    ts, ti = torch.topk(torch.from_numpy(sim_scipy), k, dim=1)
    assert (ti.numpy() == i).all()


def test_c_topk_helper_equals_numpy_selection():
    """oracle/topk.c (the multi-core selection of the CPU baseline) against the numpy path: ties to the lower row,
    -inf rows never returned, running merge across chunks, k larger than the rows available."""
    import __graft_entry__ as ge
    ge.build_oracle_helper()
    bf._TOPK_LIB = None
    assert bf._topk_lib() is not None
    c = bf.synth_rows(7, 0, 30000, 128)
    c[17] = 0
    c[900] = c[5]; c[20001] = c[5]
    q = bf.synth_queries(9, 40, 128, c)
    q[0] = c[5]
    for cuts in ([0, 30000], [0, 11000, 20500, 30000], [0, 7, 30000]):
        prep = bf.prepare_chunks_f32([(a, c[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        for k in (1, 10, 28):
            s1, i1 = bf.cosine_topk_sgemm_prepared(q, prep, k, use_c_topk=False)
            s2, i2 = bf.cosine_topk_sgemm_prepared(q, prep, k, use_c_topk=True)
            assert (i1 == i2).all() and (s1 == s2).all()
    assert i2[0, :3].tolist() == [5, 900, 20001] and 17 not in i2
    tiny = bf.prepare_chunks_f32([(0, c[:4])])
    s3, i3 = bf.cosine_topk_sgemm_prepared(q[:3], tiny, 10, use_c_topk=True)
    assert (i3[:, 4:] == -1).all() and np.isneginf(s3[:, 4:]).all() and (np.sort(i3[:, :4], axis=1) == np.arange(4)).all()
