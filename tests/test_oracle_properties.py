"""Property tests (hypothesis) of the oracle's search semantics and of the shard/merge decomposition the multi-GPU path relies on:
for ANY split of the rows into shards, merging the per-shard top-k lists (K3 semantics) equals the search over the whole DB."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import reverso_oracle as O
from revers_o_b200.sharded import shard_bounds


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 700), d=st.sampled_from([8, 33, 64]), nq=st.integers(1, 5), k=st.integers(1, 40),
       world=st.integers(1, 6), thr=st.sampled_from([None, -0.5, 0.0, 0.2, 0.9]), seed=st.integers(0, 10_000))
def test_shard_then_merge_equals_whole_search(n, d, nq, k, world, thr, seed):
    rs = np.random.RandomState(seed)
    db = O.round_to_bf16(O._cosine_prepare(rs.randn(n, d).astype(np.float32)))
    if n > 3:
        db[n - 1] = db[0]                                   # a cross-shard exact tie
    q = rs.randn(nq, d).astype(np.float32)
    whole = O.search_batch(db, q, k, thr, db_is_normalized=True)
    ids = np.full((world, nq, k), -1, np.int64)
    sc = np.full((world, nq, k), -np.inf, np.float32)
    cnt = np.zeros((world, nq), np.int32)
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        if hi > lo:
            for i, (a, b) in enumerate(O.search_batch(db[lo:hi], q, k, thr, db_is_normalized=True)):
                ids[r, i, : len(a)], sc[r, i, : len(a)], cnt[r, i] = a + lo, b, len(a)
    mi, ms, mc = O.merge_topk(ids, sc, cnt, k)
    for i, (a, b) in enumerate(whole):
        assert mc[i] == len(a)
        assert np.allclose(ms[i, : len(a)], b, atol=0, rtol=0)            # same fp32 dot products, same order of scores
        # ids may differ only inside groups of exactly equal scores that straddle the cut
        if set(mi[i, : len(a)].tolist()) != set(a.tolist()):
            assert len(a) == k and np.sum(b == b[-1]) >= 1
            assert set(mi[i, : len(a)][ms[i, : len(a)] > b[-1]].tolist()) == set(a[b > b[-1]].tolist())


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 400), d=st.sampled_from([4, 16, 50]), k=st.integers(1, 30), thr=st.floats(-1.0, 1.0), seed=st.integers(0, 10_000))
def test_threshold_walk_properties(n, d, k, thr, seed):
    """Results are descending, all >= threshold, at most `limit`, and exactly the prefix of the unthresholded ranking."""
    rs = np.random.RandomState(seed)
    db = rs.randn(n, d).astype(np.float32)
    q = rs.randn(d).astype(np.float32)
    ids, scores = O.search(db, q, k, thr)
    full_ids, full_scores = O.search(db, q, n, None)
    assert len(ids) <= min(k, n) and np.all(np.diff(scores) <= 0) and np.all(scores >= np.float32(thr))
    assert np.array_equal(scores, full_scores[: len(scores)])
    if len(ids) < min(k, n):
        assert full_scores[len(ids)] < np.float32(thr)                     # the walk stopped at the first score below it


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 150), d=st.sampled_from([8, 32]), thr=st.sampled_from([0.5, 0.9, 0.99]), seed=st.integers(0, 10_000))
def test_selfjoin_oracle_is_the_upper_triangle_of_the_gram_matrix(n, d, thr, seed):
    rs = np.random.RandomState(seed)
    db = O._cosine_prepare(rs.randn(n, d).astype(np.float32))
    db[n // 2] = db[0]
    pairs, scores = O.selfjoin_threshold(db, thr, db_is_normalized=True)
    g = db @ db.T
    want = {(i, j) for i in range(n) for j in range(i + 1, n) if g[i, j] >= thr}
    got = {(int(a), int(b)) for a, b in pairs}
    near = {(i, j) for i in range(n) for j in range(i + 1, n) if abs(g[i, j] - thr) < 1e-6}
    assert (got ^ want) <= near and (0, n // 2) in got
