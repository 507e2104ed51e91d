"""The reference's own search arithmetic (qdrant-client local mode) cannot run here and the reference ships no golden
vectors, so the oracle is "parity unpinned" (DESIGN.md §2).  What CAN be done offline is to check the oracle's cosine
ranking against independent third-party implementations of the same definition that ARE in this image: scikit-learn's
brute-force cosine k-NN / cosine_similarity and scipy's cdist.  This pins the arithmetic (normalise, dot, descending order,
threshold walk), not qdrant's tie order."""
import numpy as np
import pytest

from oracle import reverso_oracle as O

sk = pytest.importorskip("sklearn.metrics.pairwise")


@pytest.mark.parametrize("n,d,nq,k", [(2000, 64, 9, 10), (5000, 256, 4, 100), (300, 1024, 3, 500)])
def test_oracle_matches_sklearn_and_scipy(n, d, nq, k):
    from scipy.spatial.distance import cdist
    from sklearn.neighbors import NearestNeighbors
    rs = np.random.RandomState(n)
    db = rs.randn(n, d).astype(np.float32) * rs.uniform(0.1, 5.0, size=(n, 1)).astype(np.float32)   # un-normalised on purpose
    q = rs.randn(nq, d).astype(np.float32)
    got = O.search_batch(db, q, k, None)
    sim = sk.cosine_similarity(q.astype(np.float64), db.astype(np.float64))
    sim2 = 1.0 - cdist(q.astype(np.float64), db.astype(np.float64), metric="cosine")
    assert np.max(np.abs(sim - sim2)) < 1e-12
    nn = NearestNeighbors(n_neighbors=min(k, n), metric="cosine", algorithm="brute").fit(db.astype(np.float64))
    dist, idx = nn.kneighbors(q.astype(np.float64))
    for i, (ids, scores) in enumerate(got):
        kk = min(k, n)
        assert len(ids) == kk
        assert np.max(np.abs(scores - sim[i, ids])) < 2e-6                      # fp32 oracle vs fp64 third party
        assert np.all(np.diff(scores) <= 0)
        assert np.max(np.abs(scores - (1.0 - dist[i]))) < 2e-6                   # same sorted score list
        boundary = scores[-1]
        for j in set(ids.tolist()) ^ set(idx[i].tolist()):                       # sets may differ only at fp32 near-ties
            assert abs(sim[i, j] - boundary) < 2e-6


def test_oracle_threshold_walk_matches_a_filter_on_third_party_scores():
    rs = np.random.RandomState(3)
    db = rs.randn(4000, 96).astype(np.float32)
    q = db[:5] + 0.3 * rs.randn(5, 96).astype(np.float32)
    sim = sk.cosine_similarity(q.astype(np.float64), db.astype(np.float64))
    for thr in (0.2, 0.5, 0.9):
        for i, (ids, scores) in enumerate(O.search_batch(db, q, 50, thr)):
            want = np.sort(sim[i][sim[i] >= thr])[::-1][:50]
            assert abs(len(ids) - len(want)) <= 1                                 # a score within 1e-6 of thr may fall either way
            m = min(len(ids), len(want))
            assert m == 0 or np.max(np.abs(scores[:m] - want[:m])) < 2e-6
