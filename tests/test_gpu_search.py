"""GPU parity tests for K2 (search) and K3 (merge): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Tolerances are the north_star's: scores within 1e-3 absolute, identical
top-k id sets except ties within 1e-3."""
import math

import numpy as np
import pytest
import torch

from oracle import reverso_oracle as O
from conftest import assert_topk_match

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def _oracle(db_bf16: torch.Tensor, n, d, queries: torch.Tensor, k, thr=None):
    from revers_o_b200 import ops
    dbf = ops.untile_rows(db_bf16, n, d).float().cpu().numpy()
    return O.search_batch(dbf, queries.float().cpu().numpy(), k, thr, db_is_normalized=True)


def _run(db, n, d, q, k, thr=None, id_offset=0):
    from revers_o_b200 import ops
    ids, sc, cnt = ops.search_topk(db, n, d, q, k, thr, id_offset)
    torch.cuda.synchronize()
    return ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()


# ---- the tcgen05 mainloop by itself ---------------------------------------------------------------
@pytest.mark.parametrize("n,d,nq,stride", [(1000, 64, 5, 1), (4096, 1024, 16, 1), (5000, 1024, 37, 1),
                                            (3000, 1280, 256, 1), (20000, 1024, 300, 7), (777, 96, 130, 1),
                                            (9000, 1024, 200, 3)])
def test_dense_scores_match_fp32_reference(dev, n, d, nq, stride):
    from revers_o_b200 import ops, synth
    q = synth.make_queries(nq, d, seed=3, device=dev)
    db = synth.make_db(n, d, q, n_plant=8, seed=5, device=dev)
    got = ops.scores_dense(db, n, d, q, tile_stride=stride)
    torch.cuda.synchronize()
    T = ops.scan_tile_rows(nq, d)
    tiles = (n + T - 1) // T
    rows = torch.cat([torch.arange(t * T, (t + 1) * T) for t in range(0, tiles, stride)]).to(dev)
    qn = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).float()          # tensor path rounds the query
    ok = rows < n
    ref = torch.full((nq, rows.numel()), -math.inf, device=dev)
    ref[:, ok] = qn @ ops.untile_rows(db, n, d, rows[ok]).float().T
    assert got.shape == ref.shape
    assert torch.equal(torch.isinf(got), torch.isinf(ref))
    assert torch.max(torch.abs(got[:, ok] - ref[:, ok])).item() < 2e-5


# ---- full search, both paths ------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,nq,k", [
    (10_000, 1024, 1, 10),      # BASELINE config 0: the reference's own operating point
    (10_000, 1024, 3, 10),
    (50_000, 1280, 4, 100),
    (3_000, 96, 2, 10),
    (2_500, 1024, 16, 10),      # tensor path, shard below one chunk (single DENSE level)
    (10_000, 1024, 9, 10),      # tensor path, resident queries
    (60_000, 1024, 64, 100),
    (120_000, 1024, 256, 100),  # config 1 shape at reduced N
    (70_000, 1280, 300, 50),    # two query blocks
    (33_333, 2048, 40, 7),
    (50_000, 64, 40, 512),      # maximum k, one k-chunk
    (300_000, 128, 1000, 10),   # four query blocks on CTA pairs
    (20_000, 1280, 130, 100),   # PE-Core-G14 width, odd query count
])
def test_search_matches_oracle(dev, n, d, nq, k):
    from revers_o_b200 import synth
    q = synth.make_queries(nq, d, seed=7, device=dev)
    db = synth.make_db(n, d, q, n_plant=min(128, 2 * k), seed=1000, device=dev)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert_topk_match(ids, sc, cnt, _oracle(db, n, d, q, k), k, TOL, f"n{n}d{d}q{nq}k{k}")
    assert np.all(cnt == min(k, n))


def test_many_query_blocks_and_intermediate_level(dev):
    """3 query blocks (600 queries) over 1.7M rows: DENSE seed + one sampled FILTER level + the full scan.  Every
    query must come back unflagged (candidate sub-lists must fill evenly across query blocks) and exact."""
    from revers_o_b200 import ops, synth
    n, d, nq, k = 1_700_000, 256, 600, 100
    q = synth.make_queries(nq, d, seed=51, device=dev)
    db = synth.make_db(n, d, q, n_plant=128, seed=52, device=dev)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert np.all(cnt == k), f"{int(np.sum(cnt < 0))} queries overflowed"
    sel = list(range(0, nq, 50))
    dbf = ops.untile_rows(db, n, d).float().cpu().numpy()
    ref = O.search_batch(dbf, q[sel].cpu().numpy(), k, None, db_is_normalized=True)
    assert_topk_match(ids[sel], sc[sel], cnt[sel], ref, k, TOL, "multiblock")


def test_scores_also_within_tolerance_of_unrounded_fp32_db(dev):
    """The GPU DB is bf16; the reference's is fp32.  North-star tolerance (1e-3) must hold against the fp32 DB too."""
    from revers_o_b200 import ops
    rs = np.random.RandomState(0)
    n, d, nq, k = 20_000, 1024, 32, 50
    dbf = O._cosine_prepare(rs.randn(n, d).astype(np.float32))
    qf = rs.randn(nq, d).astype(np.float32)
    for i in range(nq):
        for j, a in enumerate(np.linspace(0.5, 0.99, 60)):
            v = a * qf[i] / np.linalg.norm(qf[i]) + math.sqrt(1 - a * a) * rs.randn(d).astype(np.float32) / math.sqrt(d)
            dbf[(i * 601 + j * 7) % n] = v / np.linalg.norm(v)
    src = torch.from_numpy(dbf).to(dev)
    db, _ = ops.normalize_rows(src, db=ops.db_alloc(n, d, dev))
    ids, sc, cnt = _run(db, n, d, torch.from_numpy(qf).to(dev), k)
    ref = O.search_batch(dbf, qf, k, None, db_is_normalized=True)
    assert_topk_match(ids, sc, cnt, ref, k, TOL, "fp32db")


@pytest.mark.parametrize("nq", [1, 24])
def test_score_threshold_semantics(dev, nq):
    from revers_o_b200 import synth
    n, d, k = 30_000, 1024, 50
    q = synth.make_queries(nq, d, seed=11, device=dev)
    db = synth.make_db(n, d, q, n_plant=64, seed=12, device=dev)
    for thr in (0.7, 0.95, 0.999, -1.0):
        ids, sc, cnt = _run(db, n, d, q, k, thr)
        ref = _oracle(db, n, d, q, k, thr)
        assert_topk_match(ids, sc, cnt, ref, k, TOL, f"thr{thr}")
        for qi in range(nq):
            assert np.all(sc[qi, : cnt[qi]] >= thr) and np.all(ids[qi, cnt[qi]:] == -1)


@pytest.mark.parametrize("nq", [1, 8])
def test_known_answers_identity_duplicates_limit(dev, nq):
    """KA1/KA2/KA4/KA5 through the CUDA path."""
    from revers_o_b200 import ops
    d = 128
    eye = torch.eye(d, dtype=torch.float32, device=dev)
    src = torch.cat([eye, eye[5:6], eye[5:6]], 0)                 # rows 128,129 duplicate row 5
    n = src.shape[0]
    db, _ = ops.normalize_rows(src, db=ops.db_alloc(n, d, dev))
    q = eye[[5] + list(range(1, nq))].contiguous() * 3.0
    ids, sc, cnt = _run(db, n, d, q, 3)
    assert ids[0].tolist() == [5, 128, 129] and np.allclose(sc[0], 1.0, atol=1e-6)   # ties -> lower id first
    ids, sc, cnt = _run(db, n, d, q, 200)                          # limit > N
    assert cnt[0] == n and np.all(ids[0, n:] == -1) and sorted(ids[0, :n].tolist()) == list(range(n))
    ids, sc, cnt = _run(db, n, d, -q, 5, 0.5)                      # nothing above the threshold
    assert np.all(cnt == 0) and np.all(ids == -1) and np.all(np.isneginf(sc))
    ids, sc, cnt = _run(db, 0, d, q, 5)                            # empty collection
    assert np.all(cnt == 0) and np.all(ids == -1)


def test_id_offset_and_virtual_shards_merge(dev, golden):
    """K3 on one GPU: split the DB into G virtual shards, search each with its id_offset, merge; the result must
    equal the unsharded search (merge is a pure function of the gathered lists) and the oracle's merge."""
    from revers_o_b200 import ops, synth
    from revers_o_b200.sharded import shard_bounds, pack_results
    n, d, nq, k, G = 90_000, 1024, 40, 100, 4
    q = synth.make_queries(nq, d, seed=21, device=dev)
    db = synth.make_db(n, d, q, n_plant=160, seed=22, device=dev)
    full = _run(db, n, d, q, k)
    parts = []
    for r in range(G):
        lo, hi = shard_bounds(n, G, r)
        parts.append(ops.search_topk(db[lo // 128: (hi + 127) // 128], hi - lo, d, q, k, None, lo))
    ids = torch.stack([p[0] for p in parts]); sc = torch.stack([p[1] for p in parts]); cnt = torch.stack([p[2] for p in parts])
    mi, ms, mc = ops.merge_topk(ids.contiguous(), sc.contiguous(), cnt.contiguous(), k)
    torch.cuda.synchronize()
    assert np.array_equal(mi.cpu().numpy(), full[0]) and np.allclose(ms.cpu().numpy(), full[1], atol=1e-6)
    oi, os_, oc = O.merge_topk(ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy(), k)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mc.cpu().numpy(), oc)
    # packed layout (what one all-gather delivers)
    from revers_o_b200 import _lib
    blob = torch.stack([pack_results(*p) for p in parts])
    pi = torch.empty_like(mi); ps = torch.empty_like(ms); pc = torch.empty_like(mc)
    _lib.check(_lib.load().rvo_merge_topk_packed(blob.data_ptr(), blob.stride(0), G, nq, k, pi.data_ptr(), ps.data_ptr(),
                                                 pc.data_ptr(), torch.cuda.current_stream().cuda_stream), "packed")
    torch.cuda.synchronize()
    assert torch.equal(pi, mi) and torch.equal(ps, ms) and torch.equal(pc, mc)
    # golden merge fixture
    gi, gs, gc = (torch.from_numpy(golden[f"merge_{x}"]).to(dev) for x in ("ids", "scores", "counts"))
    a, b, c = ops.merge_topk(gi, gs, gc, 5)
    assert np.array_equal(a.cpu().numpy(), golden["merge_out_ids"]) and np.array_equal(c.cpu().numpy(), golden["merge_out_counts"])


def test_golden_search_fixture(dev, golden):
    from revers_o_b200 import ops
    db = torch.from_numpy(golden["search_db_bf16_bits"].astype(np.int16)).view(torch.bfloat16)
    n, d = db.shape
    dbp = ops.tile_rows(db.to(dev))
    q = torch.from_numpy(golden["search_queries"]).to(dev)
    for tag, k, thr in (("k10", 10, None), ("k10_t07", 10, 0.7), ("k100", 100, None)):   # nq=9 -> tensor path
        ids, sc, cnt = _run(dbp, n, d, q, k, thr)
        assert np.array_equal(cnt, golden[f"search_{tag}_counts"])
        for i in range(q.shape[0]):
            c = cnt[i]
            assert np.allclose(sc[i, :c], golden[f"search_{tag}_scores"][i, :c], atol=2e-6)
            assert set(ids[i, :c].tolist()) == set(golden[f"search_{tag}_ids"][i, :c].tolist())
    for i in range(0, 9, 4):                                      # small-q path on the same fixture
        ids, sc, cnt = _run(dbp, n, d, q[i:i + 4].contiguous(), 10)
        for j in range(cnt.shape[0]):
            assert np.allclose(sc[j, :10], golden["search_k10_scores"][i + j], atol=2e-6)


def test_overflow_flag_and_exact_fallback(dev):
    """A DB made of >cap exact duplicates of the query overflows the fused path's candidate buffers: the ABI must
    flag it (-1), and the documented protocol (re-run in batches of <= RVO_SMALL_Q) must return the exact answer."""
    from revers_o_b200 import ops, synth
    n, d, nq, k = 40_000, 256, 8, 20
    q = synth.make_queries(nq, d, seed=31, device=dev)
    rows = ops.untile_rows(synth.make_db(n, d, None, seed=32, device=dev), n, d).clone()
    rows[5000:30000] = (q[0] / q[0].norm()).to(torch.bfloat16)        # 25k identical rows, all score ~1.0 for query 0
    db = ops.tile_rows(rows)
    ids, sc, cnt = ops.search_topk(db, n, d, q, k)
    torch.cuda.synchronize()
    assert cnt[0].item() == -1 and torch.all(cnt[1:] == k)
    ids, sc, cnt = ops.search_topk_exact(db, n, d, q, k)
    torch.cuda.synchronize()
    assert cnt[0].item() == k and ids[0].tolist() == list(range(5000, 5000 + k))      # ties -> lowest ids
    ref = _oracle(db, n, d, q, k)
    assert_topk_match(ids.cpu().numpy()[1:], sc.cpu().numpy()[1:], cnt.cpu().numpy()[1:], ref[1:], k, TOL, "ovf")


def test_clustered_ingest_order_is_still_exact(dev):
    """Adversarial row order for the strided threshold samples: all near neighbours sit in one contiguous run."""
    from revers_o_b200 import synth
    n, d, nq, k = 150_000, 1024, 32, 100
    q = synth.make_queries(nq, d, seed=41, device=dev)
    from revers_o_b200 import ops
    rows = ops.untile_rows(synth.make_db(n, d, None, seed=42, device=dev), n, d).clone()
    qn = q / q.norm(dim=1, keepdim=True)
    g = torch.Generator(device=dev).manual_seed(43)
    for i in range(nq):
        noise = torch.randn((300, d), generator=g, device=dev) / math.sqrt(d)
        a = torch.linspace(0.6, 0.99, 300, device=dev).view(-1, 1)
        v = a * qn[i] + torch.sqrt(1 - a * a) * noise
        rows[70_000 + i * 300: 70_000 + (i + 1) * 300] = (v / v.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    db = ops.tile_rows(rows)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert_topk_match(ids, sc, cnt, _oracle(db, n, d, q, k), k, TOL, "clustered")


def test_config1_full_size_properties_and_sampled_parity(dev):
    """BASELINE config 1 at full size (1M x 1024, Q=256, k=100): size-independent properties on every query
    (descending, in-range unique ids, every returned score reproduced by an fp32 dot with the stored row, the
    planted >=0.99 neighbour found first) plus oracle parity on a sample of queries."""
    from revers_o_b200 import ops, synth
    n, d, nq, k = 1_000_000, 1024, 256, 100
    q = synth.make_queries(nq, d, seed=7, device=dev)
    db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert np.all(cnt == k)
    assert np.all(np.diff(sc, axis=1) <= 1e-7) and ids.min() >= 0 and ids.max() < n
    assert all(len(set(r.tolist())) == k for r in ids)
    qn = (q / q.norm(dim=1, keepdim=True))
    rows = ops.untile_rows(db, n, d, torch.from_numpy(ids).to(dev).view(-1)).float().view(nq, k, d)
    redo = torch.einsum("qkd,qd->qk", rows, qn).cpu().numpy()
    assert np.max(np.abs(redo - sc)) < 1e-5
    assert np.all(sc[:, 0] > 0.98)
    sel = list(range(0, nq, 16))
    dbf = ops.untile_rows(db, n, d).float().cpu().numpy()
    ref = O.search_batch(dbf, q[sel].cpu().numpy(), k, None, db_is_normalized=True)
    assert_topk_match(ids[sel], sc[sel], cnt[sel], ref, k, TOL, "cfg1")


# ---- tightly packed top-k (what an unstructured DB looks like: many rows within the bf16 margin of the k-th) ----
def _cluster_db(dev, n, d, q, lo, hi, m, seed):
    """Background rows plus, for EVERY query, `m` rows whose cosine to it is spread over [lo, hi]."""
    from revers_o_b200 import ops, synth
    db = synth.make_db(n, d, None, seed=seed, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed + 7)
    nq = q.shape[0]
    qn = (q / q.norm(dim=1, keepdim=True)).to(dev)
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(seed + 9))[: nq * m].view(nq, m).to(dev)
    alpha = torch.linspace(lo, hi, m, device=dev).view(1, m, 1)
    for i in range(nq):
        noise = torch.randn((1, m, d), generator=g, device=dev)
        qq = qn[i].view(1, 1, d)
        noise = noise - (noise * qq).sum(-1, keepdim=True) * qq
        noise = noise / noise.norm(dim=-1, keepdim=True)
        v = alpha * qq + torch.sqrt(1 - alpha * alpha) * noise
        vb = torch.zeros((m, db.shape[1] * 64), dtype=torch.bfloat16, device=dev)
        vb[:, :d] = (v / v.norm(dim=-1, keepdim=True)).view(m, d).to(torch.bfloat16)
        r = rows[i]
        db[r // 128, :, r % 128, :] = vb.view(m, db.shape[1], 64)
    return db


@pytest.mark.parametrize("m,nq", [(600, 12), (1500, 40)])
def test_many_candidates_inside_the_margin(dev, m, nq):
    """m rows per query within 0.002 of each other at the top: far more than k candidates lie inside the bf16 admission
    margin of the k-th score.  The fused select/re-score keeps up to 2048 of them, so the result stays exact and no
    query is flagged."""
    from revers_o_b200 import synth
    n, d, k = 200_000, 1024, 100
    q = synth.make_queries(nq, d, seed=61, device=dev)
    db = _cluster_db(dev, n, d, q, 0.900, 0.902, m, seed=62)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert np.all(cnt == k), f"{int(np.sum(cnt < 0))} queries flagged"
    assert_topk_match(ids, sc, cnt, _oracle(db, n, d, q, k), k, TOL, f"cluster{m}")


def test_unplanted_background_top_k(dev):
    """No planted neighbours at all: the top-100 of 3M random rows sit ~0.004 apart in total (the densest score
    distribution a DB can have at this size).  Every query must come back exact and unflagged."""
    from revers_o_b200 import ops, synth
    n, d, nq, k = 3_000_000, 256, 48, 100
    q = synth.make_queries(nq, d, seed=71, device=dev)
    db = synth.make_db(n, d, None, seed=72, device=dev)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert np.all(cnt == k), f"{int(np.sum(cnt < 0))} queries flagged"
    sel = list(range(0, nq, 6))
    dbf = ops.untile_rows(db, n, d).float().cpu().numpy()
    ref = O.search_batch(dbf, q[sel].cpu().numpy(), k, None, db_is_normalized=True)
    assert_topk_match(ids[sel], sc[sel], cnt[sel], ref, k, TOL, "background")


def test_overflow_beyond_capacity_falls_back_exactly(dev):
    """3000 near-identical rows per query exceed the 2048-candidate capacity: rvo_search_topk flags the query
    (count -1) and the public API re-runs it through the exact fp32 scan — still the right answer, never a wrong one."""
    from revers_o_b200 import ops, synth
    n, d, nq, k = 100_000, 512, 6, 50
    q = synth.make_queries(nq, d, seed=81, device=dev)
    db = _cluster_db(dev, n, d, q, 0.9500, 0.9503, 3000, seed=82)
    ids, sc, cnt = _run(db, n, d, q, k)
    assert np.all((cnt == k) | (cnt == -1))
    ids2, sc2, cnt2 = ops.search_topk_exact(db, n, d, q, k)
    torch.cuda.synchronize()
    assert_topk_match(ids2.cpu().numpy(), sc2.cpu().numpy(), cnt2.cpu().numpy(), _oracle(db, n, d, q, k), k, TOL, "fallback")


def test_config3_shard_size_against_fp32_reference(dev):
    """One configs[3] shard at full size (12.5M x 1280 bf16 = 32 GB, the per-GPU share of the 100M x 1280 DB at 8 GPUs):
    the oracle cannot hold it, so the check is a plain fp32 GPU reference of the same definition (normalise the query,
    fp32 dot with every stored row, descending top-k) computed chunk by chunk, on the bandwidth-regime batch (Q = 16),
    the Q = 1 path and a slice of the tensor-regime batch (Q = 4096).  Tolerances as everywhere: scores 1e-3, id sets
    equal except ties within 1e-3."""
    from revers_o_b200 import ops, synth
    if torch.cuda.get_device_properties(dev).total_memory < 100e9:
        pytest.skip("needs a full-size B200")
    n, d, k = 12_500_000, 1280, 100
    q_all = synth.make_queries(4096, d, seed=7, device=dev)
    db = synth.make_db(n, d, q_all, n_plant=16, seed=2000, device=dev)
    qn = q_all[:16] / q_all[:16].norm(dim=1, keepdim=True)
    best_s = torch.full((16, 0), 0.0, device=dev)
    best_i = torch.zeros((16, 0), dtype=torch.int64, device=dev)
    nb = db.shape[0]
    for b0 in range(0, nb, 2048):                                   # 262k rows per chunk
        b1 = min(nb, b0 + 2048)
        rows = db[b0:b1].permute(0, 2, 1, 3).reshape((b1 - b0) * 128, -1)[:, :d].float()
        s = qn @ rows.T
        valid = (torch.arange(b0 * 128, b1 * 128, device=dev) < n)
        s[:, ~valid] = -2.0
        cs, ci = torch.cat([best_s, s], 1), torch.cat([best_i, torch.arange(b0 * 128, b1 * 128, device=dev).expand(16, -1)], 1)
        top = cs.topk(k, dim=1)
        best_s, best_i = top.values, torch.gather(ci, 1, top.indices)
        del rows, s, cs, ci
    ref = [(best_i[i].cpu().numpy(), best_s[i].cpu().numpy()) for i in range(16)]
    ids, sc, cnt = _run(db, n, d, q_all[:16].contiguous(), k)
    assert_topk_match(ids, sc, cnt, ref, k, TOL, "cfg3-q16")
    ids1, sc1, cnt1 = _run(db, n, d, q_all[:1].contiguous(), k)
    assert_topk_match(ids1, sc1, cnt1, ref[:1], k, TOL, "cfg3-q1")
    idsL, scL, cntL = _run(db, n, d, q_all, k)
    assert np.all(cntL == k) and np.all(np.diff(scL, axis=1) <= 1e-7)
    assert_topk_match(idsL[:16], scL[:16], cntL[:16], ref, k, TOL, "cfg3-q4096")


# ---- degenerate inputs the reference can meet (zero vectors, tiny DBs, extreme k / thresholds) -----------------------
@pytest.mark.parametrize("nq", [2, 9])
def test_degenerate_queries_and_shapes(dev, nq):
    """Zero query (qdrant leaves a zero vector as is: every score 0), a DB smaller than one 128-row tile, k = 1, the
    maximum k, k > n, a threshold above every score (empty result -> the 'no similar regions' path, core_system.py:666)
    and below every score — on both the Q <= 4 and the tensor path."""
    from revers_o_b200 import ops, synth
    from revers_o_b200._lib import RVO_MAX_K
    n, d = 90, 128
    q = synth.make_queries(nq, d, seed=91, device=dev)
    db = synth.make_db(n, d, q, n_plant=4, seed=92, device=dev)
    big_n = 70_000
    db2 = synth.make_db(big_n, d, q, n_plant=8, seed=93, device=dev)     # the same batch on a multi-level shard
    q[1] = 0.0                                                            # zero query vector (after planting: no NaN rows)
    for k, thr in ((1, None), (10, None), (RVO_MAX_K, None), (10, 1.5), (10, -1.0), (200, 0.3)):
        ids, sc, cnt = _run(db, n, d, q, k, thr)
        ref = _oracle(db, n, d, q, k, thr)
        for i in (0, 1, nq - 1):
            c = int(cnt[i])
            assert c == len(ref[i][0]) or i == 1, (k, thr, i, c, len(ref[i][0]))
            if i == 1:                                                    # all scores exactly 0: any n ids are a valid answer
                want = 0 if (thr is not None and thr > 0) else min(k, n)
                assert c == want and np.all(sc[i, :c] == 0.0) and len(set(ids[i, :c].tolist())) == c
            elif c:
                assert np.max(np.abs(sc[i, :c] - ref[i][1])) < TOL and set(ids[i, :c].tolist()) == set(ref[i][0].tolist())
            assert np.all(ids[i, c:] == -1)
    # 70k rows all scoring exactly 0 against the zero query is a tie mass beyond the fused path's capacity: it is flagged (-1)
    # on the tensor path and the documented protocol (search_topk_exact) answers it; the other queries are unaffected
    raw = _run(db2, big_n, d, q, 50)[2]
    assert raw[1] in (-1, 50) and np.all(np.delete(raw, 1) == 50)
    a, b, c = ops.search_topk_exact(db2, big_n, d, q, 50)
    torch.cuda.synchronize()
    ids, sc, cnt = a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    ref = _oracle(db2, big_n, d, q, 50)
    assert cnt[1] == 50 and np.all(sc[1, :50] == 0.0)
    keep = [i for i in range(nq) if i != 1]
    assert_topk_match(ids[keep], sc[keep], cnt[keep], [ref[i] for i in keep], 50, TOL, "degenerate")


def test_non_finite_vectors_are_stored_as_zero_vectors(dev):
    """A NaN / Inf component would turn every score against that row into NaN and float it to the top of every search (the
    reference's numpy argsort does the same).  The ingest kernel stores such a row as the zero vector instead, and a
    non-finite query scores 0 against everything."""
    from revers_o_b200 import ops
    rs = np.random.RandomState(5)
    n, d, k = 5000, 128, 5
    x = rs.randn(n, d).astype(np.float32)
    x[7, 3] = np.nan
    x[9, 0] = np.inf
    db, _ = ops.normalize_rows(torch.from_numpy(x).to(dev), db=ops.db_alloc(n, d, dev))
    rows = ops.untile_rows(db, n, d).float()
    assert torch.isfinite(rows).all() and float(rows[7].abs().sum()) == 0.0 and float(rows[9].abs().sum()) == 0.0
    q = rs.randn(6, d).astype(np.float32)
    q[2, 5] = np.nan
    # the sanitised query ties with all 5000 rows at score 0: beyond the fused path's capacity, answered by the exact protocol
    a, b, c = ops.search_topk_exact(db, n, d, torch.from_numpy(q).to(dev), k)
    torch.cuda.synchronize()
    ids, sc, cnt = a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    assert np.isfinite(sc).all() and np.all(cnt == k) and np.all(sc[2] == 0.0)
    good = [0, 1, 3, 4, 5]
    xf = x.copy(); xf[7] = 0; xf[9] = 0
    ref = O.search_batch(O.round_to_bf16(O._cosine_prepare(xf)), q[good], k, None, db_is_normalized=True)
    assert_topk_match(ids[good], sc[good], cnt[good], ref, k, TOL, "nonfinite")


# ---- the reference's operating point (Q <= 4) on shards beyond the keep-every-score regime ----------------------------
@pytest.mark.parametrize("n,d,nq,k,thr", [
    (400_000, 1024, 1, 10, None),       # sampled: [sample scan] -> [k-th best key] -> [filtered full scan] -> [exact top-k]
    (400_000, 1024, 4, 100, 0.3),
    (250_001, 1280, 3, 512, None),      # odd row count, maximum k, PE-Core-G14 width
    (131_072, 256, 2, 10, None),        # last size of the dense regime
    (131_073, 256, 2, 10, None),        # first size of the sampled regime
    (150_000, 2048, 2, 20, None),       # long rows (register-heavy instantiation)
])
def test_small_q_sampled_regime_matches_oracle(dev, n, d, nq, k, thr):
    from revers_o_b200 import synth
    q = synth.make_queries(nq, d, seed=21, device=dev)
    db = synth.make_db(n, d, q, n_plant=min(128, 2 * k), seed=22, device=dev)
    ids, sc, cnt = _run(db, n, d, q, k, thr)
    assert_topk_match(ids, sc, cnt, _oracle(db, n, d, q, k, thr), k, TOL, f"small n{n}d{d}q{nq}k{k}")


def test_small_q_paths_agree_bit_for_bit_and_dense_route_never_overflows(dev):
    """AUTO (sampled), SMALL and DENSE compute a row's score with the same arithmetic: identical ids AND scores.  A shard whose
    rows are all the same vector (every score ties, 300k survivors of any threshold) overflows the survivor lists of the
    sampled route (count -1) — the dense route and `search_topk_exact` still return the exact answer (lowest rows first)."""
    from revers_o_b200 import _lib, ops, synth
    n, d, nq, k = 300_000, 512, 3, 50
    q = synth.make_queries(nq, d, seed=31, device=dev)
    db = synth.make_db(n, d, q, n_plant=64, seed=32, device=dev)
    a = ops.search_topk(db, n, d, q, k)
    b = ops.search_topk(db, n, d, q, k, path=_lib.RVO_PATH_DENSE)
    c = ops.search_topk(db, n, d, q, k, path=_lib.RVO_PATH_SMALL)
    torch.cuda.synchronize()
    for x, y in ((a, b), (a, c)):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) and torch.equal(x[2], y[2])
    # tensor path on the same 3 queries: same ids, scores within tolerance
    t = ops.search_topk(db, n, d, q, k, path=_lib.RVO_PATH_TENSOR)
    assert torch.equal(t[0], a[0]) and float((t[1] - a[1]).abs().max()) < 1e-5
    # all-equal rows: rows 0..k-1 win every tie
    one = torch.randn((1, d), generator=torch.Generator().manual_seed(1)).to(dev)
    same = ops.db_alloc(n, d, dev)
    ops.normalize_rows(one.expand(4096, d).contiguous(), db=same, row0=0)
    blk = same[: 4096 // 128].clone()
    for b0 in range(0, same.shape[0], blk.shape[0]):
        m = min(blk.shape[0], same.shape[0] - b0)
        same[b0: b0 + m].copy_(blk[:m])
    qq = one.contiguous()
    ids, sc, cnt = ops.search_topk(same, n, d, qq, k)
    assert int(cnt[0]) in (-1, k)
    ids, sc, cnt = ops.search_topk(same, n, d, qq, k, path=_lib.RVO_PATH_DENSE)
    assert int(cnt[0]) == k and ids[0].tolist() == list(range(k)) and float(sc[0].min()) > 0.99
    ids, sc, cnt = ops.search_topk_exact(same, n, d, qq, k)
    assert int(cnt[0]) == k and ids[0].tolist() == list(range(k))


def test_submit_collect_pipeline_on_one_gpu(dev):
    """ShardedIndex.submit / collect (two batches in flight on two streams) == the synchronous search, batch by batch."""
    from revers_o_b200 import synth
    from revers_o_b200.sharded import ShardedIndex
    n, d, k = 90_000, 512, 30
    qs = [synth.make_queries(40, d, seed=60 + i, device=dev) for i in range(5)]
    db = synth.make_db(n, d, qs[0], n_plant=64, seed=61, device=dev)
    idx = ShardedIndex(db, n, d, 1000)
    want = [tuple(t.clone() for t in idx.search(q, k, 0.1)) for q in qs]
    got, prev = [], None
    for q in qs:
        t = idx.submit(q, k, 0.1)
        if prev is not None:
            got.append(tuple(x.clone() for x in idx.collect(prev)))
        prev = t
    got.append(tuple(x.clone() for x in idx.collect(prev)))
    torch.cuda.synchronize()
    for w, g in zip(want, got):
        assert torch.equal(w[0], g[0]) and torch.equal(w[1], g[1]) and torch.equal(w[2], g[2])
    a, b = idx.submit(qs[0], k), idx.submit(qs[1], k)
    with pytest.raises(Exception):
        idx.submit(qs[2], k)
    with pytest.raises(Exception):
        idx.collect(b)                      # tickets come back in submission order
    idx.collect(a), idx.collect(b)
