"""CPU: the oracle against hand-computable known answers and against the committed
golden fixtures (regression pin of the restatement)."""
import numpy as np

from oracle import reverso_oracle as O


def test_ka1_identity_db():
    db = np.eye(16, dtype=np.float32)
    for j in (0, 5, 15):
        ids, sc = O.search(db, db[j], 3, None)
        assert ids[0] == j and sc[0] == 1.0 and np.all(sc[1:] == 0.0)


def test_ka2_duplicate_rows_both_returned():
    db = np.eye(8, dtype=np.float32)
    db[6] = db[2]
    ids, sc = O.search(db, db[2], 2, None)
    assert set(ids.tolist()) == {2, 6} and np.all(sc == 1.0)


def test_ka3_threshold_equal_is_kept():
    db = np.array([[1, 0], [0.6, 0.8], [0, 1]], dtype=np.float32)
    ids, sc = O.search(db, np.array([1, 0], np.float32), 10, 0.6)
    thr = np.dot(O._cosine_prepare(db)[1], np.array([1, 0], np.float32))
    ids2, _ = O.search(db, np.array([1, 0], np.float32), 10, float(thr))
    assert ids2.tolist() == [0, 1]  # score == threshold is kept (walk stops at the first score < threshold)
    ids3, _ = O.search(db, np.array([1, 0], np.float32), 10, float(np.nextafter(thr, np.float32(2))))
    assert ids3.tolist() == [0]


def test_ka4_limit_above_n():
    db = np.random.RandomState(0).randn(5, 8).astype(np.float32)
    ids, sc = O.search(db, db[0], 50, None)
    assert len(ids) == 5 and np.all(np.diff(sc) <= 0)


def test_ka5_no_hit_above_threshold():
    db = np.eye(4, dtype=np.float32)
    ids, sc = O.search(db, -db[0], 5, 0.5)
    assert len(ids) == 0


def test_ka6_unnormalised_inputs_same_ranking():
    rs = np.random.RandomState(1)
    db, q = rs.randn(200, 32).astype(np.float32), rs.randn(32).astype(np.float32)
    a, sa = O.search(db, q, 10, None)
    b, sb = O.search(db * rs.uniform(0.1, 9, (200, 1)).astype(np.float32), 7.5 * q, 10, None)
    assert a.tolist() == b.tolist() and np.allclose(sa, sb, atol=1e-6)


def test_ka7_all_ones_mask_is_token_mean():
    rs = np.random.RandomState(2)
    f = rs.randn(1, 12, 16).astype(np.float32)
    emb, counts, _ = O.mask_pool(f, np.ones((1, 1, 12), np.uint8))
    ref = O.l2_normalize(O.global_embedding(f)[0])  # core_system.py:345-346 + :407
    assert counts.tolist() == [1] and np.allclose(emb[0], ref, atol=1e-6)


def test_ka8_single_patch_mask():
    rs = np.random.RandomState(3)
    f = rs.randn(1, 9, 8).astype(np.float32)
    m = np.zeros((1, 1, 9), np.uint8)
    m[0, 0, 4] = 1
    emb, _, _ = O.mask_pool(f, m)
    assert np.allclose(emb[0], f[0, 4] / np.linalg.norm(f[0, 4]), atol=1e-6)


def test_ka9_empty_mask_dropped_and_shifted():
    rs = np.random.RandomState(4)
    f = rs.randn(1, 9, 8).astype(np.float32)
    m = np.zeros((1, 3, 9), np.uint8)
    m[0, 0, :3] = 1
    m[0, 2, 5:] = 1
    emb, counts, src = O.mask_pool(f, m)
    assert counts.tolist() == [2] and src.tolist() == [[0, 0], [0, 2]] and emb.shape == (2, 8)


def test_ka10_float_mask_threshold():
    m = np.array([[0.5, 0.51], [0.2, 0.9]], np.float32)
    assert O.binarize_mask(m).tolist() == [[0, 1], [0, 1]]  # strictly greater than 0.5 (core_system.py:399)
    assert O.binarize_mask(np.array([[True, False]])).tolist() == [[1, 0]]


def test_reference_mode_every_region_gets_global_embedding():
    rs = np.random.RandomState(5)
    tokens = rs.randn(1, 10, 8).astype(np.float32)
    masks = [np.ones((4, 4), bool), np.zeros((4, 4), bool), np.ones((4, 4), np.float32)]
    embs = O.extract_embeddings_reference(tokens, masks)
    g = O.l2_normalize(tokens.mean(axis=1)[0])
    assert len(embs) == 2 and all(np.allclose(e, g) for e in embs)


def test_region_cap_50():
    f = np.random.RandomState(6).randn(1, 4, 8).astype(np.float32)
    emb, counts, _ = O.mask_pool(f, np.ones((1, 64, 4), np.uint8), max_regions=O.MAX_REGIONS)
    assert counts.tolist() == [50]


def test_qdrant_duck_type_roundtrip():
    from types import SimpleNamespace as NS
    c = O.QdrantLocalOracle(path="x")
    c.recreate_collection("c", vectors_config=NS(size=4, distance="Cosine"))
    c.upsert("c", [NS(id="a", vector=[2, 0, 0, 0], payload={"f": 1}), NS(id="b", vector=[0, 3, 0, 0], payload={"f": 2})])
    hits = c.search("c", [1, 0.1, 0, 0], limit=5, score_threshold=0.5)
    assert [h.payload["f"] for h in hits] == [1] and abs(hits[0].score - 1 / np.sqrt(1.01)) < 1e-6
    assert [x.name for x in c.get_collections().collections] == ["c"]


def test_merge_topk_tie_by_lower_id():
    ids = np.array([[[5, 9]], [[3, 7]]], np.int64)
    sc = np.array([[[0.9, 0.5]], [[0.9, 0.8]]], np.float32)
    oi, os_, oc = O.merge_topk(ids, sc, np.full((2, 1), 2, np.int32), 3)
    assert oi.tolist() == [[3, 5, 7]] and oc.tolist() == [3]


def test_bf16_rounding_helpers():
    x = np.array([1.0, 1.00390625, 1.001953125, -3.14159, 0.0], np.float32)
    r = O.round_to_bf16(x)
    assert r[0] == 1.0 and r[1] == 1.0 and np.allclose(O.bf16_bits_to_f32(O.bf16_bits(x)), r)


def test_mask_to_patch_grid_rules():
    m = np.zeros((48, 48), bool)
    m[:24, :24] = True
    g = O.mask_to_patch_grid(m, 4)
    assert g.reshape(4, 4)[:2, :2].all() and g.sum() == 4
    tiny = np.zeros((48, 48), bool)
    tiny[5, 5] = True
    assert O.mask_to_patch_grid(tiny, 4).sum() == 1  # never silently empty


def test_golden_search(golden):
    db = O.bf16_bits_to_f32(golden["search_db_bf16_bits"])
    q = golden["search_queries"]
    for tag, k, thr in (("k10", 10, None), ("k10_t07", 10, 0.7), ("k100", 100, None), ("kall", 3000, None)):
        res = O.search_batch(db, q, k, thr, db_is_normalized=True)
        for i, (ids, sc) in enumerate(res):
            n = golden[f"search_{tag}_counts"][i]
            assert len(ids) == n
            assert np.array_equal(golden[f"search_{tag}_scores"][i, :n], sc)
            assert set(golden[f"search_{tag}_ids"][i, :n].tolist()) == set(ids.tolist())
    assert golden["search_kall_counts"].tolist() == [2000] * 9


def test_golden_pool_and_merge(golden):
    feats = O.bf16_bits_to_f32(golden["pool_feats_bf16_bits"])
    emb, counts, src = O.mask_pool(feats, golden["pool_masks"])
    assert np.array_equal(counts, golden["pool_counts"]) and np.allclose(emb, golden["pool_emb"], atol=1e-6)
    assert counts.tolist() == [5, 5, 5]
    emb4, counts4, _ = O.mask_pool(feats, golden["pool_masks"], max_regions=4)
    assert np.array_equal(counts4, golden["pool_counts_cap4"]) and np.allclose(emb4, golden["pool_emb_cap4"], atol=1e-6)
    mi, ms, mc = O.merge_topk(golden["merge_ids"], golden["merge_scores"], golden["merge_counts"], 5)
    assert np.array_equal(mi, golden["merge_out_ids"]) and np.array_equal(mc, golden["merge_out_counts"])


def test_fair_batched_variant_agrees():
    rs = np.random.RandomState(8)
    db = O._cosine_prepare(rs.randn(500, 32).astype(np.float32))
    q = rs.randn(6, 32).astype(np.float32)
    ids, sc = O.search_batch_fair(db, q, 7)
    for i, (a, b) in enumerate(O.search_batch(db, q, 7, None, db_is_normalized=True)):
        assert a.tolist() == ids[i].tolist() and np.allclose(b, sc[i], atol=1e-6)
