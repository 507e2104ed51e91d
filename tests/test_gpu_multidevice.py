"""GPU: the row-sharded search behind the reference's own entry points, in ONE process: a collection is
saved, loaded onto several shards through `SimpleReverso.load_database`, and `search_similar` / `B200VectorDB.search` return the
same hits as the single-device path.  On a 1-GPU box the shards are virtual (the same device listed twice or three times);
with two or more GPUs the real devices are used as well.  Also: implicit persistence (qdrant-local persists on upsert,
core_system.py:521,621), crash consistency of the write-through, search vs in-place upsert, NaN ingest."""
import json
import os
import threading

import numpy as np
import pytest
import torch

from oracle import reverso_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def device_sets():
    sets = [[0, 0], [0, 0, 0]]
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        sets.append(list(range(min(torch.cuda.device_count(), 8))))
    return sets


def build(path, n, d, seed=0, **kw):
    from revers_o_b200.vector_db import B200VectorDB, models
    rs = np.random.RandomState(seed)
    v = rs.randn(n, d).astype(np.float32)
    v[n // 2] = v[n // 2 - 1]                       # an exact duplicate: ties must resolve the same way on every layout
    db = B200VectorDB(path=path, **kw)
    db.recreate_collection("simple_reverso_t", vectors_config=models.VectorParams(size=d, distance=models.Distance.COSINE))
    for lo in range(0, n, 1000):
        db.upsert_batch("simple_reverso_t", [f"id-{i:07d}" for i in range(lo, min(n, lo + 1000))], v[lo: lo + 1000],
                        [{"filename": f"f{i}.jpg", "bbox": [i, 0, 1, 1], "image_source": ""} for i in range(lo, min(n, lo + 1000))])
    return db, v


@pytest.mark.parametrize("devices", device_sets())
def test_sharded_collection_equals_single_device(dev, tmp_path, devices):
    from revers_o_b200.core_system import SimpleReverso
    from revers_o_b200.vector_db import B200VectorDB
    n, d = 7000, 256
    root = tmp_path / "simple_reverso_db"
    single, v = build(str(root / "t"), n, d, device=dev)              # written through while ingesting: nothing to save()
    rs = np.random.RandomState(5)
    q = np.concatenate([v[[n // 2, 17, 4242]] + 0.05 * rs.randn(3, d).astype(np.float32), rs.randn(30, d).astype(np.float32)])
    # the reference's entry points: load_database -> search_similar (core_system.py:90-119, 650-676)
    r1 = SimpleReverso(db_root=str(root), device=dev)
    rn = SimpleReverso(db_root=str(root), devices=devices)
    assert r1.load_database("t").startswith("✅") and rn.load_database("t").startswith("✅")
    cn = rn.vector_db._coll(rn.current_database)
    assert [s.n for s in cn.shards if s.n] and sum(s.n for s in cn.shards) == n and len([s for s in cn.shards if s.n]) == len(devices)
    assert all(s.row0 % 128 == 0 for s in cn.shards)
    for j in (0, 1, 5):                                               # Q = 1, what search_similar issues (core_system.py:657)
        r1.region_embeddings = rn.region_embeddings = [torch.from_numpy(q[j])]
        t1, i1 = r1.search_similar(0.2, 10)
        tn, in_ = rn.search_similar(0.2, 10)
        assert t1 == tn and [x["filename"] for x in i1] == [x["filename"] for x in in_] and len(i1) >= 1
    # batched: ids, scores, counts identical to the single-device path (same kernels, same tie rule: lower global row first)
    for k, thr in ((10, None), (100, 0.1)):
        a = r1.vector_db.search_batch(r1.current_database, q, k, thr)
        b = rn.vector_db.search_batch(rn.current_database, q, k, thr)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.allclose(a[1], b[1], atol=1e-6)
    # ... and to the oracle (bf16-valued DB)
    from revers_o_b200 import ops
    c1 = r1.vector_db._coll(r1.current_database)
    ref = O.search_batch(ops.untile_rows(c1.vectors, n, d).float().cpu().numpy(), q, 10, None, db_is_normalized=True)
    b = rn.vector_db.search_batch(rn.current_database, q, 10)
    for i, (rid, rsc) in enumerate(ref):
        assert np.allclose(b[1][i], rsc, atol=1e-3) and set(b[0][i].tolist()) == set(rid.tolist())
    # hits carry the uuid/payload of the GLOBAL row
    hits = rn.vector_db.search(rn.current_database, v[4242].tolist(), limit=1)
    assert hits[0].id == "id-0004242" and hits[0].payload["filename"] == "f4242.jpg" and hits[0].score > 0.999


def test_incremental_ingest_spills_to_the_next_shard_and_rebalances(dev, tmp_path):
    from revers_o_b200.vector_db import B200VectorDB
    n, d = 3000, 128
    db, v = build(None, n, d, devices=[0, 0, 0], shard_rows=1024)
    c = db._coll("simple_reverso_t")
    assert [s.n for s in c.shards] == [1024, 1024, 952] and [s.row0 for s in c.shards] == [0, 1024, 2048]
    one, _ = build(None, n, d, device=dev)
    q = v[[5, 1500, 2999]]
    a, b = one.search_batch("simple_reverso_t", q, 7), db.search_batch("simple_reverso_t", q, 7)
    assert np.array_equal(a[0], b[0]) and np.allclose(a[1], b[1], atol=1e-6)
    # an upsert of an existing id overwrites its row on whatever shard holds it
    from revers_o_b200.vector_db import models
    db.upsert("simple_reverso_t", [models.PointStruct(id="id-0001500", vector=v[7].tolist(), payload={"filename": "new"})])
    h = db.search("simple_reverso_t", v[7].tolist(), limit=2, score_threshold=0.999)
    assert sorted(x.id for x in h) == ["id-0000007", "id-0001500"] and db.count("simple_reverso_t") == n
    db.rebalance("simple_reverso_t")
    assert [s.n for s in c.shards] == [1024, 1024, 952] or sum(s.n for s in c.shards) == n
    h2 = db.search("simple_reverso_t", v[7].tolist(), limit=2, score_threshold=0.999)
    assert sorted(x.id for x in h2) == ["id-0000007", "id-0001500"]


def test_implicit_persistence_and_crash_consistency(dev, tmp_path):
    """qdrant-local persists on upsert (core_system.py:521,621: no save call anywhere in the reference): a second client on the
    same path sees every batch; bytes written after the last meta.json are ignored and cut off."""
    from revers_o_b200.vector_db import B200VectorDB, models
    p = str(tmp_path / "db")
    db, v = build(p, 2500, 64, device=dev)
    files = set(os.listdir(p))
    assert {"meta.json", "simple_reverso_t.bf16", "simple_reverso_t.ids", "simple_reverso_t.payload.jsonl",
            "simple_reverso_t.payload.idx"} <= files
    meta = json.load(open(os.path.join(p, "meta.json")))["collections"]["simple_reverso_t"]
    assert meta["n"] == 2500 and meta["ids_bytes"] == 2500 * 36 and meta["ids_dtype"].endswith("S36")
    again = B200VectorDB(path=p, device=dev)
    assert again.count("simple_reverso_t") == 2500
    h = again.search("simple_reverso_t", v[2499].tolist(), limit=1)
    assert h[0].id == "id-0002499" and h[0].payload == {"filename": "f2499.jpg", "bbox": [2499, 0, 1, 1], "image_source": ""}
    assert not again._coll("simple_reverso_t").payloads._ram         # payloads are parsed lazily from the log
    # overwrite + append through the second client, then a torn tail (crash after the data, before meta.json)
    again.upsert("simple_reverso_t", [models.PointStruct(id="id-0000003", vector=v[9].tolist(), payload={"filename": "x"}),
                                      models.PointStruct(id="brand-new-id-longer-than-the-36-characters-of-a-uuid", vector=v[1].tolist(), payload=None)])
    for f in ("simple_reverso_t.payload.jsonl", "simple_reverso_t.ids", "simple_reverso_t.bf16"):
        with open(os.path.join(p, f), "ab") as fh:
            fh.write(b"\x01garbage after the last meta.json")
    third = B200VectorDB(path=p, device=dev)
    assert third.count("simple_reverso_t") == 2501
    h = third.search("simple_reverso_t", v[9].tolist(), limit=2, score_threshold=0.999)
    assert sorted((x.id, (x.payload or {}).get("filename")) for x in h) == [("id-0000003", "x"), ("id-0000009", "f9.jpg")]
    assert third.search("simple_reverso_t", v[1].tolist(), limit=2)[1].id in ("brand-new-id-longer-than-the-36-characters-of-a-uuid", "id-0000001")
    third.upsert("simple_reverso_t", [models.PointStruct(id="after-crash", vector=v[2].tolist(), payload={"ok": 1})])
    fourth = B200VectorDB(path=p, device=dev)
    assert fourth.count("simple_reverso_t") == 2502 and fourth.search("simple_reverso_t", v[2].tolist(), limit=2)[0].score > 0.999
    assert {x.id for x in fourth.search("simple_reverso_t", v[2].tolist(), limit=2)} == {"after-crash", "id-0000002"}
    # recreate_collection starts fresh on disk as well (core_system.py:600)
    fourth.recreate_collection("simple_reverso_t", vectors_config=models.VectorParams(size=64, distance=models.Distance.COSINE))
    assert B200VectorDB(path=p, device=dev).count("simple_reverso_t") == 0


def test_failed_write_rolls_the_host_tables_back(dev):
    from revers_o_b200.vector_db import B200VectorDB, models
    db, v = build(None, 300, 64, device=dev)
    with pytest.raises(Exception):
        db.upsert("simple_reverso_t", [models.PointStruct(id="a", vector=v[0].tolist(), payload=None),
                                       models.PointStruct(id="b", vector=[1.0, 2.0], payload=None)])
    c = db._coll("simple_reverso_t")
    assert db.count("simple_reverso_t") == 300 and len(c.ids) == 300 and len(c.payloads) == 300
    db.upsert("simple_reverso_t", [models.PointStruct(id="a", vector=v[0].tolist(), payload={"k": 1})])
    assert db.count("simple_reverso_t") == 301 and c.ids[300] == "a"


def test_search_never_sees_a_half_written_upsert(dev):
    """Concurrent Gradio callbacks (ui.py:20 shares one instance): searches overlap each other, an in-place upsert is exclusive."""
    from revers_o_b200.vector_db import B200VectorDB, models
    db, v = build(None, 2000, 128, device=dev)
    stop, errors = threading.Event(), []
    a, b = v[10].copy(), -v[10]                      # row "flip" alternates between two normalised vectors

    def writer():
        i = 0
        while not stop.is_set():
            db.upsert("simple_reverso_t", [models.PointStruct(id="flip", vector=(a if i % 2 else b).tolist(), payload={"i": i})])
            i += 1

    def reader():
        try:
            while not stop.is_set():
                h = db.search("simple_reverso_t", a.tolist(), limit=3)
                s = [x.score for x in h if x.id == "flip"]
                # either the row scores ~1 (it holds a) or it is absent (it holds -a): never something in between
                assert all(abs(x - 1.0) < 1e-2 for x in s), s
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    db.upsert("simple_reverso_t", [models.PointStruct(id="flip", vector=a.tolist(), payload=None)])
    ts = [threading.Thread(target=writer)] + [threading.Thread(target=reader) for _ in range(3)]
    [t.start() for t in ts]
    import time
    time.sleep(2.0)
    stop.set()
    [t.join() for t in ts]
    assert not errors, errors[0]


def test_non_finite_and_zero_rows_never_poison_a_search(dev):
    """the fused ingest (rvo_mask_pool_to_db) and the CUDA-core pooling apply the same rule as rvo_normalize_rows —
    a region over all-zero features (0 * inf) or NaN features is stored as the zero vector, so it scores 0, not NaN."""
    from revers_o_b200 import _lib, ops
    from revers_o_b200.vector_db import B200VectorDB, models
    B, M, P, D = 3, 4, 64, 256
    g = torch.Generator().manual_seed(0)
    feats = torch.randn((B, P, D), generator=g).to(dev).to(torch.bfloat16)
    feats[1] = 0                                      # image 1: every region's mean is the zero vector
    feats[2, 5, 7] = float("nan")                     # image 2: one NaN feature under region 0's mask
    masks = torch.zeros((B, M, P), dtype=torch.uint8, device=dev)
    masks[:, 0, :32] = 1
    masks[:, 1, 32:] = 1
    for path in (0, 1):                               # tensor-core kernel and the CUDA-core kernels
        _lib.set_option("pool_path", path)
        try:
            out, counts, src, total = ops.mask_pool(feats, masks)
            t = int(total.item())
            o = out[:t].cpu().numpy()
            assert t == 6 and np.isfinite(o).all()
            assert np.allclose(np.linalg.norm(o[[0, 1]], axis=1), 1, atol=1e-4) and np.abs(o[[2, 3, 4]]).max() == 0
            # region 1 of image 2 does not cover the NaN patch: the CUDA-core kernels (which only visit covered patches) keep it;
            # on tensor cores the pooling is a GEMM and 0 * NaN = NaN, so the whole image's regions become zero vectors
            nrm5 = float(np.linalg.norm(o[5]))
            assert abs(nrm5 - 1) < 1e-4 if path == 1 else (nrm5 == 0.0 or abs(nrm5 - 1) < 1e-4)
        finally:
            _lib.set_option("pool_path", 0)
    db = B200VectorDB(device=dev)
    db.recreate_collection("r", vectors_config=models.VectorParams(size=D, distance=models.Distance.COSINE))
    assert db.ingest_regions("r", feats, masks) == 6
    q = feats[0, :32].float().mean(0).cpu().numpy()
    ids, sc, cnt = db.search_batch("r", q[None], 6)
    assert np.isfinite(sc[0, : cnt[0]]).all() and ids[0, 0] == 0 and sc[0, 0] > 0.99 and np.sum(np.abs(sc[0, : cnt[0]]) < 1e-6) >= 3
    assert cnt[0] == 6
    with pytest.raises(Exception):
        db.search("r", [float("nan")] * D)
