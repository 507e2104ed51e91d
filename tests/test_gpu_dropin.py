"""GPU: the drop-in `SimpleReverso` + `B200VectorDB` driven with the call pattern of the unmodified ui.py
(ui.py:29-159), a fake encoder/detector standing in for the third-party models, checked against the oracle's
duck-typed qdrant-local restatement fed the same points."""
import os
from types import SimpleNamespace as NS

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


class FakeEncoder:
    """Deterministic stand-in for pe.CLIP.encode_image: [1,3,H,W] -> [1, 1+24*24, 1024] tokens."""
    def __init__(self, dev, tokens=True):
        g = torch.Generator().manual_seed(0)
        self.w = torch.randn((3, 1024), generator=g).to(dev)
        self.tokens = tokens

    def encode_image(self, x):
        p = torch.nn.functional.adaptive_avg_pool2d(x.float(), 24).flatten(2).transpose(1, 2)   # [1,576,3]
        t = torch.tanh(p @ self.w + torch.linspace(-1, 1, 1024, device=x.device))
        if not self.tokens:
            return t.mean(1)
        return torch.cat([t.mean(1, keepdim=True), t], 1)


def preprocess(pil):
    a = np.asarray(pil.resize((336, 336)), dtype=np.float32) / 255.0
    return torch.from_numpy(a).permute(2, 0, 1)


def detector(pil, prompt):
    H, W = pil.height, pil.width
    m = np.zeros((4, H, W), bool)
    m[0, : H // 2, : W // 2] = True
    m[1, H // 4:, W // 3:] = True
    # m[2] stays empty -> dropped (core_system.py:402-404)
    m[3, H // 2:, :] = True
    return NS(mask=m, confidence=np.array([0.9, 0.8, 0.7, 0.6]), class_id=np.array([0, 0, 0, 0]),
              xyxy=np.array([[0, 0, W // 2, H // 2]] * 4), __len__=lambda: 4)


class Det(NS):
    def __len__(self):
        return len(self.mask)


def make_image(seed, size=(96, 80)):
    from PIL import Image
    rs = np.random.RandomState(seed)
    coarse = Image.fromarray((rs.rand(5, 6, 3) * 255).astype(np.uint8))       # low-frequency content: images differ globally
    img = np.asarray(coarse.resize(size, Image.BILINEAR), dtype=np.float32)
    img += rs.rand(size[1], size[0], 3) * 20
    return Image.fromarray(np.clip(img, 0, 255).astype(np.uint8))


@pytest.mark.parametrize("mode", ["reference", "pooled"])
def test_ui_call_pattern(dev, tmp_path, mode):
    from revers_o_b200.core_system import SimpleReverso
    det = lambda pil, prompt: Det(**vars(detector(pil, prompt)))
    r = SimpleReverso(encoder=FakeEncoder(dev), preprocess=preprocess, detector=det, parity_mode=mode,
                      db_root=str(tmp_path / "simple_reverso_db"), device=dev)
    folder = tmp_path / "imgs"
    folder.mkdir()
    for i in range(6):
        make_image(i).save(folder / f"img{i}.png")
    # ui.py:86-94 build_database_ui
    msgs = []
    status = r.create_database(folder_path=str(folder), database_name="t", text_prompt="object .", use_direct_pe=False,
                               resume_from_checkpoint=True, include_subfolders=False,
                               progress_callback=lambda m, p=None: msgs.append(m))            # keywords, as ui.py:86-94
    assert "ready for searching" in status and r.vector_db and r.current_database == "simple_reverso_t"
    assert any(m.startswith("🔄 Processing 1/6: img0.png") for m in msgs) and any("Stored batch 1/1 (18 points)" in m for m in msgs)
    assert "✅ Found 4 regions, extracted 3 embeddings in img0.png" in msgs and "✅ Successfully processed: 6 images" in msgs
    assert r.vector_db.count(r.current_database) == 6 * 3                 # one region of four is empty
    assert r.list_databases() == ["t"]
    # ui.py:47-53 detect_and_extract_ui on a query image that is in the DB
    q_img = make_image(2)
    assert r.detect_regions(q_img, "object .") == 4
    embs, metas = r.extract_embeddings(q_img)
    assert len(embs) == 3 and [m["detection_index"] for m in metas] == [0, 1, 3]
    assert all(abs(float(e.norm()) - 1) < 1e-5 for e in embs) and embs[0].device.type == "cpu"
    if mode == "reference":
        assert torch.equal(embs[0], embs[1])                               # global embedding for every region
    else:
        assert not torch.allclose(embs[0], embs[1])
    # ui.py:131 search_similar
    text, items = r.search_similar(0.5, 5)
    assert items and items[0]["filename"] == "img2.png" and items[0]["score"] > 0.999
    assert all(a["score"] >= b["score"] for a, b in zip(items, items[1:]))
    # ui.py:125-133 region selection by swapping region_embeddings
    keep = r.region_embeddings
    r.region_embeddings = [keep[2]]
    text2, items2 = r.search_similar(0.5, 5)
    r.region_embeddings = keep
    assert items2[0]["score"] > 0.999
    # same points through the oracle's qdrant-local restatement -> same hits
    oc = O.QdrantLocalOracle()
    c = r.vector_db._coll(r.current_database)
    oc.recreate_collection("x", size=c.dim)
    from revers_o_b200 import ops
    vec = ops.untile_rows(c.vectors, c.n, c.dim).float().cpu().numpy()
    oc.upsert("x", [NS(id=i, vector=v, payload=p) for i, v, p in zip(c.ids, vec, c.payloads)])
    ref = oc.search("x", keep[0].numpy(), limit=5, score_threshold=0.5)
    got = r.vector_db.search(r.current_database, keep[0].numpy(), limit=5, score_threshold=0.5)
    assert len(ref) == len(got) and all(abs(a.score - b.score) < 1e-3 for a, b in zip(ref, got))
    # status-string convention (core_system.py:652-666)
    assert r.search_similar(1.01, 5)[0].startswith("❌ No similar regions")
    # restart: load_database from disk (ui.py:207-214)
    r2 = SimpleReverso(db_root=str(tmp_path / "simple_reverso_db"), device=dev)
    assert r2.load_database("t").startswith("✅") and r2.load_database("nope").startswith("❌")
    r2.region_embeddings = [keep[0]]
    assert r2.search_similar(0.5, 5)[1][0]["filename"] == "img2.png"
    assert r2.search_similar_batch(0.5, 3)[1][0][0]["filename"] == "img2.png"
    assert r.delete_database("t").startswith("✅") and r.list_databases() == []


def test_direct_pe_and_pooled_output(dev, tmp_path):
    from revers_o_b200.core_system import SimpleReverso
    r = SimpleReverso(encoder=FakeEncoder(dev, tokens=False), preprocess=preprocess, device=dev,
                      db_root=str(tmp_path / "db"))
    embs, metas = r.process_image_direct_pe(make_image(1))
    assert len(embs) == 1 and metas[0]["detected_class"] == "full_image" and abs(float(embs[0].norm()) - 1) < 1e-6
    feats = FakeEncoder(dev, tokens=False).encode_image(preprocess(make_image(1)).unsqueeze(0).to(dev))
    ref = O.l2_normalize(feats[0].cpu().numpy())
    assert np.allclose(embs[0].numpy(), ref, atol=1e-6)
    assert r.search_similar()[0].startswith("❌ No database loaded")


def test_vector_db_upsert_overwrite_and_batch(dev):
    from revers_o_b200.vector_db import B200VectorDB, models
    db = B200VectorDB(device=dev)
    db.recreate_collection("c", vectors_config=models.VectorParams(size=100, distance=models.Distance.COSINE))
    rs = np.random.RandomState(0)
    v = rs.randn(250, 100).astype(np.float32)
    for j in range(0, 250, 100):                                           # batches of 100, core_system.py:612
        db.upsert("c", [models.PointStruct(id=f"p{i}", vector=v[i].tolist(), payload={"i": i}) for i in range(j, min(j + 100, 250))])
    db.upsert("c", [models.PointStruct(id="p7", vector=v[3].tolist(), payload={"i": 7, "new": True})])   # overwrite
    assert db.count("c") == 250
    hits = db.search("c", v[3].tolist(), limit=3, score_threshold=0.99)
    assert sorted(h.payload["i"] for h in hits) == [3, 7] and all(h.score > 0.999 for h in hits)
    ids, sc, cnt = db.search_batch("c", v[:40], 5)
    from revers_o_b200 import ops
    ref = O.search_batch(ops.untile_rows(db._coll("c").vectors, 250, 100).float().cpu().numpy(), v[:40], 5, None,
                         db_is_normalized=True)
    for i, (a, b) in enumerate(ref):
        assert np.allclose(sc[i], b, atol=1e-3)
    dup, dsc = db.find_near_duplicates("c", 0.999)
    assert sorted(map(sorted, dup)) == [["p3", "p7"]] and dsc[0] > 0.999
    with pytest.raises(Exception):
        db.upsert("c", [models.PointStruct(id="bad", vector=[1.0, 2.0], payload=None)])


def test_saved_collection_loads_shard_wise(dev, tmp_path):
    """save() -> ShardedIndex.from_disk for each of 3 virtual ranks -> per-shard search + K3 merge == search of the whole DB."""
    import numpy as np
    from revers_o_b200 import ops
    from revers_o_b200.sharded import ShardedIndex
    from revers_o_b200.vector_db import B200VectorDB, models
    rs = np.random.RandomState(0)
    n, d, nq, k = 5000, 256, 9, 20
    vdb = B200VectorDB(device=dev)
    vdb.recreate_collection("c", vectors_config=models.VectorParams(size=d, distance=models.Distance.COSINE))
    vdb.upsert_batch("c", [f"id{i}" for i in range(n)], rs.randn(n, d).astype(np.float32))
    vdb.save(str(tmp_path))
    q = torch.from_numpy(rs.randn(nq, d).astype(np.float32)).to(dev)
    fi, fs, fc = vdb.search_batch("c", q, k, as_device=True)
    parts = []
    for r in range(3):
        idx = ShardedIndex.from_disk(str(tmp_path), "c", dev, rank=r, world=3)
        parts.append(idx.search_local(q, k))
    mi, ms, mc = ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                                torch.stack([p[2] for p in parts]), k)
    torch.cuda.synchronize()
    assert torch.equal(mi, fi) and torch.equal(mc, fc) and torch.allclose(ms, fs, atol=1e-6)
    # the torchrun-side index carries the same id / payload tables (memory-mapped id column, lazily parsed payload log)
    hits = idx.hits(mi[0].cpu().numpy(), ms[0].cpu().numpy(), int(mc[0]))
    want = vdb.search("c", q[0].cpu().numpy().tolist(), limit=k)
    assert [h.id for h in hits] == [h.id for h in want] and hits[0].id.startswith("id") and hits[0].payload == want[0].payload


def test_search_batch_input_kinds_agree(dev):
    """numpy, pageable torch, PINNED torch (no staging copy) and device tensors give the same answer."""
    from revers_o_b200.vector_db import B200VectorDB, models
    rs = np.random.RandomState(1)
    n, d, nq, k = 3000, 128, 7, 5
    vdb = B200VectorDB(device=dev)
    vdb.recreate_collection("c", vectors_config=models.VectorParams(size=d, distance=models.Distance.COSINE))
    vdb.upsert_batch("c", list(range(n)), rs.randn(n, d).astype(np.float32))
    q = rs.randn(nq, d).astype(np.float32)
    ref = vdb.search_batch("c", q, k)
    for other in (torch.from_numpy(q), torch.from_numpy(q).pin_memory(), torch.from_numpy(q).to(dev)):
        got = vdb.search_batch("c", other, k)
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))


def test_pipelined_search_matches_the_blocking_call(dev):
    """search_batch_async (two batches in flight on their own streams, scratch and pinned buffers) returns exactly what
    search_batch returns, keeps writers out until the result is collected, and survives an upsert between batches."""
    from collections import deque
    from revers_o_b200.vector_db import B200VectorDB, models
    rs = np.random.RandomState(4)
    n, d, nq, k = 40_000, 256, 48, 20
    vdb = B200VectorDB(device=dev)
    vdb.recreate_collection("c", vectors_config=models.VectorParams(size=d, distance=models.Distance.COSINE))
    vdb.upsert_batch("c", list(range(n)), rs.randn(n, d).astype(np.float32))
    batches = [rs.randn(nq, d).astype(np.float32) for _ in range(7)]
    want = [vdb.search_batch("c", b, k, 0.05) for b in batches]
    got, inflight = [], deque()
    for i, b in enumerate(batches):
        inflight.append(vdb.search_batch_async("c", torch.from_numpy(b).pin_memory() if i % 2 else b, k, 0.05))
        if len(inflight) == 2:
            got.append(inflight.popleft().result())
    while inflight:
        got.append(inflight.popleft().result())
    for w, g in zip(want, got):
        assert all(np.array_equal(a, b) for a, b in zip(w, g))
    h1 = vdb.search_batch_async("c", batches[0], k)
    h2 = vdb.search_batch_async("c", batches[1], k)
    with pytest.raises(Exception):
        vdb.search_batch_async("c", batches[2], k)          # depth 2: a third batch needs a result collected first
    r1, r2 = h1.result(), h2.result()
    with pytest.raises(Exception):
        h1.result()
    h3 = vdb.search_batch_async("c", batches[2], k)
    with pytest.raises(Exception):          # a write while this thread's own search is in flight would wait for itself: refused
        vdb.upsert("c", [models.PointStruct(id=-5, vector=batches[3][1].tolist(), payload=None)])
    h3.result()
    h4 = vdb.search_batch_async("c", batches[2], k)
    del h4                                  # a dropped handle releases the collection
    import gc
    gc.collect()
    # an upsert between batches is seen by the next one (and could not run while a batch was in flight)
    new = batches[3][0] / np.linalg.norm(batches[3][0])
    vdb.upsert("c", [models.PointStruct(id=n + 7, vector=new.tolist(), payload=None)])
    ids, sc, cnt = vdb.search_batch_async("c", batches[3], 1).result()
    assert ids[0, 0] == n and sc[0, 0] > 0.999
