"""Small end-to-end pass over every kernel (for compute-sanitizer memcheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
for (n, d, nq, k) in [(3000, 96, 2, 10), (5000, 1024, 9, 10), (20000, 1024, 64, 20), (40000, 256, 300, 50), (300, 128, 8, 200)]:
    q = synth.make_queries(nq, d, seed=1, device=dev)
    db = synth.make_db(n, d, q, n_plant=16, seed=2, device=dev)
    ids, sc, cnt = ops.search_topk(db, n, d, q, k, 0.2)
    torch.cuda.synchronize()
    print("search", n, d, nq, k, cnt[:4].tolist())
for (B, M, G, D) in [(3, 8, 24, 1024), (2, 50, 16, 1280), (2, 5, 7, 96), (160, 12, 24, 256), (5, 3, 7, 128)]:
    f, m = synth.make_maskpool_inputs(B, M, G, D, seed=3, device=dev, n_empty=min(2, M - 1))
    out, counts, src, total = ops.mask_pool(f, m)
    torch.cuda.synchronize()
    print("pool", B, M, G, D, int(total.item()))
    dbp = ops.db_alloc(100 + B * M, D, dev)
    c2, s2, t2, _ = ops.mask_pool_to_db(f, m, dbp, 100)      # fused ingest (two-call route for shapes outside the TC kernel)
    torch.cuda.synchronize()
    print("pool->db", int(t2.item()))
a = ops.search_topk(db, n, d, q, 5)
mi, ms, mc = ops.merge_topk(torch.stack([a[0], a[0]]), torch.stack([a[1], a[1]]), torch.stack([a[2], a[2]]), 5)
torch.cuda.synchronize()
print("merge ok", mc[:3].tolist())

# round 2: small-Q regimes (dense, sampled + filter, dense route), fp16 / wide pooling (region groups), hot-list off, self-join,
# one-process sharded collection, pipelined lanes
for (n, d, nq, k, path) in [(9000, 1024, 1, 10, 0), (140000, 256, 3, 20, 0), (140000, 256, 2, 20, _lib.RVO_PATH_DENSE),
                            (20000, 2048, 2, 5, 0), (6000, 128, 4, 512, 0)]:
    q = synth.make_queries(nq, d, seed=4, device=dev)
    db = synth.make_db(n, d, q, n_plant=16, seed=5, device=dev)
    ids, sc, cnt = ops.search_topk(db, n, d, q, k, None, 7, path=path)
    torch.cuda.synchronize()
    print("small-q", n, d, nq, k, path, cnt.tolist())
_lib.set_option("hot", 0)
q = synth.make_queries(40, 512, seed=6, device=dev)
db = synth.make_db(70000, 512, q, n_plant=16, seed=7, device=dev)
print("hot off", ops.search_topk(db, 70000, 512, q, 30)[2][:3].tolist())
_lib.set_option("hot", 1)
print("hot on", ops.search_topk(db, 70000, 512, q, 30)[2][:3].tolist())
for (B, M, G, D, dt) in [(3, 64, 24, 1280, torch.bfloat16), (2, 33, 12, 4096, torch.bfloat16), (4, 20, 24, 1024, torch.float16)]:
    f, m = synth.make_maskpool_inputs(B, M, G, D, seed=8, device=dev, n_empty=2)
    out, counts, src, total = ops.mask_pool(f.to(dt), m)
    dbp = ops.db_alloc(B * M, D, dev)
    ops.mask_pool_to_db(f.to(dt), m, dbp, 0)
    torch.cuda.synchronize()
    print("pool r2", B, M, G, D, dt, int(total.item()))
sj = synth.make_selfjoin_db(9000, 256, 0.05, dev, seed=9, cluster=200)
p, s_ = ops.selfjoin_exact(sj, 9000, 256, 0.95, cand_cap=512)
print("selfjoin", len(p))
from revers_o_b200.vector_db import B200VectorDB, models
v = B200VectorDB(devices=[0, 0], shard_rows=1024)
v.recreate_collection("c", vectors_config=models.VectorParams(size=128, distance=models.Distance.COSINE))
v.upsert_batch("c", list(range(3000)), np.random.RandomState(0).randn(3000, 128).astype(np.float32))
qq = np.random.RandomState(1).randn(9, 128).astype(np.float32)
print("sharded", v.search_batch("c", qq, 5)[2].tolist(), v.search("c", qq[0].tolist(), limit=2)[0].id)
one = B200VectorDB()
one.recreate_collection("c", vectors_config=models.VectorParams(size=128, distance=models.Distance.COSINE))
one.upsert_batch("c", list(range(3000)), np.random.RandomState(0).randn(3000, 128).astype(np.float32))
h1 = one.search_batch_async("c", qq, 5); h2 = one.search_batch_async("c", qq[::-1].copy(), 5)
print("lanes", h1.result()[2].tolist(), h2.result()[2].tolist())
