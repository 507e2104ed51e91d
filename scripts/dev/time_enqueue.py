import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from revers_o_b200 import _lib, ops, synth
from revers_o_b200.vector_db import B200VectorDB, models
dev = torch.device("cuda:0")
n, d, nq, k = 1_000_000, 1024, 256, 100
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
for _ in range(5): ops.search_topk(db, n, d, q, k)
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(50):
    t0 = time.perf_counter(); r = ops.search_topk(db, n, d, q, k); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    enq.append(t1 - t0); tot.append(t2 - t0)
print("ops.search_topk: enqueue %.1f us, total %.1f us" % (1e6 * np.median(enq), 1e6 * np.median(tot)))
vdb = B200VectorDB(device=dev)
vdb.recreate_collection("b", vectors_config=models.VectorParams(size=d, distance=models.Distance.COSINE))
c = vdb._coll("b"); c.vectors, c.n = db, n
qh = q.cpu().numpy()
for _ in range(5): vdb.search_batch("b", qh, k)
ts = []
for _ in range(50):
    t0 = time.perf_counter(); vdb.search_batch("b", qh, k); ts.append(time.perf_counter() - t0)
print("search_batch(host): %.1f us" % (1e6 * np.median(ts)))
# pieces
pin = torch.empty((nq, d), dtype=torch.float32).pin_memory(); qd = torch.empty_like(q)
ts = []
for _ in range(50):
    t0 = time.perf_counter(); pin.numpy()[...] = qh; t1 = time.perf_counter(); qd.copy_(pin, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1))
print("stage copy %.1f us, H2D+sync %.1f us" % (1e6 * np.median([a for a, b in ts]), 1e6 * np.median([b for a, b in ts])))
