"""Small end-to-end pass over every kernel (for compute-sanitizer memcheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from revers_o_b200 import ops, synth
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
