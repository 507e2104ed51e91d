"""Generates tests/golden/golden.npz — small seeded inputs with the outputs of the CPU oracle, plus
hand-computable known-answer cases.

The reference ships no golden vectors and neither it nor its search dependency (qdrant-client) can be
imported here, so these fixtures pin the ORACLE's restatement, not the reference's
output: parity stays "unpinned" in the sense of the task statement.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reverso_oracle as O  # noqa: E402


def main():
    rs = np.random.RandomState(1234)
    out = {}
    # search: 2000 x 96 DB (D not a multiple of 64 on purpose), 9 queries, with planted neighbours + duplicates
    n, d, nq = 2000, 96, 9
    db = rs.randn(n, d).astype(np.float32)
    q = rs.randn(nq, d).astype(np.float32)
    for i in range(nq):
        for j, a in enumerate(np.linspace(0.4, 0.98, 12)):
            v = a * q[i] / np.linalg.norm(q[i]) + np.sqrt(1 - a * a) * rs.randn(d).astype(np.float32) / np.sqrt(d)
            db[(i * 97 + j * 13) % n] = v
    db[1500] = db[7]  # KA2 duplicate rows
    dbn = O.round_to_bf16(O._cosine_prepare(db))  # the GPU DB is its bf16 values
    out["search_db_bf16_bits"] = O.bf16_bits(O._cosine_prepare(db))
    out["search_queries"] = q
    for k, thr, tag in ((10, None, "k10"), (10, 0.7, "k10_t07"), (100, None, "k100"), (3000, None, "kall")):
        res = O.search_batch(dbn, q, k, thr, db_is_normalized=True)
        ids = np.full((nq, k), -1, np.int64)
        sc = np.full((nq, k), -np.inf, np.float32)
        cnt = np.zeros(nq, np.int32)
        for i, (a, b) in enumerate(res):
            ids[i, : len(a)], sc[i, : len(a)], cnt[i] = a, b, len(a)
        out[f"search_{tag}_ids"], out[f"search_{tag}_scores"], out[f"search_{tag}_counts"] = ids, sc, cnt
    # mask pool: 3 images, 5x5 grid, D=64, 6 regions (one empty, one full, one single patch)
    B, g, D, M = 3, 5, 64, 6
    feats = O.round_to_bf16(rs.randn(B, g * g, D).astype(np.float32))
    masks = (rs.rand(B, M, g * g) < 0.3).astype(np.uint8)
    masks[:, 1] = 0
    masks[:, 2] = 1
    masks[:, 3] = 0
    masks[:, 3, 7] = 1
    emb, counts, src = O.mask_pool(feats, masks)
    out["pool_feats_bf16_bits"] = O.bf16_bits(feats)
    out["pool_masks"] = masks
    out["pool_emb"], out["pool_counts"], out["pool_src"] = emb, counts, src
    emb50, counts50, _ = O.mask_pool(feats, masks, max_regions=4)
    out["pool_emb_cap4"], out["pool_counts_cap4"] = emb50, counts50
    # merge: 3 shards x 4 queries x k=5
    G, Q, k = 3, 4, 5
    sc = np.sort(rs.rand(G, Q, k).astype(np.float32), axis=2)[:, :, ::-1].copy()
    sc[1, 0, 0] = sc[0, 0, 0]  # tie across shards -> lower id first
    ids = rs.permutation(G * Q * k).reshape(G, Q, k).astype(np.int64)
    cnt = np.full((G, Q), k, np.int32)
    cnt[2, 1] = 2
    cnt[0, 3] = 0
    mi, ms, mc = O.merge_topk(ids, sc, cnt, k)
    out["merge_ids"], out["merge_scores"], out["merge_counts"] = ids, sc, cnt
    out["merge_out_ids"], out["merge_out_scores"], out["merge_out_counts"] = mi, ms, mc
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden.npz"), **out)
    print("wrote golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
