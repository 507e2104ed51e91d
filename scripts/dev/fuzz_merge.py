"""Differential fuzz of K3 (`rvo_merge_topk`): G sorted per-shard lists per query with short lists, empty lists, equal scores across
shards (ties go to the lower id), an overflow flag (-1) that must propagate, against a plain numpy merge.  Bit-exact."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from revers_o_b200 import ops
dev = torch.device("cuda:0")
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for c in range(cases):
    G = int(rs.choice([1, 2, 3, 4, 8, 16, 64]))
    nq = int(rs.choice([1, 2, 7, 64, 256]))
    k = int(rs.choice([1, 5, 10, 100, 512]))
    if G * k > 4096:
        G = max(1, 4096 // k)
    ids = np.full((G, nq, k), -1, np.int64)
    sc = np.full((G, nq, k), -np.inf, np.float32)
    cnt = np.zeros((G, nq), np.int32)
    coarse = rs.rand() < 0.5           # few distinct scores: many cross-shard ties
    for g in range(G):
        for q in range(nq):
            m = int(rs.choice([0, 1, k // 2, k, k])) if k > 1 else int(rs.randint(0, 2))
            s = rs.rand(m).astype(np.float32)
            if coarse:
                s = np.round(s * 4) / 4
            i = np.unique(rs.randint(0, 1 << 40, size=m).astype(np.int64))                # unique within the list ...
            while len(i) < m:
                i = np.unique(np.concatenate([i, rs.randint(0, 1 << 40, size=m - len(i)).astype(np.int64)]))
            i = rs.permutation(i) * G + g                                                  # ... and across shards
            order = np.lexsort((i, -s))
            ids[g, q, :m], sc[g, q, :m], cnt[g, q] = i[order], s[order], m
    flagged = set()
    if rs.rand() < 0.3:
        for _ in range(int(rs.randint(1, 4))):
            g, q = int(rs.randint(G)), int(rs.randint(nq))
            cnt[g, q] = -1
            flagged.add(q)
    oi, os_, oc = ops.merge_topk(torch.from_numpy(ids).to(dev), torch.from_numpy(sc).to(dev), torch.from_numpy(cnt).to(dev), k)
    torch.cuda.synchronize()
    oi, os_, oc = oi.cpu().numpy(), os_.cpu().numpy(), oc.cpu().numpy()
    try:
        for q in range(nq):
            if q in flagged:
                assert oc[q] < 0, f"q{q}: a shard's overflow flag was dropped (count {oc[q]})"
                continue
            ai = np.concatenate([ids[g, q, : cnt[g, q]] for g in range(G)])
            as_ = np.concatenate([sc[g, q, : cnt[g, q]] for g in range(G)])
            order = np.lexsort((ai, -as_))[:k]
            assert oc[q] == len(order), f"q{q}: count {oc[q]} vs {len(order)}"
            assert np.array_equal(oi[q, : oc[q]], ai[order]) and np.array_equal(os_[q, : oc[q]], as_[order]), f"q{q}: lists differ"
    except AssertionError as e:
        bad += 1
        print(f"MISMATCH case {c} G={G} nq={nq} k={k} coarse={coarse}: {e}")
print(f"fuzz: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
