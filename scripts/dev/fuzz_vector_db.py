"""Model-based fuzz of the drop-in boundary (B200VectorDB, the qdrant-client surface core_system.py uses): random sequences of
upsert (new ids, overwrites, duplicates inside one batch, int / uuid-string ids, payloads or none), bulk upsert_batch, search,
search_batch, collection re-creation and close + reopen from disk (implicit persistence), checked after every step against a plain
numpy model of "normalise, round to bf16, cosine top-k with threshold, ties by insertion order".
`python scripts/dev/fuzz_vector_db.py [steps] [seed] [n_devices]`; exit 1 on the first mismatch."""
import os, shutil, sys, tempfile, uuid
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from revers_o_b200.vector_db import B200VectorDB, models
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
TOL = 1e-4


def bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).float().numpy()


class Model:
    def __init__(self, dim):
        self.dim, self.order, self.row, self.vec, self.payload = dim, [], {}, [], []

    def upsert(self, pid, v, payload):
        v = np.asarray(v, np.float32)
        n = np.sqrt(np.sum(v.astype(np.float64) ** 2))
        u = bf16_round((v / n).astype(np.float32)) if n > 0 and np.isfinite(n) else np.zeros_like(v)
        key = str(pid)
        if key in self.row:
            r = self.row[key]
            self.vec[r], self.payload[r] = u, payload
        else:
            self.row[key] = len(self.order)
            self.order.append(pid); self.vec.append(u); self.payload.append(payload)

    def search(self, q, k, thr):
        if not self.order:
            return []
        q = np.asarray(q, np.float64)
        qn = (q / np.sqrt(np.sum(q * q))).astype(np.float32)
        sc = np.stack(self.vec).astype(np.float32) @ qn
        idx = np.lexsort((np.arange(len(sc)), -sc))[: k + 4]
        return [(int(i), float(sc[i])) for i in idx if thr is None or sc[i] >= thr - TOL]


def check(hits, want, m, k, thr, tag):
    got = [(m.row[str(h.id)], h.score, h.payload) for h in hits]
    assert len(got) <= k, f"{tag}: {len(got)} hits for limit {k}"
    strict = [w for w in want if thr is None or w[1] >= thr + TOL][:k]
    assert len(got) >= len(strict), f"{tag}: {len(got)} hits, model has at least {len(strict)}"
    assert all(got[i][1] >= got[i + 1][1] - 1e-7 for i in range(len(got) - 1)), f"{tag}: not descending"
    wmap = dict(want)
    for i, (r, s, p) in enumerate(got):
        assert p == m.payload[r], f"{tag}: payload of row {r}: {p} vs {m.payload[r]}"
        if thr is not None:
            assert s >= thr - TOL, f"{tag}: score {s} below threshold {thr}"
        if i < len(want):
            assert abs(s - want[i][1]) <= TOL, f"{tag}: rank {i} score {s} vs {want[i][1]}"
        if r not in wmap:   # may only be a near-tie of the boundary
            edge = want[min(k, len(want)) - 1][1]
            assert abs(s - edge) <= 2 * TOL, f"{tag}: row {r} (score {s}) not in the model's top list (edge {edge})"


root = tempfile.mkdtemp(prefix="rvo_fuzz_")
try:
    kw = dict(devices=(list(range(ndev)) if torch.cuda.device_count() >= ndev else [0] * ndev), shard_rows=int(rs.choice([256, 1000, 4096]))) if ndev > 1 else dict(device="cuda:0")
    db = B200VectorDB(path=root, **kw)
    colls = {}
    history = []
    for step in range(steps):
        op = rs.choice(["upsert", "upsert", "bulk", "search", "search", "batch", "reopen", "recreate", "count"],
                       p=[0.2, 0.1, 0.1, 0.2, 0.15, 0.1, 0.05, 0.03, 0.07])
        if not colls or op == "recreate":
            name = f"c{rs.randint(3)}"
            dim = int(rs.choice([8, 64, 256, 1024]))
            db.recreate_collection(name, vectors_config=models.VectorParams(size=dim, distance=models.Distance.COSINE))
            colls[name] = Model(dim)
            history.append(f"step {step} recreate {name} d={dim}")
            continue
        name = list(colls)[rs.randint(len(colls))]
        m = colls[name]
        tag = f"step {step} {op} {name}(n={len(m.order)}, d={m.dim})"
        history.append(tag)
        if op == "upsert":
            pts = []
            for _ in range(int(rs.randint(1, 40))):
                if m.order and rs.rand() < 0.3:
                    pid = m.order[rs.randint(len(m.order))]                       # overwrite
                else:
                    pid = str(uuid.UUID(int=int(rs.randint(1 << 62)))) if rs.rand() < 0.6 else int(rs.randint(1 << 40))
                v = rs.randn(m.dim).astype(np.float32) * float(rs.choice([1e-3, 1.0, 50.0]))
                if rs.rand() < 0.03:
                    v[:] = 0
                payload = None if rs.rand() < 0.2 else {"image": f"img_{rs.randint(1000)}.jpg", "box": rs.randint(0, 99, 4).tolist()}
                pts.append(models.PointStruct(id=pid, vector=v.tolist() if rs.rand() < 0.5 else v, payload=payload))
            if len(pts) > 2 and rs.rand() < 0.3:
                pts.append(models.PointStruct(id=pts[0].id, vector=(rs.randn(m.dim)).astype(np.float32), payload={"dup": True}))
            db.upsert(name, pts)
            for p in pts:
                m.upsert(p.id, p.vector, p.payload)
        elif op == "bulk":
            cnt = int(rs.choice([100, 700, 3000]))
            ids = [int(1 << 41) + step * 10000 + i for i in range(cnt)]
            vecs = rs.randn(cnt, m.dim).astype(np.float32)
            pay = [{"i": i} for i in range(cnt)]
            db.upsert_batch(name, ids, vecs, pay, assume_new=bool(rs.rand() < 0.5))
            for i in range(cnt):
                m.upsert(ids[i], vecs[i], pay[i])
        elif op == "search":
            k = int(rs.choice([1, 5, 10, 20]))
            thr = None if rs.rand() < 0.4 else float(rs.choice([0.0, 0.1, 0.3, 0.9]))
            if m.order and rs.rand() < 0.4:     # a stored vector as the query (the UI's "find this region again"): score ~1
                q = m.vec[rs.randint(len(m.order))] + 0.05 * rs.randn(m.dim).astype(np.float32)
                if not np.any(q):
                    q = rs.randn(m.dim).astype(np.float32)
            else:
                q = rs.randn(m.dim).astype(np.float32)
            hits = db.search(collection_name=name, query_vector=q.tolist(), limit=k, score_threshold=thr)
            check(hits, m.search(q, k, thr), m, k, thr, tag)
        elif op == "batch":
            nq, k = int(rs.choice([2, 5, 33])), int(rs.choice([3, 10]))
            thr = None if rs.rand() < 0.5 else 0.05
            Q = rs.randn(nq, m.dim).astype(np.float32)
            ids, sc, cn = db.search_batch(name, Q, k, thr)
            c = db._coll(name)
            for j in range(nq):
                hits = [models.ScoredPoint(id=c.ids[int(i)], version=0, score=float(s), payload=c.payloads[int(i)])
                        for i, s in zip(ids[j, : cn[j]], sc[j, : cn[j]])]
                check(hits, m.search(Q[j], k, thr), m, k, thr, f"{tag} q{j}")
        elif op == "reopen":
            db.close()
            del db
            db = B200VectorDB(path=root, **kw)
            assert sorted(c.name for c in db.get_collections().collections) == sorted(colls), f"{tag}: collections differ after reopen"
        elif op == "count":
            assert db.count(name) == len(m.order), f"{tag}: count {db.count(name)} vs {len(m.order)}"
        for nm, mm in colls.items():
            assert db.count(nm) == len(mm.order), f"{tag}: afterwards count({nm}) = {db.count(nm)} vs model {len(mm.order)}"
        if step % 50 == 49:
            print(f"{step + 1} steps done: " + ", ".join(f"{n}: {len(mm.order)} rows x {mm.dim}" for n, mm in colls.items()), flush=True)
    print(f"fuzz: {steps} steps, 0 mismatches")
except AssertionError as e:
    print("MISMATCH", e)
    print("history:\n  " + "\n  ".join(history[-25:]))
    sys.exit(1)
finally:
    shutil.rmtree(root, ignore_errors=True)
