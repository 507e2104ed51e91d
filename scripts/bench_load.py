"""Host side of the DB at scale: ingest N rows with uuid-like string ids and payloads, persist, then load the
saved collection (i) on one device, (ii) shard-wise over 4 (virtual) devices, (iii) as a torchrun-style shard (ShardedIndex.from_disk),
and time every phase.  What matters: the id column and the payload index are read as flat arrays (no per-row parsing), payloads are
parsed only for the hits returned.   python scripts/bench_load.py [N] [D]"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from revers_o_b200.sharded import ShardedIndex  # noqa: E402
from revers_o_b200.vector_db import B200VectorDB, models  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
base = tempfile.mkdtemp(prefix="rvo_load_")
free = shutil.disk_usage(base).free
need = n * d * 2 + n * 120
if free < 1.3 * need:
    n = int(free / 1.3 / (d * 2 + 120)) // 128 * 128
    print(f"# disk: {free / 1e9:.1f} GB free -> N reduced to {n}")
path = os.path.join(base, "db")
out = {"rows": n, "dim": d}
t0 = time.perf_counter()
db = B200VectorDB(path=path, device=dev)
db.autosave = False                                   # bulk ingest: one write-through at the end
db.recreate_collection("c", vectors_config=models.VectorParams(size=d, distance=models.Distance.COSINE))
g = torch.Generator(device=dev).manual_seed(0)
step = 1 << 20
for lo in range(0, n, step):
    m = min(step, n - lo)
    v = torch.randn((m, d), generator=g, device=dev)
    db.upsert_batch("c", [f"{i:08x}-0000-4000-8000-{i:012x}" for i in range(lo, lo + m)], v,
                    [{"filename": f"f{i}.jpg", "bbox": [0, 0, 1, 1]} for i in range(lo, lo + m)], assume_new=True, persist=False)
torch.cuda.synchronize()
out["ingest_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
db.save()
out["save_s"] = time.perf_counter() - t0
out["bytes_on_disk"] = sum(os.path.getsize(os.path.join(path, f)) for f in os.listdir(path))
q = torch.randn((4, d), generator=g, device=dev)
ref = db.search_batch("c", q.cpu().numpy(), 5)
ref_hit = db.search("c", q[0].cpu().numpy().tolist(), limit=1)[0]
del db
torch.cuda.empty_cache()

t0 = time.perf_counter()
one = B200VectorDB(path=path, device=dev)
torch.cuda.synchronize()
out["load_one_device_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
hit = one.search("c", q[0].cpu().numpy().tolist(), limit=1)[0]
out["first_search_incl_payload_ms"] = (time.perf_counter() - t0) * 1e3
assert hit.id == ref_hit.id and hit.payload == ref_hit.payload
assert all(np.array_equal(a, b) for a, b in zip(one.search_batch("c", q.cpu().numpy(), 5), ref))
c = one._coll("c")
out["payload_dicts_in_ram_after_load"] = len(c.payloads._ram)
t0 = time.perf_counter()
rows = c.ids.lookup([f"{i:08x}-0000-4000-8000-{i:012x}" for i in (0, n // 2, n - 1)])
out["first_id_lookup_builds_sorted_index_s"] = time.perf_counter() - t0
assert rows.tolist() == [0, n // 2, n - 1]
del one, c
torch.cuda.empty_cache()

t0 = time.perf_counter()
four = B200VectorDB(path=path, devices=[0, 0, 0, 0])
torch.cuda.synchronize()
out["load_4_shards_one_process_s"] = time.perf_counter() - t0
assert all(np.array_equal(a, b) for a, b in zip(four.search_batch("c", q.cpu().numpy(), 5), ref))
del four
torch.cuda.empty_cache()

t0 = time.perf_counter()
idx = ShardedIndex.from_disk(path, "c", dev, rank=3, world=8)
torch.cuda.synchronize()
out["load_rank3_of_8_shard_s"] = time.perf_counter() - t0
out["rank3_rows"] = idx.n_local
ids, sc, cnt = idx.search_local(q, 1)
h = idx.hits(ids[0].cpu().numpy(), sc[0].cpu().numpy(), int(cnt[0]))
assert h and h[0].payload["filename"] == f"f{int(ids[0, 0])}.jpg"
out["gb_per_s_one_device"] = n * d * 2 / 1e9 / out["load_one_device_s"]
print(json.dumps(out))
shutil.rmtree(base, ignore_errors=True)
