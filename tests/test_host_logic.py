"""CPU: host-side logic of the drop-in (no GPU, no compute calls): patch-grid reduction vs the oracle's rule,
sharding maths, result packing, status-string convention of the boundary."""
import numpy as np
import torch

from oracle import reverso_oracle as O
from revers_o_b200.core_system import SimpleReverso, binarize_mask, mask_to_patch_grid
from revers_o_b200.sharded import pack_results, packed_bytes, selfjoin_blocks, shard_bounds, unpack_results


def test_patch_grid_matches_oracle_rule():
    rs = np.random.RandomState(0)
    for H, W, g in ((336, 336, 24), (480, 640, 24), (97, 61, 16), (20, 20, 24)):
        for _ in range(4):
            m = np.zeros((H, W), bool)
            y0, x0 = rs.randint(0, H - 1), rs.randint(0, W - 1)
            m[y0: y0 + rs.randint(1, H), x0: x0 + rs.randint(1, W)] = True
            assert np.array_equal(mask_to_patch_grid(m, g), O.mask_to_patch_grid(m, g))
    f = rs.rand(50, 70).astype(np.float32)
    assert np.array_equal(binarize_mask(f), O.binarize_mask(f))
    assert np.array_equal(mask_to_patch_grid(f, 8), O.mask_to_patch_grid(f, 8))


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 100, 100_000_000):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(lo % 128 == 0 or lo == n for lo, _ in b)             # shards are whole blocks of the tiled storage
            if n >= 128 * w:
                sizes = [hi - lo for lo, hi in b]
                assert max(sizes) - min(sizes) <= 128
    assert shard_bounds(100_000_000, 8, 3) == (37_500_160, 50_000_128)


def test_selfjoin_blocks_partition_rows_and_balance_triangle():
    n = 2_000_000
    for w in (1, 2, 8):
        parts = [selfjoin_blocks(n, w, r) for r in range(w)]
        rows = sorted(x for p in parts for x in p)
        assert rows[0][0] == 0 and rows[-1][1] == n and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        work = [sum((hi - lo) * (n - lo) for lo, hi in p) for p in parts]     # rows scanned per query block
        assert max(work) / min(work) < 1.02


def test_pack_unpack_roundtrip():
    nq, k, G = 5, 7, 3
    g = torch.Generator().manual_seed(0)
    parts = []
    for r in range(G):
        ids = torch.randint(0, 1 << 40, (nq, k), generator=g)
        sc = torch.randn((nq, k), generator=g)
        cnt = torch.randint(-1, k + 1, (nq,), generator=g, dtype=torch.int32)
        parts.append((ids, sc, cnt))
    blobs = torch.stack([pack_results(*p) for p in parts])
    assert blobs.shape == (G, packed_bytes(nq, k)) and packed_bytes(nq, k) % 8 == 0
    ids, sc, cnt = unpack_results(blobs, nq, k)
    for r in range(G):
        assert torch.equal(ids[r], parts[r][0]) and torch.equal(sc[r], parts[r][1]) and torch.equal(cnt[r], parts[r][2])


def test_status_string_convention_without_gpu(tmp_path):
    """Guards of core_system.py:93-97,322-324,652-653: status strings / empty results, never exceptions."""
    r = SimpleReverso(db_root=str(tmp_path / "db"), device="cpu")
    assert r.list_databases() == []
    assert r.load_database("") == "❌ Please provide a database name"
    assert r.load_database("missing") == "❌ Database not found: missing"
    assert r.delete_database("missing").startswith("❌") and r.unlock_database("").startswith("❌")
    assert r.search_similar()[0].startswith("❌ No query embeddings")
    r.region_embeddings = [torch.zeros(4)]
    assert r.search_similar()[0].startswith("❌ No database loaded")
    assert r.extract_embeddings(None) == ([], [])
    assert r.detect_regions(np.zeros((4, 4, 3), np.uint8), "x") == 0
    assert r.create_database(str(tmp_path / "nope"), "d").startswith("❌ Folder not found")
    r.request_stop()
    assert r._stop_requested is True


def test_read_shard_blocks_covers_the_file_exactly(tmp_path):
    """Shard-wise view of the on-disk tiled format: the ranks' block ranges partition the file, row counts add up, the first
    global row of a rank is the end of the previous one (host logic only — no GPU)."""
    import json
    import numpy as np
    from revers_o_b200.vector_db import read_shard_blocks
    n, dim, d_pad = 1000, 100, 128
    blocks, nk = (n + 127) // 128, d_pad // 64
    raw = np.arange(blocks * nk * 128 * 64, dtype=np.int64).astype(np.int16).reshape(blocks, nk, 128, 64)
    raw.tofile(tmp_path / "c.bf16")
    (tmp_path / "meta.json").write_text(json.dumps({"format": "revers_o_b200/1", "collections": {
        "c": {"dim": dim, "d_pad": d_pad, "n": n, "distance": "Cosine", "blocks": blocks}}}))
    for world in (1, 2, 3, 8, 11):
        got, rows, nxt = [], 0, 0
        for r in range(world):
            view, n_local, row0, d = read_shard_blocks(str(tmp_path), "c", world, r)
            assert d == dim and row0 == nxt and (row0 % 128 == 0 or n_local == 0)
            assert view.shape[0] == (n_local + 127) // 128
            got.append(np.asarray(view))
            rows += n_local
            nxt = row0 + n_local
        assert rows == n and np.array_equal(np.concatenate(got), raw)
    import pytest
    from revers_o_b200._lib import RvoError
    with pytest.raises(RvoError):
        read_shard_blocks(str(tmp_path), "missing", 2, 0)


def test_rw_lock_readers_overlap_writers_exclude_and_self_deadlock_is_refused():
    """vector_db._RWLock: searches (readers) overlap, upserts (writers) are exclusive and re-entrant, and a thread that still
    holds a read side (an in-flight search_batch_async) is refused the write side instead of waiting for itself."""
    import threading
    import time

    import pytest

    from revers_o_b200._lib import RvoError
    from revers_o_b200.vector_db import _RWLock
    lk = _RWLock()
    inside, peak, log = [0], [0], []

    def reader():
        with lk.read():
            inside[0] += 1
            peak[0] = max(peak[0], inside[0])
            time.sleep(0.05)
            inside[0] -= 1

    ts = [threading.Thread(target=reader) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert peak[0] >= 2                                     # readers ran together

    def writer():
        with lk.write():
            with lk.write():                                # re-entrant for the owner
                with lk.read():                             # the writer may read its own state
                    log.append(("w", inside[0]))
                    time.sleep(0.05)

    r = lk.read()
    r.__enter__()                                           # an outstanding search on THIS thread
    with pytest.raises(RvoError):
        with lk.write():
            pass
    w = threading.Thread(target=writer)
    w.start()
    time.sleep(0.02)
    assert not log                                          # the writer waits for the reader
    r.__exit__(None, None, None)
    w.join()
    assert log == [("w", 0)]
    with lk.write():
        pass
