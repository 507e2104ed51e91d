"""CPU: host-side id / payload tables (revers_o_b200/tables.py) — array-backed ids with qdrant's upsert-overwrite semantics,
append-only payload log with lazy reads."""
import uuid

import numpy as np

from revers_o_b200.tables import IdTable, PayloadStore


def test_id_table_uuid_strings_append_lookup_overwrite():
    t = IdTable()
    ids = [str(uuid.uuid4()) for _ in range(1000)]              # core_system.py:574: one uuid4 string per point
    assert (t.append(ids[:500]) == np.arange(500)).all()
    assert (t.append(ids[400:600]) == np.arange(400, 600)).all()   # known ids keep their row (overwrite), new ones append
    assert t[123] == ids[123] and len(t) == 600 and t.kind == "str" and t.array().dtype == np.dtype("S36")
    assert (t.append(ids[600:], assume_new=True) == np.arange(600, 1000)).all()
    assert (t.lookup([ids[5], ids[999], "nope"]) == np.array([5, 999, -1])).all()
    assert list(t)[:3] == ids[:3] and t[-1] == ids[-1]
    t2 = IdTable.from_array(t.array().copy())                    # what a load from <collection>.ids produces: index built lazily
    assert (t2.lookup([ids[7], ids[998]]) == [7, 998]).all()
    assert list(t2.append([ids[3], "new-one", "new-one"])) == [3, 1000, 1000]   # duplicate inside a batch: one row
    t3 = IdTable.from_array(t.array().copy())
    longer = ids[4] + "-suffix-beyond-36-characters"             # shares a 36-character prefix with a stored id: NOT that id
    assert list(t3.append([ids[4], longer])) == [4, 1000] and t3[1000] == longer and t3[4] == ids[4]
    t.truncate(700)                                              # roll-back of a failed write
    assert len(t) == 700 and t.lookup([ids[800]])[0] == -1 and t.lookup([ids[650]])[0] == 650
    assert list(t.append([ids[800]])) == [700]


def test_id_table_longer_strings_ints_and_mixed():
    t = IdTable()
    t.append(["a" * 36])
    t.append(["b" * 50])                                         # wider than the column: widened, old ids intact
    assert t[0] == "a" * 36 and t[1] == "b" * 50 and t.lookup(["b" * 50])[0] == 1 and t.lookup(["b" * 49])[0] == -1
    ti = IdTable()
    assert list(ti.append([5, 9, 5])) == [0, 1, 0] and ti[1] == 9 and ti.kind == "int" and ti.disk_dtype() == "<i8"
    assert ti.lookup(["5"])[0] == -1                             # a string is not the integer id
    ti.append(["x"])                                             # mixed types: python-object fallback keeps everything
    assert ti.kind == "obj" and ti[2] == "x" and ti.lookup([9])[0] == 1 and ti.disk_dtype() is None


def test_id_table_large_sorted_index():
    t = IdTable()
    n = 200_000
    t.append(list(range(0, 2 * n, 2)), assume_new=True)
    got = t.lookup([0, 2 * n - 2, 7, 123456])
    assert list(got) == [0, n - 1, -1, 61728]
    assert list(t.append([123456, 1])) == [61728, n]


def test_payload_store_log_is_append_only_and_lazy(tmp_path):
    log, idx = str(tmp_path / "l"), str(tmp_path / "i")
    p = PayloadStore()
    p.append_or_set(np.arange(3), [{"a": 1}, None, {"b": [1, 2]}])
    lb, ib = p.flush(log, idx, 0, 0)
    size1 = (tmp_path / "l").stat().st_size
    p.append_or_set(np.array([1, 3]), [{"c": 3}, {"d": 4}])      # overwrite row 1, append row 3
    lb2, ib2 = p.flush(log, idx, lb, ib)
    assert (tmp_path / "l").read_bytes()[:size1] == (tmp_path / "l").read_bytes()[:size1] and lb2 > lb and ib2 == ib + 32
    q = PayloadStore.open(log, idx, 4, ib2)                      # only the (row, offset) index is read
    assert not q._ram and [q[i] for i in range(4)] == [{"a": 1}, {"c": 3}, {"b": [1, 2]}, {"d": 4}]
    # a torn write after the last meta.json (bytes beyond what it vouches for) is cut off by the next flush
    with open(log, "ab") as f:
        f.write(b'{"row": 9, "payl')
    q.append_or_set(np.array([4]), [{"e": 5}])
    lb3, ib3 = q.flush(log, idx, lb2, ib2)
    r = PayloadStore.open(log, idx, 5, ib3)
    assert r[4] == {"e": 5} and r[1] == {"c": 3} and len(r) == 5
    # the state meta.json vouched for BEFORE the second flush is still readable (crash between flush and meta replace)
    old = PayloadStore.open(log, idx, 3, ib)
    assert [old[i] for i in range(3)] == [{"a": 1}, None, {"b": [1, 2]}]


def test_payload_store_mark_all_dirty_rewrites_records_with_ids(tmp_path):
    """A collection whose ids stop fitting a typed column (an int id joins uuid strings) moves its ids into the log records:
    every row persisted so far must get a fresh record that carries its id, payloads read back from the old log."""
    log, idx = str(tmp_path / "c.jsonl"), str(tmp_path / "c.idx")
    ids = IdTable()
    ids.append(["a", "b", "c"])
    s = PayloadStore()
    s.append_or_set(np.arange(3), [{"i": 0}, None, {"i": 2}])
    lb, ib = s.flush(log, idx, 0, 0, ids=ids, drop_ram=True)          # typed ids: records carry no id
    assert b'"id"' not in open(log, "rb").read()
    s2 = PayloadStore.open(log, idx, 3, ib)
    ids.append([7])                                                    # mixed types -> object kind
    assert ids.kind == IdTable.KIND_OBJ
    s2.append_or_set(np.array([3]), [{"i": 3}])
    s2.mark_all_dirty(3)
    lb2, ib2 = s2.flush(log, idx, lb, ib, ids=ids)
    s3 = PayloadStore.open(log, idx, 4, ib2)
    assert [s3[r] for r in range(4)] == [{"i": 0}, None, {"i": 2}, {"i": 3}]
    import json
    byrow = {}
    for line in open(log, "rb").read()[:lb2].splitlines():
        rec = json.loads(line)
        byrow[rec["row"]] = rec.get("id")
    assert [byrow[r] for r in range(4)] == ["a", "b", "c", 7]


def test_id_table_random_operations_against_a_dict_model():
    """Random appends (new, known, duplicates inside a batch, assume_new bulk), lookups, roll-backs and array round trips for
    string-only, int-only and mixed collections, against a plain dict + list model."""
    for seed, flavour in enumerate(["str", "int", "mixed", "str", "int"]):
        rs = np.random.RandomState(seed)
        t, order, row = IdTable(), [], {}

        def fresh():
            if flavour == "str" or (flavour == "mixed" and rs.rand() < 0.5):
                return str(uuid.UUID(int=int(rs.randint(1 << 62)))) if rs.rand() < 0.8 else "id-" + "x" * int(rs.randint(1, 60))
            return int(rs.randint(-(1 << 40), 1 << 40))

        for step in range(300):
            op = rs.choice(["append", "bulk", "lookup", "truncate", "reload"], p=[0.45, 0.1, 0.3, 0.05, 0.1])
            if op == "append":
                ids = [order[rs.randint(len(order))] if order and rs.rand() < 0.3 else fresh() for _ in range(int(rs.randint(1, 30)))]
                if len(ids) > 2 and rs.rand() < 0.3:
                    ids.append(ids[0])
                got = t.append(ids)
                for i, r in zip(ids, got.tolist()):
                    if i not in row:
                        row[i] = len(order)
                        order.append(i)
                    assert row[i] == r, (flavour, step, i, r, row[i])
            elif op == "bulk":
                ids = []
                while len(ids) < 200:
                    i = fresh()
                    if i not in row and i not in ids:
                        ids.append(i)
                got = t.append(ids, assume_new=True)
                assert got.tolist() == list(range(len(order), len(order) + len(ids)))
                for i in ids:
                    row[i] = len(order)
                    order.append(i)
            elif op == "lookup":
                ids = [order[rs.randint(len(order))] if order and rs.rand() < 0.6 else fresh() for _ in range(20)]
                assert t.lookup(ids).tolist() == [row.get(i, -1) for i in ids], (flavour, step)
            elif op == "truncate" and order:
                n = int(rs.randint(0, len(order) + 1))
                t.truncate(n)
                for i in order[n:]:
                    del row[i]
                del order[n:]
            elif op == "reload" and order and t.disk_dtype() is not None:   # typed columns round-trip through their array
                t = IdTable.from_array(np.frombuffer(t.array().tobytes(), dtype=np.dtype(t.disk_dtype())).copy())
            assert len(t) == len(order)
            if order:
                j = int(rs.randint(len(order)))
                assert t[j] == order[j]
        assert list(t) == order


def test_payload_store_random_operations_against_a_dict_model(tmp_path):
    """Random writes / overwrites, write-through flushes, reopen from the log and roll-backs against a dict."""
    for seed in range(3):
        rs = np.random.RandomState(seed)
        log, idx = str(tmp_path / f"p{seed}.jsonl"), str(tmp_path / f"p{seed}.idx")
        s, model, lb, ib, persisted_n = PayloadStore(), [], 0, 0, 0
        for step in range(200):
            op = rs.choice(["write", "flush", "reopen", "truncate", "read"], p=[0.4, 0.2, 0.1, 0.05, 0.25])
            if op == "write":
                rows, pays = [], []
                for _ in range(int(rs.randint(1, 20))):
                    r = int(rs.randint(len(model))) if model and rs.rand() < 0.3 else len(model)
                    p = None if rs.rand() < 0.2 else {"step": step, "v": rs.randint(0, 9, 3).tolist(), "s": "é" * int(rs.randint(0, 4))}
                    if r == len(model):
                        model.append(p)
                    else:
                        model[r] = p
                    rows.append(r)
                    pays.append(p)
                s.append_or_set(np.asarray(rows), pays)
            elif op == "flush":
                lb, ib = s.flush(log, idx, lb, ib, drop_ram=bool(rs.rand() < 0.5))
                persisted_n = len(model)
            elif op == "reopen" and persisted_n == len(model) and not s.dirty_rows:
                s.close()
                s = PayloadStore.open(log, idx, len(model), ib)
            elif op == "truncate" and len(model) > persisted_n:
                n = int(rs.randint(persisted_n, len(model) + 1))      # a failed upsert rolls back rows that were never persisted
                s.truncate(n)
                del model[n:]
            assert len(s) == len(model)
            if model:
                j = int(rs.randint(len(model)))
                assert s[j] == model[j], (seed, step, j)
        s.close()
