"""Host-side tables of a collection: point ids and payloads, array-backed so that they scale to the metric's 100M rows.

The reference hands qdrant one `PointStruct(id=str(uuid4()), vector=list, payload=dict)` per region
(core_system.py:608-609) and reads `.payload` back from every hit (core_system.py:671-676).  qdrant-local keeps both in
python dicts and pickles them into sqlite; at 100M rows that is tens of GB of python objects and a minutes-long load.  Here:

  IdTable       row -> id in ONE numpy array (fixed-width bytes for strings such as uuids, int64 for integer ids; python
                objects only as a fallback for mixed/other types), id -> row through a sorted index built lazily plus a small dict
                for the rows appended since (an upsert of an existing id overwrites its row, as qdrant does).
  PayloadStore  an append-only JSONL log on disk plus an int64 offset per row; payloads of a loaded collection are parsed
                only when a hit is returned.  Rows created in this process keep their dict in RAM until they are persisted.

Pure host code (numpy + json): covered by the CPU test suite.
"""
from __future__ import annotations

import json
import os
from typing import Any, Iterable

import numpy as np


# ----------------------------------------------------------------------------------------------------------------------
class IdTable:
    KIND_EMPTY, KIND_STR, KIND_INT, KIND_OBJ = "empty", "str", "int", "obj"
    _REINDEX_TAIL = 1 << 20      # rows appended since the last sorted index before it is rebuilt

    def __init__(self):
        self.kind = self.KIND_EMPTY
        self.n = 0
        self._arr: np.ndarray | None = None      # capacity >= n
        self._sorted: np.ndarray | None = None   # argsort of _arr[:_sorted_n] (KIND_STR / KIND_INT)
        self._sorted_n = 0
        self._tail: dict = {}                    # id (canonical form) -> row for rows >= _sorted_n (all rows for KIND_OBJ)
        self._tail_complete_from = 0             # _tail covers rows [_tail_complete_from, n)

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_array(cls, arr: np.ndarray) -> "IdTable":
        t = cls()
        a = np.asarray(arr)
        t.n = len(a)
        if a.dtype.kind == "S":
            t.kind = cls.KIND_STR
        elif a.dtype.kind in "iu":
            t.kind, a = cls.KIND_INT, a.astype(np.int64, copy=False)
        elif a.dtype.kind == "O":
            t.kind = cls.KIND_OBJ
        else:
            raise TypeError(f"unsupported id array dtype {a.dtype}")
        t._arr = a
        t._sorted, t._sorted_n = None, 0
        t._tail, t._tail_complete_from = {}, t.n      # nothing indexed yet: built on the first lookup
        if t.kind == cls.KIND_OBJ:
            t._tail = {v: i for i, v in enumerate(a.tolist())}
            t._tail_complete_from = 0
        return t

    @staticmethod
    def _classify(ids: list):
        if all(isinstance(i, str) for i in ids):
            return IdTable.KIND_STR
        if all(isinstance(i, (int, np.integer)) and not isinstance(i, bool) for i in ids):
            return IdTable.KIND_INT
        return IdTable.KIND_OBJ

    def _to_kind(self, kind: str) -> None:
        """Widen the representation (empty -> anything, str/int -> obj)."""
        if self.kind == kind:
            return
        if self.kind == self.KIND_EMPTY:
            self.kind = kind
            self._arr = (np.zeros(0, dtype="S36") if kind == self.KIND_STR else
                         np.zeros(0, dtype=np.int64) if kind == self.KIND_INT else np.zeros(0, dtype=object))
            return
        # mixed id types: fall back to python objects (small collections only)
        vals = list(self)
        self.kind = self.KIND_OBJ
        self._arr = np.empty(len(vals), dtype=object)
        self._arr[:] = vals
        self._sorted, self._sorted_n = None, 0
        self._tail = {v: i for i, v in enumerate(vals)}
        self._tail_complete_from = 0

    def _encode(self, ids: list) -> np.ndarray:
        if self.kind == self.KIND_STR:
            enc = np.array([s.encode("utf-8") for s in ids]) if ids else np.zeros(0, dtype="S1")
            if enc.dtype.itemsize > self._arr.dtype.itemsize:      # longer strings than seen so far: widen the column
                self._arr = self._arr.astype(f"S{enc.dtype.itemsize}")
            return enc.astype(self._arr.dtype)
        if self.kind == self.KIND_INT:
            return np.asarray(ids, dtype=np.int64)
        out = np.empty(len(ids), dtype=object)
        out[:] = ids
        return out

    def _decode(self, v):
        if self.kind == self.KIND_STR:
            return bytes(v).decode("utf-8")
        if self.kind == self.KIND_INT:
            return int(v)
        return v

    def _reserve(self, n: int) -> None:
        if self._arr is None or len(self._arr) < n:
            cap = max(n, int(len(self._arr) * 1.5) if self._arr is not None else 0, 1024)
            new = np.zeros(cap, dtype=self._arr.dtype) if self.kind != self.KIND_OBJ else np.empty(cap, dtype=object)
            new[: self.n] = self._arr[: self.n]
            self._arr = new

    # ---- id -> row --------------------------------------------------------------------------------------------------
    def _ensure_index(self) -> None:
        if self.kind in (self.KIND_EMPTY, self.KIND_OBJ):
            return
        if self._tail_complete_from > self._sorted_n or len(self._tail) > self._REINDEX_TAIL:
            live = self._arr[: self.n]
            self._sorted = np.argsort(live, kind="stable")
            self._sorted_n = self.n
            self._tail, self._tail_complete_from = {}, self.n

    def lookup(self, ids: list) -> np.ndarray:
        """Rows of `ids` (-1 where unknown), int64 [len(ids)]."""
        out = np.full(len(ids), -1, dtype=np.int64)
        if self.n == 0 or not ids:
            return out
        kind = self._classify(ids)
        if self.kind == self.KIND_OBJ:
            for j, i in enumerate(ids):
                out[j] = self._tail.get(i, -1)
            return out
        if kind != self.kind:
            return out        # an id of another type cannot be present
        self._ensure_index()
        enc = self._encode(ids) if kind != self.KIND_STR else np.array([s.encode("utf-8") for s in ids])
        if self._sorted_n:
            keys = self._arr[: self._sorted_n]
            if kind == self.KIND_STR:
                fits = np.char.str_len(enc) <= keys.dtype.itemsize     # a longer string cannot equal a stored id
                e = enc.astype(keys.dtype)                             # (the cast truncates: masked out by `fits`)
            else:
                fits, e = True, enc
            pos = np.searchsorted(keys, e, sorter=self._sorted)
            pos = np.minimum(pos, self._sorted_n - 1)
            rows = self._sorted[pos]
            hit = (keys[rows] == e) & fits
            out[hit] = rows[hit]
        if self._tail:
            for j, i in enumerate(ids):
                r = self._tail.get(i)
                if r is not None:
                    out[j] = r
        return out

    # ---- append -----------------------------------------------------------------------------------------------------
    def append(self, ids: Iterable, assume_new: bool = False) -> np.ndarray:
        """Register `ids`; returns their rows.  Known ids keep their row (overwrite semantics of qdrant's upsert), new ones get
        rows n, n+1, ... in order.  `assume_new` skips the lookup (bulk ingest of freshly generated ids)."""
        ids = list(ids)
        if not ids:
            return np.zeros(0, np.int64)
        kind = self._classify(ids)
        if self.kind == self.KIND_EMPTY:
            self._to_kind(kind)
        elif kind != self.kind:
            self._to_kind(self.KIND_OBJ)
        rows = np.full(len(ids), -1, np.int64) if assume_new else self.lookup(ids)
        new_pos = []
        seen: dict = {}
        for j, i in enumerate(ids):       # an id repeated inside one batch maps to one row (the last vector wins, as in qdrant)
            if rows[j] >= 0:
                continue
            r = seen.get(i)
            if r is None:
                r = self.n + len(new_pos)
                seen[i] = r
                new_pos.append(j)
            rows[j] = r
        if new_pos:
            new_ids = [ids[j] for j in new_pos]
            enc = self._encode(new_ids)
            self._reserve(self.n + len(new_ids))
            self._arr[self.n: self.n + len(new_ids)] = enc
            if not assume_new or self.kind == self.KIND_OBJ:
                self._tail.update(seen)
            else:
                self._tail_complete_from = self.n + len(new_ids)   # bulk rows are indexed lazily (next lookup re-sorts)
            self.n += len(new_ids)
        return rows

    def truncate(self, n: int) -> None:
        """Roll back to the first n rows (a failed write must not leave ids without vectors)."""
        if n >= self.n:
            return
        self._tail = {k: r for k, r in self._tail.items() if r < n}
        if self._sorted_n > n:                  # the sorted index mentions dropped rows: rebuild lazily
            self._sorted, self._sorted_n = None, 0
            if self.kind != self.KIND_OBJ:
                self._tail = {}
                self._tail_complete_from = n
        self._tail_complete_from = min(self._tail_complete_from, n)
        self.n = n

    # ---- row -> id --------------------------------------------------------------------------------------------------
    def __len__(self):
        return self.n

    def __getitem__(self, row):
        if isinstance(row, slice):
            return [self._decode(v) for v in self._arr[: self.n][row]]
        row = int(row)
        if row < 0:
            row += self.n
        if not 0 <= row < self.n:
            raise IndexError(row)
        return self._decode(self._arr[row])

    def __iter__(self):
        for v in (self._arr[: self.n] if self._arr is not None else ()):
            yield self._decode(v)

    def array(self) -> np.ndarray:
        """The live rows as one numpy array (fixed-width bytes / int64 / object)."""
        return self._arr[: self.n] if self._arr is not None else np.zeros(0, dtype="S36")

    # ---- persistence ------------------------------------------------------------------------------------------------
    def disk_dtype(self):
        if self.kind == self.KIND_STR:
            return self._arr.dtype.str
        if self.kind == self.KIND_INT:
            return "<i8"
        return None        # objects go into the JSONL log instead


# ----------------------------------------------------------------------------------------------------------------------
class PayloadStore:
    """row -> payload dict.  RAM for rows created in this process, lazily parsed JSONL for rows loaded from disk."""

    def __init__(self):
        self.n = 0
        self._ram: dict[int, Any] = {}            # rows whose payload lives in this process (new or overwritten)
        self._log_path: str | None = None         # JSONL log holding the persisted rows
        self._offsets = np.zeros(0, np.int64)     # byte offset of the newest record of each persisted row (-1: none)
        self._fh = None
        self.dirty_rows: list[int] = []           # rows whose newest payload is not in the log yet

    def append_or_set(self, rows: np.ndarray, payloads: list) -> None:
        for r, p in zip(rows.tolist(), payloads):
            self._ram[r] = p
            self.dirty_rows.append(r)
            if r >= self.n:
                self.n = r + 1

    def mark_all_dirty(self, n: int) -> None:
        """Bring the payloads of rows [0, n) into RAM and mark them unwritten: the next flush appends a fresh record for each
        (used once when a collection's ids stop fitting a typed column and move into the log records)."""
        for r in range(min(n, self.n)):
            if r not in self._ram:
                self._ram[r] = self[r]
        self.dirty_rows.extend(range(min(n, self.n)))

    def truncate(self, n: int) -> None:
        if n >= self.n:
            return
        self._ram = {r: p for r, p in self._ram.items() if r < n}
        self.dirty_rows = [r for r in self.dirty_rows if r < n]
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, row):
        row = int(row)
        if row < 0:
            row += self.n
        if not 0 <= row < self.n:
            raise IndexError(row)
        if row in self._ram:
            return self._ram[row]
        if row < len(self._offsets) and self._offsets[row] >= 0 and self._log_path:
            if self._fh is None:
                self._fh = open(self._log_path, "rb")
            self._fh.seek(int(self._offsets[row]))
            return json.loads(self._fh.readline())["payload"]
        return None

    def __iter__(self):
        for r in range(self.n):
            yield self[r]

    # ---- persistence: append-only log + (row, offset) index ------------------------------------------------------
    def flush(self, log_path: str, idx_path: str, valid_log_bytes: int, valid_idx_bytes: int, ids: "IdTable | None" = None,
              drop_ram: bool = False) -> tuple[int, int]:
        """Append the dirty rows' records to `log_path` (after truncating it to `valid_log_bytes`, what meta.json vouches for)
        and their (row, offset) pairs to `idx_path`.  Returns the new valid sizes of both files."""
        for path, valid in ((log_path, valid_log_bytes), (idx_path, valid_idx_bytes)):
            if not os.path.exists(path):
                open(path, "wb").close()
            if os.path.getsize(path) != valid:
                with open(path, "r+b") as f:
                    f.truncate(valid)
        if self.dirty_rows:
            rows = sorted(set(self.dirty_rows))
            pairs = np.empty((len(rows), 2), np.int64)
            with open(log_path, "ab") as f:
                pos = valid_log_bytes
                for j, r in enumerate(rows):
                    rec = {"row": r, "payload": self._ram.get(r)}
                    if ids is not None and ids.kind == IdTable.KIND_OBJ:
                        rec["id"] = ids[r]
                    line = (json.dumps(rec, default=str) + "\n").encode("utf-8")
                    f.write(line)
                    pairs[j] = (r, pos)
                    pos += len(line)
            with open(idx_path, "ab") as f:
                f.write(pairs.tobytes())
            valid_log_bytes, valid_idx_bytes = pos, valid_idx_bytes + pairs.nbytes
            if len(self._offsets) < self.n:
                grown = np.full(self.n, -1, np.int64)
                grown[: len(self._offsets)] = self._offsets
                self._offsets = grown
            self._offsets[pairs[:, 0]] = pairs[:, 1]
            self._log_path = log_path
            if self._fh is not None:
                self._fh.close()
                self._fh = None
            self.dirty_rows = []
            if drop_ram:
                self._ram = {}
        return valid_log_bytes, valid_idx_bytes

    @classmethod
    def open(cls, log_path: str, idx_path: str, n: int, valid_idx_bytes: int) -> "PayloadStore":
        """Attach to a persisted log: only the (row, offset) index is read (16 bytes per record); payloads are parsed on demand."""
        s = cls()
        s.n = n
        s._log_path = log_path
        s._offsets = np.full(n, -1, np.int64)
        if valid_idx_bytes:
            pairs = np.fromfile(idx_path, dtype=np.int64, count=valid_idx_bytes // 8).reshape(-1, 2)
            pairs = pairs[pairs[:, 0] < n]
            # later records of a row supersede earlier ones: assign in file order, keeping the last
            order = np.argsort(pairs[:, 0], kind="stable")
            pr = pairs[order]
            last = np.ones(len(pr), bool)
            last[:-1] = pr[1:, 0] != pr[:-1, 0]
            s._offsets[pr[last, 0]] = pr[last, 1]
        return s

    def close(self):
        if self._fh is not None:
            self._fh.close()
            self._fh = None
