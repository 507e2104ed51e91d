"""B200VectorDB — the object `SimpleReverso.vector_db` holds instead of `QdrantClient(path=...)`.

It duck-types exactly the five qdrant-client methods core_system.py uses:
  QdrantClient(path=str)                         core_system.py:100,521
  .get_collections().collections[i].name         core_system.py:104-107
  .recreate_collection(collection_name, vectors_config=VectorParams(size, distance=COSINE))   :600-603
  .upsert(collection_name, points=[PointStruct(id, vector, payload)])                          :621
  .search(collection_name, query_vector, limit, score_threshold) -> [obj(.score,.payload)]     :659-664
plus the batched entry point `search_batch` (new; reduces to `search` row by row).

Vectors live on the GPU as L2-normalised bf16 rows in the TILED storage of include/revers_o_b200.h
([row/128][col/64][128][64]: one contiguous 16 KiB TMA box per tile; the DB is DEFINED as its bf16 values); ids (uuid strings) and payload dicts stay in host lists indexed by row.  All arithmetic
(normalise, scan, select, re-score) runs in the CUDA library; there is no CPU fallback.

On-disk format: <path>/meta.json + <path>/<collection>.bf16 (the tiled storage
as is, mmap-able, loadable shard-wise by 128-row block) + <path>/<collection>.payload.jsonl.
"""
from __future__ import annotations

import json
import os
import threading
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Iterable

import numpy as np
import torch

from . import ops
from ._lib import RVO_MAX_K, RvoError


class Distance:
    COSINE = "Cosine"


@dataclass
class VectorParams:
    size: int
    distance: str = Distance.COSINE


@dataclass
class PointStruct:
    id: Any
    vector: Any
    payload: dict | None = None


@dataclass
class ScoredPoint:
    """What core_system.py:671-676 consumes: `.payload` (dict) and `.score` (float)."""
    id: Any
    version: int
    score: float
    payload: dict | None
    vector: Any = None


# `from qdrant_client.http import models` replacement: models.VectorParams / Distance / PointStruct
models = SimpleNamespace(Distance=Distance, VectorParams=VectorParams, PointStruct=PointStruct,
                         ScoredPoint=ScoredPoint)


class _Collection:
    def __init__(self, name: str, dim: int, device: torch.device):
        self.name, self.dim, self.device = name, int(dim), device
        self.d_pad = ops.d_pad_of(self.dim)
        self.n = 0
        self.vectors = ops.db_alloc(0, self.dim, device)  # tiled bf16 storage [blocks, d_pad/64, 128, 64]
        self.ids: list = []
        self.payloads: list = []
        self.row_of: dict = {}

    def reserve(self, rows: int) -> None:
        have = ops.db_capacity(self.vectors)
        if rows <= have:
            return
        new = ops.db_alloc(max(rows, int(have * 1.5), 1024), self.dim, self.device)
        if self.vectors.shape[0]:
            new[: self.vectors.shape[0]].copy_(self.vectors)  # whole row blocks: device-to-device memcpy
        self.vectors = new


class B200VectorDB:
    def __init__(self, path: str | None = None, device: str | torch.device | None = None, **_):
        if not torch.cuda.is_available():
            raise RvoError("B200VectorDB needs a CUDA (sm_100) device: there is no CPU fallback")
        self.path = path
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lock = threading.RLock()  # Gradio callbacks share one instance across threads (ui.py:20)
        self._collections: dict[str, _Collection] = {}
        self._staging: dict = {}
        if path and os.path.exists(os.path.join(path, "meta.json")):
            self._load(path)

    # ---- qdrant-client surface ---------------------------------------------------------------
    def get_collections(self):
        with self._lock:
            return SimpleNamespace(collections=[SimpleNamespace(name=n) for n in self._collections])

    def recreate_collection(self, collection_name: str, vectors_config: VectorParams | None = None, **kw):
        size = vectors_config.size if vectors_config is not None else kw["size"]
        distance = getattr(vectors_config, "distance", Distance.COSINE)
        if str(getattr(distance, "value", distance)).lower() != "cosine":
            raise RvoError("only Distance.COSINE is supported (the reference uses no other, core_system.py:602)")
        with self._lock:
            self._collections[collection_name] = _Collection(collection_name, size, self.device)
        return True

    def upsert(self, collection_name: str, points: Iterable):
        """Normalise (qdrant COSINE semantics) and append/overwrite rows.  `vector` may be a python list
        (core_system.py:608), a numpy array or a torch tensor."""
        with self._lock:
            c = self._coll(collection_name)
            pts = list(points)
            if not pts:
                return SimpleNamespace(status="completed")
            vecs = np.asarray([np.asarray(_get(p, "vector"), dtype=np.float32) for p in pts], dtype=np.float32)
            if vecs.ndim != 2 or vecs.shape[1] != c.dim:
                raise RvoError(f"Wrong input: Vector dimension error: expected dim: {c.dim}, got {vecs.shape[-1]}")
            rows = []
            for p in pts:
                pid = _get(p, "id")
                r = c.row_of.get(pid)
                if r is None:
                    r = len(c.ids)
                    c.ids.append(pid)
                    c.payloads.append(_get(p, "payload"))
                    c.row_of[pid] = r
                else:
                    c.payloads[r] = _get(p, "payload")
                rows.append(r)
            c.reserve(len(c.ids))
            self._write_rows(c, rows, torch.from_numpy(vecs))
            c.n = len(c.ids)
            return SimpleNamespace(status="completed")

    def search(self, collection_name: str, query_vector, limit: int = 10, score_threshold: float | None = None, **_):
        """core_system.py:659-664.  One query; returns ScoredPoint-like hits, score descending."""
        q = np.asarray(query_vector, dtype=np.float32).reshape(1, -1)
        ids, scores, counts = self.search_batch(collection_name, q, limit, score_threshold)
        c = self._coll(collection_name)
        n = int(counts[0])
        return [ScoredPoint(id=c.ids[int(i)], version=0, score=float(s), payload=c.payloads[int(i)])
                for i, s in zip(ids[0, :n], scores[0, :n])]

    # ---- batched entry point (new) -------------------------------------------------------------
    def search_batch(self, collection_name: str, queries, limit: int = 10, score_threshold: float | None = None,
                     as_device: bool = False):
        """queries: [Q, D] float32 (numpy / torch, host or device; a PINNED torch CPU tensor is copied to the GPU without
        a staging copy).  Returns (ids [Q,k] int64 row numbers, scores [Q,k] float32, counts [Q] int32); numpy unless
        `as_device`."""
        with self._lock:
            c = self._coll(collection_name)
            n = c.n
            vectors = c.vectors
        k = int(limit)
        if k > RVO_MAX_K:
            raise RvoError(f"limit={k} above the supported maximum {RVO_MAX_K}")
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            qd = queries.to(dtype=torch.float32).contiguous()
            if qd.dim() == 1:
                qd = qd.unsqueeze(0)
            if qd.shape[1] != c.dim:
                raise RvoError(f"Wrong input: Vector dimension error: expected dim: {c.dim}, got {qd.shape[1]}")
            if as_device:
                return ops.search_topk_exact(vectors, n, c.dim, qd, k, score_threshold)
            io = self._io_plan(qd.shape[0], c.dim, k)
        elif (isinstance(queries, torch.Tensor) and queries.is_pinned() and queries.dtype == torch.float32
              and queries.dim() == 2 and queries.is_contiguous()):
            # caller-owned pinned host memory: DMA straight from it, no staging copy
            if queries.shape[1] != c.dim:
                raise RvoError(f"Wrong input: Vector dimension error: expected dim: {c.dim}, got {queries.shape[1]}")
            io = self._io_plan(queries.shape[0], c.dim, k)
            if as_device:
                io.q_dev.copy_(queries, non_blocking=True)
                return ops.search_topk_exact(vectors, n, c.dim, io.q_dev, k, score_threshold)
            if os.environ.get("RVO_ZC_IN", "0") == "1":
                qd = queries        # the normalise kernel reads the queries straight from the caller's pinned memory
            else:
                io.q_dev.copy_(queries, non_blocking=True)
                qd = io.q_dev
        else:
            qh = queries.detach().cpu().numpy() if isinstance(queries, torch.Tensor) else queries
            qh = np.ascontiguousarray(qh, dtype=np.float32)
            if qh.ndim == 1:
                qh = qh[None]
            if qh.shape[1] != c.dim:
                raise RvoError(f"Wrong input: Vector dimension error: expected dim: {c.dim}, got {qh.shape[1]}")
            # host -> pinned staging -> device, all on the current stream; the views are cached per (Q, D, k) and thread
            io = self._io_plan(qh.shape[0], c.dim, k)
            np.copyto(io.q_stage_np, qh)
            if as_device:
                io.q_dev.copy_(io.q_stage, non_blocking=True)
                return ops.search_topk_exact(vectors, n, c.dim, io.q_dev, k, score_threshold)
            if os.environ.get("RVO_ZC_IN", "0") == "1":
                qd = io.q_stage     # pinned staging: the normalise kernel reads it over PCIe, no H2D copy launch
            else:
                io.q_dev.copy_(io.q_stage, non_blocking=True)
                qd = io.q_dev
        # results land in ONE blob [ids int64 | scores f32 | counts i32]: a device blob + one D2H copy by default.  With
        # RVO_ZC_OUT=1 the last kernel stores straight into the pinned host blob (and with RVO_ZC_IN=1 the first kernel reads
        # the queries from pinned host memory): measured equal within noise on one GPU and 30-40 us per step WORSE with two
        # processes on one box, so the copies stay the default
        ops.search_topk(vectors, n, c.dim, qd, k, score_threshold, out=(io.ids, io.scores, io.counts))
        if io.res_dev is not io.res_host:
            io.res_host.copy_(io.res_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        out_i, out_s, out_c = io.ids_np.copy(), io.scores_np.copy(), io.counts_np.copy()
        if out_c.min() < 0:  # overflow protocol of rvo_search_topk: exact fp32 scan in batches of <= RVO_SMALL_Q
            bad = np.nonzero(out_c < 0)[0]
            qbad = qd[torch.from_numpy(bad).to(qd.device)].to(self.device).contiguous()
            a, b, cc = ops.search_topk_exact(vectors, n, c.dim, qbad, k, score_threshold)
            out_i[bad], out_s[bad], out_c[bad] = a.cpu().numpy(), b.cpu().numpy(), cc.cpu().numpy()
        return out_i, out_s, out_c

    def _io_plan(self, nq: int, d: int, k: int):
        """Pinned staging + device buffers + every view of them for one (Q, D, k) shape, per calling thread; building the
        ~20 tensor/numpy views costs more host time than the copies they describe, so they are made once."""
        key = ("io", nq, d, k, threading.get_ident())
        io = self._staging.get(key)
        if io is None:
            nb_i, nb_s, nb_c = nq * k * 8, nq * k * 4, nq * 4
            nb = (nb_i + nb_s + nb_c + 7) // 8 * 8
            q_stage = torch.empty((nq, d), dtype=torch.float32).pin_memory()
            res_host = torch.empty(nb, dtype=torch.uint8).pin_memory()     # device-addressable (UVA): kernels write into it
            res_dev = res_host if os.environ.get("RVO_ZC_OUT", "0") == "1" else torch.empty(nb, dtype=torch.uint8, device=self.device)
            host = res_host.numpy()
            io = SimpleNamespace(
                q_stage=q_stage, q_stage_np=q_stage.numpy(),
                q_dev=torch.empty((nq, d), dtype=torch.float32, device=self.device),
                res_dev=res_dev, res_host=res_host,
                ids=res_dev[:nb_i].view(torch.int64).view(nq, k),
                scores=res_dev[nb_i: nb_i + nb_s].view(torch.float32).view(nq, k),
                counts=res_dev[nb_i + nb_s: nb_i + nb_s + nb_c].view(torch.int32),
                ids_np=host[:nb_i].view(np.int64).reshape(nq, k),
                scores_np=host[nb_i: nb_i + nb_s].view(np.float32).reshape(nq, k),
                counts_np=host[nb_i + nb_s: nb_i + nb_s + nb_c].view(np.int32))
            if len(self._staging) > 64:      # shapes come and go (UI: Q = 1; batch jobs: a few sizes): keep the cache bounded
                self._staging.clear()
            self._staging[key] = io
        return io

    # ---- bulk ingest: tensors in, no python float lists --------------------
    def upsert_batch(self, collection_name: str, ids: list, vectors, payloads: list | None = None):
        with self._lock:
            c = self._coll(collection_name)
            v = vectors if isinstance(vectors, torch.Tensor) else torch.from_numpy(np.asarray(vectors, np.float32))
            if v.dim() != 2 or v.shape[1] != c.dim or v.shape[0] != len(ids):
                raise RvoError(f"upsert_batch: bad shape {tuple(v.shape)} for {len(ids)} ids, dim {c.dim}")
            payloads = payloads if payloads is not None else [None] * len(ids)
            rows = []
            for pid, pay in zip(ids, payloads):
                r = c.row_of.get(pid)
                if r is None:
                    r = len(c.ids)
                    c.ids.append(pid)
                    c.payloads.append(pay)
                    c.row_of[pid] = r
                else:
                    c.payloads[r] = pay
                rows.append(r)
            c.reserve(len(c.ids))
            self._write_rows(c, rows, v)
            c.n = len(c.ids)

    def ingest_regions(self, collection_name: str, feats: torch.Tensor, masks: torch.Tensor, payload_fn=None,
                       max_regions: int = 0):
        """Batched ingest: patch features [B,P,D] bf16 + patch-grid masks [B,M,P] uint8 on the GPU ->
        every non-empty region's mask-pooled, L2-normalised embedding appended to the collection by ONE kernel that writes
        bf16 rows straight into the tiled DB (no fp32 embeddings in HBM, no python float lists, core_system.py:363-408 +
        :596-622).  `payload_fn(image_index, region_index) -> (id, payload)`; default: uuid4 id, {"image": b, "region": m}.
        Returns the number of rows appended."""
        import uuid
        with self._lock:
            c = self._coll(collection_name)
            B, P, D = feats.shape
            M = masks.shape[1]
            if D != c.dim:
                raise RvoError(f"Wrong input: Vector dimension error: expected dim: {c.dim}, got {D}")
            lim = M if max_regions <= 0 else min(M, max_regions)
            row0 = len(c.ids)
            c.reserve(row0 + B * lim)
            counts, src, total, _ = ops.mask_pool_to_db(feats, masks, c.vectors, row0, max_regions)
            m = int(total.item())                      # the one host sync of an ingest batch
            origin = src[:m].cpu().numpy()
            for o in origin:
                b, r = divmod(int(o), M)
                pid, pay = payload_fn(b, r) if payload_fn else (str(uuid.uuid4()), {"image": b, "region": r})
                c.row_of[pid] = len(c.ids)
                c.ids.append(pid)
                c.payloads.append(pay)
            c.n = len(c.ids)
            return m

    def find_near_duplicates(self, collection_name: str, threshold: float = 0.95, max_pairs: int = 1 << 22):
        """All pairs of stored points with cosine >= threshold (BASELINE config 4: keyframe near-duplicate self-join;
        generalises `score_threshold`, core_system.py:663).  Returns (list of (id_a, id_b), scores float32 [n])."""
        with self._lock:
            c = self._coll(collection_name)
            n, vectors = c.n, c.vectors
        if n < 2:
            return [], np.zeros(0, np.float32)
        pairs, scores, count, over = ops.selfjoin_threshold(vectors, n, c.dim, threshold, out_cap=max_pairs)
        m = int(count.item())
        if int(over.item()) > 0 or m > max_pairs:
            raise RvoError(f"near-duplicate join overflowed ({m} pairs, {int(over.item())} candidate lists): raise max_pairs / cand_cap")
        p = pairs[:m].cpu().numpy()
        return [(c.ids[int(a)], c.ids[int(b)]) for a, b in p], scores[:m].cpu().numpy()

    def count(self, collection_name: str) -> int:
        return self._coll(collection_name).n

    # ---- persistence -----------------------------------------------------------------------------
    def save(self, path: str | None = None) -> None:
        path = path or self.path
        if not path:
            raise RvoError("no path to save to")
        os.makedirs(path, exist_ok=True)
        with self._lock:
            meta = {"format": "revers_o_b200/1", "collections": {}}
            for name, c in self._collections.items():
                blocks = (c.n + ops.TILE_ROWS - 1) // ops.TILE_ROWS
                meta["collections"][name] = {"dim": c.dim, "d_pad": c.d_pad, "n": c.n, "distance": "Cosine",
                                             "layout": "tiled[block][d_pad/64][128][64] bf16", "blocks": blocks}
                raw = c.vectors[:blocks].contiguous().view(torch.int16).cpu().numpy()
                raw.tofile(os.path.join(path, f"{name}.bf16"))  # the tiled storage as is: loadable shard-wise by block
                with open(os.path.join(path, f"{name}.payload.jsonl"), "w") as f:
                    for pid, pay in zip(c.ids, c.payloads):
                        f.write(json.dumps({"id": pid, "payload": pay}, default=str) + "\n")
            with open(os.path.join(path, "meta.json"), "w") as f:
                json.dump(meta, f)

    def _load(self, path: str) -> None:
        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        for name, m in meta.get("collections", {}).items():
            c = _Collection(name, m["dim"], self.device)
            n = int(m["n"])
            if n:
                blocks = (n + ops.TILE_ROWS - 1) // ops.TILE_ROWS
                raw = np.fromfile(os.path.join(path, f"{name}.bf16"), dtype=np.int16).reshape(
                    blocks, c.d_pad // ops.TILE_COLS, ops.TILE_ROWS, ops.TILE_COLS)
                c.reserve(n)
                c.vectors[:blocks].copy_(torch.from_numpy(raw).view(torch.bfloat16))
            with open(os.path.join(path, f"{name}.payload.jsonl")) as f:
                for line in f:
                    rec = json.loads(line)
                    c.row_of[rec["id"]] = len(c.ids)
                    c.ids.append(rec["id"])
                    c.payloads.append(rec["payload"])
            c.n = n
            self._collections[name] = c

    def close(self) -> None:
        pass

    def _pinned(self, name: str, nbytes: int) -> torch.Tensor:
        """Persistent pinned host staging buffers, one set per calling thread (Gradio worker threads)."""
        key = (name, threading.get_ident())
        buf = self._staging.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 4096) * 2, dtype=torch.uint8).pin_memory()
            self._staging[key] = buf
        return buf[: (nbytes + 7) // 8 * 8]

    def _device_buf(self, name: str, nbytes: int) -> torch.Tensor:
        key = ("dev_" + name, threading.get_ident())
        buf = self._staging.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 4096) * 2, dtype=torch.uint8, device=self.device)
            self._staging[key] = buf
        return buf[: (nbytes + 7) // 8 * 8]

    # ---- helpers ---------------------------------------------------------------------------------
    def _coll(self, name: str) -> _Collection:
        c = self._collections.get(name)
        if c is None:
            raise RvoError(f"Collection {name} not found")
        return c

    def _to_device_f32(self, x) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            t = x.to(dtype=torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if not t.is_cuda:
            t = t.pin_memory().to(self.device, non_blocking=True) if t.numel() > 4096 else t.to(self.device)
        return t.contiguous()

    def _write_rows(self, c: _Collection, rows: list, host_or_dev: torch.Tensor) -> None:
        src = self._to_device_f32(host_or_dev)
        i = 0
        while i < len(rows):  # one normalise-and-store launch per run of consecutive destination rows
            j = i + 1
            while j < len(rows) and rows[j] == rows[j - 1] + 1:
                j += 1
            ops.normalize_rows(src[i:j], db=c.vectors, row0=rows[i])
            i = j


def read_shard_blocks(path: str, collection_name: str, world: int, rank: int):
    """On-disk format, shard-wise: memory-map `<path>/<collection>.bf16` (the tiled storage as is) and
    return (view int16 [blocks_of_this_rank, d_pad/64, 128, 64], n_local, first_global_row, dim) for rank `rank` of `world`
    — whole 128-row blocks, the same split as `sharded.shard_bounds`.  Pure host code (numpy), no copy until the caller
    reads the view; each GPU process of a box maps only its slice of the file."""
    from .sharded import shard_bounds
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    m = meta.get("collections", {}).get(collection_name)
    if m is None:
        raise RvoError(f"Collection {collection_name} not found")
    n, dim, d_pad = int(m["n"]), int(m["dim"]), int(m["d_pad"])
    nk = d_pad // ops.TILE_COLS
    blocks = (n + ops.TILE_ROWS - 1) // ops.TILE_ROWS
    lo, hi = shard_bounds(n, world, rank)
    b0, b1 = lo // ops.TILE_ROWS, (hi + ops.TILE_ROWS - 1) // ops.TILE_ROWS
    if blocks == 0 or hi <= lo:
        return np.zeros((0, nk, ops.TILE_ROWS, ops.TILE_COLS), np.int16), 0, lo, dim
    mm = np.memmap(os.path.join(path, f"{collection_name}.bf16"), dtype=np.int16, mode="r",
                   shape=(blocks, nk, ops.TILE_ROWS, ops.TILE_COLS))
    return mm[b0:b1], hi - lo, lo, dim


def _get(p, name):
    return getattr(p, name) if hasattr(p, name) else p[name]
