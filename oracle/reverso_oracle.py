"""CPU oracle for the revers-o region-similarity hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's algorithm for the path named by
BASELINE.json `north_star`.  It is the checker for `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.  Nothing under `revers_o_b200/` may
import it: the product path is the CUDA library and must fail loudly without it.

PARITY: EMBEDDING HALF PINNED, SEARCH HALF UNPINNED.  The reference (`/root/reference`) ships no tests, fixtures or golden
vectors for this path.  What pins this module instead:

  * Embedding half (`l2_normalize`, `global_embedding`, `binarize_mask`, `extract_embeddings_reference`): checked against
    OUTPUTS OF THE REFERENCE ITSELF, run in this container — tests/golden/make_reference_trace.py loads the unmodified
    /root/reference/core_system.py + ui.py (stand-ins only for the third-party packages that are not installed), drives the
    UI callbacks and records every vector the reference computed and handed to the vector DB; the committed fixtures
    (tests/golden/reference_trace.{json,npz}) are compared with this restatement to 2e-7
    (tests/test_reference_source.py::test_oracle_embedding_half_is_pinned_to_reference_outputs), and the generator is re-run
    against the live reference every round (::test_committed_trace_is_what_the_unmodified_reference_does_today).
  * Search half (`search`, `QdrantLocalOracle`): UNPINNED.  The arithmetic lives in an un-vendored, un-pinned third-party
    dependency that is not installed in this image and cannot be fetched (no network, no wheel in /opt/wheelhouse):
    `qdrant-client>=1.3.0` (requirements.txt:39), *local mode* (`QdrantClient(path=...)`), call sites
    core_system.py:100,521,600-603,621,659-664.  Its published algorithm for a COSINE collection, restated here:  vectors are
    stored as float32 and L2-normalised at upsert; at search the query is L2-normalised, `scores = vectors @ query` over all
    live points in float32, candidates are visited in `np.argsort(scores)[::-1]` order (unstable sort => order among equal
    scores is unspecified), the walk stops at the first `score < score_threshold` (so a score equal to the threshold is KEPT) or
    after `limit` hits.  It is anchored on the reference's call sites (recorded argument for argument in the trace above: python
    lists of floats, limit, score_threshold, what the reference does with `.score` / `.payload`), on hand-computable
    known-answer cases (tests/golden/golden.npz, tests/test_oracle.py) and on scikit-learn / scipy cross-checks.
  * `mask_pool` restates the reference's STATED design (main.py:8-9) — the reference does not implement it — and
    is pinned to the reference only in its degenerate case (all-ones mask == `features.mean(dim=1)`, core_system.py:346).
  * `facebookresearch/perception_models` @ unpinned HEAD (setup.sh:230) — the PE encoder — is out of scope (random-init features
    per north_star); only the output layouts accepted by core_system.py:345-353 matter and are restated in `global_embedding`.

numpy only (no torch import: SURVEY.md §6 notes a 40x gemv slowdown when torch is imported first).
"""
from __future__ import annotations

import numpy as np

EPSILON = np.float32(1.1920929e-7)  # qdrant local uses float32 eps to avoid 0/0 on zero vectors

MAX_REGIONS = 50  # core_system.py:363  `min(len(self.detected_regions), 50)`


# ----------------------------------------------------------------------------------------------
# Embedding half
# ----------------------------------------------------------------------------------------------
def l2_normalize(x: np.ndarray) -> np.ndarray:
    """`e / e.norm()` — core_system.py:407 (region), :381 (fallback), :447 (direct).  No epsilon:
    a zero vector yields NaN/inf exactly like the reference (empty masks are skipped before this,
    core_system.py:402-404)."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (x / np.linalg.norm(x.astype(np.float32), axis=-1, keepdims=True)).astype(np.float32)


def global_embedding(features: np.ndarray) -> np.ndarray:
    """core_system.py:345-353 / :443-445.  `[B,N,D]` tokens -> `mean(dim=1)`; `[B,D]` pooled -> as is;
    anything else is an error."""
    f = np.asarray(features, dtype=np.float32)
    if f.ndim == 3:
        return f.mean(axis=1, dtype=np.float32)
    if f.ndim == 2:
        return f
    raise ValueError(f"Unexpected feature shape: {f.shape}")


def binarize_mask(mask: np.ndarray) -> np.ndarray:
    """core_system.py:398-400.  bool -> uint8; float -> (m > 0.5); other -> astype(uint8)."""
    m = np.asarray(mask)
    if m.dtype == bool:
        return m.astype(np.uint8)
    if np.issubdtype(m.dtype, np.floating):
        return (m > 0.5).astype(np.uint8)
    return m.astype(np.uint8)


def mask_to_patch_grid(mask: np.ndarray, grid: int = 24) -> np.ndarray:
    """Image-resolution mask (H,W) -> binary patch-grid mask (grid*grid,) uint8.

    The reference never reduces a mask to the patch grid (mask pooling is intended
    but unimplemented), so this rule is OURS and fixed here: PE's transform squashes the image to a
    square (core_system.py:200), so the patch grid is a linear squash of the image; patch (r,c)
    covers rows [r*H/grid,(r+1)*H/grid) x cols [c*W/grid,(c+1)*W/grid) (integer floor bounds, at
    least one pixel); a patch is IN the region when more than half of its pixels are set — the same
    0.5 rule the reference applies to float masks (core_system.py:399).  If that leaves a
    non-empty image mask with no patch, the patch with the largest covered fraction is taken so
    that a non-empty region (core_system.py:402) never silently becomes empty."""
    m = binarize_mask(mask)
    H, W = m.shape
    rb = (np.arange(grid + 1) * H) // grid
    cb = (np.arange(grid + 1) * W) // grid
    frac = np.zeros((grid, grid), dtype=np.float64)
    ii = np.zeros((H + 1, W + 1), dtype=np.int64)
    ii[1:, 1:] = np.cumsum(np.cumsum(m.astype(np.int64), axis=0), axis=1)
    for r in range(grid):
        r0, r1 = rb[r], max(rb[r + 1], rb[r] + 1)
        r1 = min(r1, H)
        for c in range(grid):
            c0, c1 = cb[c], max(cb[c + 1], cb[c] + 1)
            c1 = min(c1, W)
            area = max((r1 - r0) * (c1 - c0), 1)
            s = ii[r1, c1] - ii[r0, c1] - ii[r1, c0] + ii[r0, c0]
            frac[r, c] = s / area
    out = (frac > 0.5).astype(np.uint8)
    if out.sum() == 0 and m.sum() > 0:
        out.flat[int(np.argmax(frac))] = 1
    return out.reshape(-1)


def mask_pool(feats: np.ndarray, masks: np.ndarray, max_regions: int | None = None):
    """Mask-pooled region embeddings, the reference's STATED design (main.py:8-9, tutorial.md:53-58,
    north_star) restated so that it degenerates exactly to what the reference computes today:

      e_m = sum_p w[m,p] * F[p,:] / sum_p w[m,p]      (all-ones mask == features.mean(dim=1),
                                                       core_system.py:345-346)
      e_m = e_m / ||e_m||_2                            (core_system.py:407)
      regions with an empty mask are DROPPED and later regions shift up (core_system.py:402-404)
      only the first `max_regions` regions of an image are visited (core_system.py:363)

    feats  [B,P,D] float (any float dtype, accumulated in float32)
    masks  [B,M,P] binary (nonzero == in region)
    returns (emb [sum M', D] float32, counts [B] int32, src [sum M', 2] int32 (image, region))
    """
    feats = np.asarray(feats)
    masks = np.asarray(masks)
    B, P, D = feats.shape
    M = masks.shape[1]
    lim = M if max_regions is None else min(M, max_regions)
    out, counts, src = [], np.zeros(B, dtype=np.int32), []
    for b in range(B):
        f = feats[b].astype(np.float32)
        for m in range(lim):
            w = masks[b, m] != 0
            n = int(w.sum())
            if n == 0:
                continue
            e = f[w].sum(axis=0, dtype=np.float32) / np.float32(n)
            out.append(l2_normalize(e))
            src.append((b, m))
            counts[b] += 1
    emb = np.stack(out).astype(np.float32) if out else np.zeros((0, D), np.float32)
    return emb, counts, np.asarray(src, dtype=np.int32).reshape(-1, 2)


def extract_embeddings_reference(features: np.ndarray, masks: np.ndarray, max_regions: int = MAX_REGIONS):
    """What core_system.py:363-408 computes TODAY for one image: every region with a non-empty
    mask receives the normalised GLOBAL embedding (`region_embedding = global_embedding[0]`,
    core_system.py:406-407).  features: [1,N,D] or [1,D]; masks: [M, ...] any shape per region."""
    g = l2_normalize(global_embedding(features)[0])
    out = []
    for i in range(min(len(masks), max_regions)):
        if binarize_mask(masks[i]).sum() == 0:
            continue
        out.append(g.copy())
    return out


# ----------------------------------------------------------------------------------------------
# Search half — qdrant-client local mode, COSINE distance
# ----------------------------------------------------------------------------------------------
def _cosine_prepare(vectors: np.ndarray) -> np.ndarray:
    v = np.array(vectors, dtype=np.float32, copy=True)
    if v.ndim == 1:
        n = np.linalg.norm(v)
        return v / np.where(n != 0.0, n, EPSILON)
    n = np.linalg.norm(v, axis=-1)[:, np.newaxis]
    return v / np.where(n != 0.0, n, EPSILON)


def search(db: np.ndarray, query: np.ndarray, limit: int, score_threshold: float | None = None,
           db_is_normalized: bool = False):
    """One query, exactly as the reference issues it (core_system.py:657-664 -> qdrant local).

    db [N,D] float32 (normalised here unless `db_is_normalized`), query [D].
    Returns (ids int64 [n], scores float32 [n]) with n <= limit, scores descending."""
    v = np.asarray(db, dtype=np.float32) if db_is_normalized else _cosine_prepare(db)
    q = _cosine_prepare(np.asarray(query, dtype=np.float32))
    if v.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float32)
    scores = np.dot(v, q).astype(np.float32)
    order = np.argsort(scores)[::-1]
    ids, out = [], []
    for idx in order:
        if len(ids) >= limit:
            break
        s = scores[idx]
        if score_threshold is not None and s < np.float32(score_threshold):
            break
        ids.append(int(idx))
        out.append(s)
    return np.asarray(ids, dtype=np.int64), np.asarray(out, dtype=np.float32)


def search_batch(db: np.ndarray, queries: np.ndarray, limit: int, score_threshold: float | None = None,
                 db_is_normalized: bool = False):
    """The batched entry point reduced to the Q=1 semantics row by row."""
    v = np.asarray(db, dtype=np.float32) if db_is_normalized else _cosine_prepare(db)
    return [search(v, q, limit, score_threshold, db_is_normalized=True) for q in np.asarray(queries)]


def search_batch_fair(db_normalized: np.ndarray, queries: np.ndarray, limit: int):
    """'Fair batched CPU' variant: one sgemm + argpartition.  Same results as
    `search_batch` up to tie order; used only as a second, labelled CPU timing."""
    q = _cosine_prepare(np.asarray(queries, dtype=np.float32))
    s = q @ db_normalized.T
    k = min(limit, s.shape[1])
    part = np.argpartition(-s, k - 1, axis=1)[:, :k]
    ps = np.take_along_axis(s, part, axis=1)
    o = np.argsort(-ps, axis=1)
    return np.take_along_axis(part, o, axis=1).astype(np.int64), np.take_along_axis(ps, o, axis=1)


class QdrantLocalOracle:
    """Duck-type of the five `QdrantClient` methods core_system.py uses:
    get_collections (:104-107), recreate_collection (:600-603), upsert (:621), search (:659-664)."""

    class _Hit:
        def __init__(self, id, score, payload):
            self.id, self.score, self.payload, self.version, self.vector = id, float(score), payload, 0, None

    class _Coll:
        def __init__(self, name):
            self.name = name

    class _Colls:
        def __init__(self, names):
            self.collections = [QdrantLocalOracle._Coll(n) for n in names]

    def __init__(self, path=None):
        self.path = path
        self._c = {}

    def get_collections(self):
        return QdrantLocalOracle._Colls(list(self._c))

    def recreate_collection(self, collection_name, vectors_config=None, size=None, **_):
        dim = size if size is not None else getattr(vectors_config, "size", None)
        self._c[collection_name] = {"dim": int(dim), "vec": np.zeros((0, int(dim)), np.float32), "ids": [], "payload": []}

    def upsert(self, collection_name, points):
        c = self._c[collection_name]
        for p in points:
            pid = p.id if hasattr(p, "id") else p["id"]
            vec = np.asarray(p.vector if hasattr(p, "vector") else p["vector"], dtype=np.float32)
            pay = p.payload if hasattr(p, "payload") else p.get("payload")
            if vec.shape != (c["dim"],):
                raise ValueError(f"Wrong input: Vector dimension error: expected dim: {c['dim']}, got {vec.shape[0]}")
            vec = _cosine_prepare(vec)
            if pid in c["ids"]:
                i = c["ids"].index(pid)
                c["vec"][i] = vec
                c["payload"][i] = pay
            else:
                c["ids"].append(pid)
                c["payload"].append(pay)
                c["vec"] = np.concatenate([c["vec"], vec[None]], axis=0)

    def search(self, collection_name, query_vector, limit=10, score_threshold=None, **_):
        c = self._c[collection_name]
        ids, scores = search(c["vec"], np.asarray(query_vector, np.float32), limit, score_threshold,
                             db_is_normalized=True)
        return [QdrantLocalOracle._Hit(c["ids"][i], s, c["payload"][i]) for i, s in zip(ids, scores)]


# ----------------------------------------------------------------------------------------------
# Multi-GPU merge (no reference counterpart: the reference is single-process)
# ----------------------------------------------------------------------------------------------
def merge_topk(ids: np.ndarray, scores: np.ndarray, counts: np.ndarray, k: int):
    """Select the k best of G per-shard lists per query.  ids [G,Q,kk] int64 (global ids),
    scores [G,Q,kk] float32, counts [G,Q] valid entries per list.  Order: score descending,
    ties by lower id (our documented policy; the reference's tie order is unspecified).
    Returns (ids [Q,k] int64 padded with -1, scores [Q,k] float32 padded with -inf, counts [Q])."""
    G, Q, kk = ids.shape
    oi = np.full((Q, k), -1, np.int64)
    os_ = np.full((Q, k), -np.inf, np.float32)
    oc = np.zeros(Q, np.int32)
    for q in range(Q):
        ci = np.concatenate([ids[g, q, : counts[g, q]] for g in range(G)])
        cs = np.concatenate([scores[g, q, : counts[g, q]] for g in range(G)])
        order = np.lexsort((ci, -cs.astype(np.float64)))[:k]
        n = len(order)
        oi[q, :n], os_[q, :n], oc[q] = ci[order], cs[order], n
    return oi, os_, oc


def selfjoin_threshold(db: np.ndarray, threshold: float, db_is_normalized: bool = False):
    """Near-duplicate self-join (BASELINE config 5; generalises `score_threshold`,
    core_system.py:663, with every stored vector as a query).  Returns the sorted list of pairs
    (i, j), i < j, with cos >= threshold, and their scores."""
    v = np.asarray(db, dtype=np.float32) if db_is_normalized else _cosine_prepare(db)
    s = v @ v.T
    i, j = np.nonzero(np.triu(s >= np.float32(threshold), k=1))
    return np.stack([i, j], axis=1).astype(np.int64), s[i, j].astype(np.float32)


# ----------------------------------------------------------------------------------------------
# bf16 helpers (the GPU DB is DEFINED as its bf16 values; the oracle searches those values upcast)
# ----------------------------------------------------------------------------------------------
def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16 -> float32 (pure numpy bit arithmetic)."""
    a = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((a + 0x7FFF + ((a >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def round_to_bf16_inplace(x: np.ndarray) -> np.ndarray:
    """Same rounding for FINITE float32 data, in place, uint32 arithmetic (fast path for bench DB generation)."""
    a = x.view(np.uint32)
    a += np.uint32(0x7FFF) + ((a >> np.uint32(16)) & np.uint32(1))
    a &= np.uint32(0xFFFF0000)
    return x


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> uint16 bf16 bit patterns (nearest-even)."""
    a = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((a + 0x7FFF + ((a >> 16) & 1)) >> 16).astype(np.uint16).reshape(np.shape(x))


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
