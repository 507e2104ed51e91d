"""Tensor-level wrappers over the C ABI.  PyTorch is used for device-memory ownership and streams only;
every arithmetic step runs in the hand-written CUDA library (no torch ops on the data path)."""
from __future__ import annotations

import math
import threading

import torch

from . import _lib
from ._lib import RVO_MAX_K, RVO_SMALL_Q, RvoError, check

_ws_lock = threading.Lock()
_workspaces: dict = {}


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, thread) scratch buffer, grown on demand, 1024-byte aligned (torch allocations are 512-byte
    aligned, so one KiB of slack is kept and the view is offset)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), threading.get_ident())
    with _ws_lock:
        buf = _workspaces.get(key)
        if buf is None or buf.numel() < nbytes + 1024:
            buf = torch.empty(int(nbytes * 1.25) + 2048, dtype=torch.uint8, device=device)
            _workspaces[key] = buf
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RvoError(f"{name} must be a CUDA tensor: the B200 library has no CPU path")


def d_pad_of(d: int) -> int:
    return (d + 63) // 64 * 64


def normalize_rows(src: torch.Tensor, dst_bf16: torch.Tensor | None = None, want_f32: bool = False):
    """L2-normalise fp32 rows (core_system.py:407,447; qdrant COSINE upsert).  Returns (bf16 [n,d_pad], f32|None)."""
    require_cuda(src, "src")
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    n, d = src.shape
    if dst_bf16 is None:
        dst_bf16 = torch.empty((n, d_pad_of(d)), dtype=torch.bfloat16, device=src.device)
    assert dst_bf16.dtype == torch.bfloat16 and dst_bf16.shape[0] >= n and dst_bf16.stride(1) == 1
    f32 = torch.empty((n, d), dtype=torch.float32, device=src.device) if want_f32 else None
    lib = _lib.load()
    check(lib.rvo_normalize_rows(_ptr(src), n, d, src.stride(0), _ptr(dst_bf16), dst_bf16.stride(0), _ptr(f32),
                                 _stream(src.device)), "rvo_normalize_rows")
    return dst_bf16, f32


def mask_pool(feats: torch.Tensor, masks: torch.Tensor, max_regions: int = 0):
    """K1.  feats bf16 [B,P,D], masks uint8 [B,M,P] -> (out f32 [B*M, D] (first `total` rows valid),
    counts int32 [B], src int32 [B*M], total int32 [1]) — all device tensors, no host sync."""
    require_cuda(feats, "feats")
    require_cuda(masks, "masks")
    assert feats.dtype == torch.bfloat16 and feats.is_contiguous() and feats.dim() == 3
    assert masks.dtype == torch.uint8 and masks.is_contiguous() and masks.dim() == 3
    B, P, D = feats.shape
    M = masks.shape[1]
    assert masks.shape == (B, M, P)
    dev = feats.device
    out = torch.empty((B * M, D), dtype=torch.float32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    src = torch.empty((B * M,), dtype=torch.int32, device=dev)
    total = torch.empty((1,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.rvo_mask_pool_workspace_bytes(B, M, P, D)
    ws = workspace(dev, nbytes)
    check(lib.rvo_mask_pool(_ptr(feats), _ptr(masks), B, M, P, D, int(max_regions), _ptr(out), _ptr(counts), _ptr(src),
                            _ptr(total), _ptr(ws), nbytes, _stream(dev)), "rvo_mask_pool")
    return out, counts, src, total


def search_topk(db: torch.Tensor, n_rows: int, d: int, queries: torch.Tensor, k: int,
                score_threshold: float | None = None, id_offset: int = 0, out=None):
    """K2.  db bf16 [>=n_rows, d_pad] normalised rows, queries f32 [nq, d].
    Returns device tensors (ids int64 [nq,k], scores f32 [nq,k], counts int32 [nq]); async on the current stream.
    counts[q] == -1 marks an overflowed query (see `search_topk_exact`)."""
    require_cuda(db, "db")
    require_cuda(queries, "queries")
    assert db.dtype == torch.bfloat16 and db.dim() == 2 and db.stride(1) == 1
    assert queries.dtype == torch.float32 and queries.dim() == 2 and queries.is_contiguous()
    nq = queries.shape[0]
    assert queries.shape[1] == d and n_rows <= db.shape[0]
    if not (1 <= k <= RVO_MAX_K):
        raise RvoError(f"k={k} outside 1..{RVO_MAX_K}")
    dev = queries.device
    if out is None:
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
    else:
        ids, scores, counts = out
    lib = _lib.load()
    nbytes = lib.rvo_search_workspace_bytes(n_rows, d, nq, k)
    if nbytes == 0:
        raise RvoError(f"rvo_search_workspace_bytes rejected n_rows={n_rows} d={d} nq={nq} k={k}: "
                       + lib.rvo_last_error().decode())
    ws = workspace(dev, nbytes)
    thr = -math.inf if score_threshold is None else float(score_threshold)
    check(lib.rvo_search_topk(_ptr(db), n_rows, d, db.stride(0), _ptr(queries), nq, k, thr, int(id_offset), _ptr(ids),
                              _ptr(scores), _ptr(counts), _ptr(ws), nbytes, _stream(dev)), "rvo_search_topk")
    return ids, scores, counts


def search_topk_exact(db, n_rows, d, queries, k, score_threshold=None, id_offset=0):
    """`search_topk` plus the documented overflow protocol: queries flagged -1 by the fused path are re-run in
    batches of <= RVO_SMALL_Q through the exact fp32 scan, which cannot overflow.  Synchronises (reads counts)."""
    ids, scores, counts = search_topk(db, n_rows, d, queries, k, score_threshold, id_offset)
    bad = (counts < 0).nonzero().flatten().tolist()  # host sync; pathological inputs only take the branch
    for i in range(0, len(bad), RVO_SMALL_Q):
        sel = bad[i:i + RVO_SMALL_Q]
        sub = queries[sel].contiguous()
        a, b, c = search_topk(db, n_rows, d, sub, k, score_threshold, id_offset)
        for j, q in enumerate(sel):
            ids[q].copy_(a[j])
            scores[q].copy_(b[j])
            counts[q] = c[j]
    return ids, scores, counts


def padded_queries(nq: int, d: int) -> int:
    r = _lib.load().rvo_padded_queries(nq, d)
    if r < 0:
        check(r, "rvo_padded_queries")
    return r


def scores_dense(db: torch.Tensor, n_rows: int, d: int, queries: torch.Tensor, row_stride: int = 1,
                 n_sample: int | None = None) -> torch.Tensor:
    """Dense tensor-core score block (tests/diagnostics): [nq, n_sample] fp32."""
    require_cuda(db, "db")
    nq = queries.shape[0]
    if n_sample is None:
        n_sample = (n_rows + row_stride - 1) // row_stride
    nq_pad = padded_queries(nq, d)
    out = torch.zeros((nq_pad, n_sample), dtype=torch.float32, device=db.device)
    nbytes = nq_pad * d_pad_of(d) * 2 + 4096
    ws = workspace(db.device, nbytes)
    lib = _lib.load()
    check(lib.rvo_scores_dense(_ptr(db), n_rows, d, db.stride(0), _ptr(queries), nq, row_stride, n_sample, _ptr(out),
                               out.stride(0), _ptr(ws), nbytes, _stream(db.device)), "rvo_scores_dense")
    return out[:nq]


def merge_topk(ids: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor, k: int):
    """K3.  ids int64 [G,nq,k], scores f32 [G,nq,k], counts int32 [G,nq] -> merged (ids, scores, counts)."""
    require_cuda(ids, "ids")
    G, nq, kk = ids.shape
    assert kk == k and ids.is_contiguous() and scores.is_contiguous() and counts.is_contiguous()
    dev = ids.device
    oi = torch.empty((nq, k), dtype=torch.int64, device=dev)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=dev)
    oc = torch.empty((nq,), dtype=torch.int32, device=dev)
    check(_lib.load().rvo_merge_topk(_ptr(ids), _ptr(scores), _ptr(counts), G, nq, k, _ptr(oi), _ptr(os_), _ptr(oc),
                                     _stream(dev)), "rvo_merge_topk")
    return oi, os_, oc
