"""Synthetic workloads of the BASELINE.json shapes.  Input generation only — torch ops here
produce TEST/BENCH INPUTS (random unit vectors, planted neighbours, rectangle masks); none of this is on
the product path.

Search DB: rows x ~ N(0, I_D), L2-normalised in fp32, rounded to bf16 (the DB is defined as its bf16
values).  Random unit vectors in 1024-d all score ~0 +- 0.03 against a query, so for each query `n_plant`
near neighbours `a*q + sqrt(1-a^2)*n_perp` with `a` spread over [0.5, 0.99] are planted at random rows to
make the top-k non-degenerate.
"""
from __future__ import annotations

import numpy as np
import torch


def make_queries(nq: int, d: int, seed: int = 7, device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn((nq, d), generator=g, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    # queries are handed over UN-normalised (scaled) on purpose: the path must normalise them
    scale = 0.5 + torch.rand((nq, 1), generator=g)
    return (q * scale).to(device)


def make_db(n: int, d: int, queries: torch.Tensor | None = None, n_plant: int = 128, seed: int = 1000,
            device="cpu", chunk: int = 1 << 18) -> torch.Tensor:
    """Tiled bf16 DB storage (ops.db_alloc layout) of n L2-normalised rows, generated on `device` chunk by chunk.
    `ops.untile_rows(db, n, d)` gives the row-major view for checkers."""
    from . import ops
    dev = torch.device(device)
    tiled = ops.db_alloc(n, d, dev)
    nk = tiled.shape[1]
    d_pad = nk * 64
    chunk = max(128, chunk // 128 * 128)
    g = torch.Generator(device=dev).manual_seed(seed)
    # chunk by chunk straight into the tiled storage [block][k-chunk][128][64]: peak extra memory is one chunk
    # (a 128 GB shard of the 100M x 1280 DB at 2 GPUs leaves no room for a row-major alias)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        x = torch.randn((hi - lo, d), generator=g, dtype=torch.float32, device=dev)
        x = x / x.norm(dim=1, keepdim=True)
        nb_c = (hi - lo + 127) // 128
        pad = torch.zeros((nb_c * 128, d_pad), dtype=torch.bfloat16, device=dev)
        pad[: hi - lo, :d] = x.to(torch.bfloat16)
        del x
        tiled[lo // 128: lo // 128 + nb_c].copy_(pad.view(nb_c, 128, nk, 64).permute(0, 2, 1, 3))
        del pad
    if queries is not None and n_plant > 0 and n >= queries.shape[0]:   # fewer rows than queries: nothing to plant
        nq = queries.shape[0]
        qn = (queries.float() / queries.float().norm(dim=1, keepdim=True)).to(dev)
        gp = torch.Generator(device="cpu").manual_seed(seed + 1)
        per = min(n_plant, max(1, n // max(nq, 1)))
        if n <= (1 << 24):
            rows = torch.randperm(n, generator=gp)[: nq * per]
        else:  # big shards: distinct rows without materialising a permutation of n
            rows = torch.randint(0, n, (2 * nq * per,), generator=gp).unique()
            rows = rows[torch.randperm(rows.numel(), generator=gp)][: nq * per]
        rows = rows.view(nq, per).to(dev)
        alpha = torch.linspace(0.5, 0.99, per, device=dev).view(1, per, 1)
        for lo in range(0, nq, 64):
            hi = min(nq, lo + 64)
            noise = torch.randn((hi - lo, per, d), generator=g, dtype=torch.float32, device=dev)
            qq = qn[lo:hi].unsqueeze(1)
            noise = noise - (noise * qq).sum(-1, keepdim=True) * qq
            noise = noise / noise.norm(dim=-1, keepdim=True)
            v = alpha * qq + torch.sqrt(1 - alpha * alpha) * noise
            v = v / v.norm(dim=-1, keepdim=True)
            vb = torch.zeros(((hi - lo) * per, d_pad), dtype=torch.bfloat16, device=dev)
            vb[:, :d] = v.reshape(-1, d).to(torch.bfloat16)
            r = rows[lo:hi].reshape(-1)
            tiled[r // 128, :, r % 128, :] = vb.view(-1, nk, 64)
    return tiled


def make_maskpool_inputs(B: int, M: int, grid: int, D: int, seed: int = 11, device="cpu", n_empty: int = 2):
    """cfg3: feats [B, grid*grid, D] ~ N(0,1) bf16 (stand-in for random-init PE patch features), masks
    [B, M, grid*grid] uint8 random axis-aligned rectangles with area fraction ~ U(0.02, 0.30); `n_empty`
    regions per image are forced empty to exercise the skip (core_system.py:402-404)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    P = grid * grid
    feats = torch.randn((B, P, D), generator=g, dtype=torch.float32, device=dev).to(torch.bfloat16)
    rs = np.random.RandomState(seed)
    masks = np.zeros((B, M, grid, grid), dtype=np.uint8)
    frac = rs.uniform(0.02, 0.30, size=(B, M))
    aspect = rs.uniform(0.5, 2.0, size=(B, M))
    h = np.clip(np.round(np.sqrt(frac * P * aspect)), 1, grid).astype(int)
    w = np.clip(np.round(frac * P / h), 1, grid).astype(int)
    y0 = (rs.uniform(size=(B, M)) * (grid - h + 1)).astype(int)
    x0 = (rs.uniform(size=(B, M)) * (grid - w + 1)).astype(int)
    for b in range(B):
        empty = set(rs.choice(M, size=min(n_empty, M), replace=False).tolist()) if n_empty else set()
        for m in range(M):
            if m in empty:
                continue
            masks[b, m, y0[b, m]: y0[b, m] + h[b, m], x0[b, m]: x0[b, m] + w[b, m]] = 1
    return feats, torch.from_numpy(masks.reshape(B, M, P)).to(dev)


def make_selfjoin_db(n: int, d: int, dup_frac: float, device, seed: int = 0, cluster: int = 0) -> torch.Tensor:
    """configs[4] input: random unit rows with `dup_frac` of them replaced by near copies (cos ~0.93..0.999) of earlier rows
    (video keyframes of slowly changing scenes); `cluster` > 0 additionally makes that many rows exact copies of row 0 (a static
    scene: far more near-duplicates of one frame than a candidate list holds).  Tiled bf16 storage."""
    from . import ops
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n, d), generator=g, device=dev)
    x = x / x.norm(dim=1, keepdim=True)
    nd = int(n * dup_frac)
    src = torch.randint(0, n // 2, (nd,), generator=g, device=dev)
    dst = n // 2 + torch.randperm(n - n // 2, generator=g, device=dev)[:nd]
    a = torch.empty(nd, device=dev).uniform_(0.93, 0.999, generator=g).view(-1, 1)
    noise = torch.randn((nd, d), generator=g, device=dev)
    noise = noise - (noise * x[src]).sum(1, keepdim=True) * x[src]
    noise = noise / noise.norm(dim=1, keepdim=True)
    x[dst] = a * x[src] + torch.sqrt(1 - a * a) * noise
    if cluster > 0:
        rows = torch.randperm(n - 1, generator=g, device=dev)[:cluster] + 1
        x[rows] = x[0].clone()
    return ops.tile_rows(x.to(torch.bfloat16))


def bf16_to_f32_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()
