"""Differential fuzz of K1 (mask pooling, with and without the fused ingest): random batch sizes, region counts, grids (with and
without a class token), feature widths on and off the tensor path, bf16 / fp16 features, region caps, empty regions — against an
fp32 restatement on the GPU (torch einsum, checker code).  `python scripts/dev/fuzz_maskpool.py [cases] [seed]`; exit 1 on mismatch."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from revers_o_b200 import ops
dev = torch.device("cuda:0")
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
torch.backends.cuda.matmul.allow_tf32 = False


def make(c):
    D = int(rs.choice([32, 96, 128, 256, 512, 768, 1024, 1152, 1280, 1536, 2048, 4096]))
    grid = int(rs.choice([7, 8, 12, 16, 23, 24, 27, 32]))
    P = grid * grid + int(rs.rand() < 0.3)                      # + class token
    M = int(rs.choice([1, 2, 3, 8, 16, 17, 33, 48, 49, 50, 64]))
    B = int(rs.choice([1, 2, 3, 7, 40, 148, 149, 300]))
    while B * P * D > 1.5e8 and B > 1:
        B = max(1, B // 2)
    dtype = torch.float16 if rs.rand() < 0.4 else torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(int(rs.randint(1 << 30)))
    feats = torch.randn((B, P, D), generator=g, device=dev).to(dtype)
    kind = rs.choice(["rect", "dense", "sparse"])
    if kind == "rect":
        m = np.zeros((B, M, P), dtype=np.uint8)
        gg = m[:, :, : grid * grid].reshape(B, M, grid, grid)
        h = rs.randint(1, grid + 1, (B, M)); w = rs.randint(1, grid + 1, (B, M))
        y = (rs.rand(B, M) * (grid - h + 1)).astype(int); x = (rs.rand(B, M) * (grid - w + 1)).astype(int)
        for b in range(B):
            for r in range(M):
                gg[b, r, y[b, r]: y[b, r] + h[b, r], x[b, r]: x[b, r] + w[b, r]] = rs.choice([1, 255])   # any non-zero byte counts
        masks = torch.from_numpy(m)
    else:
        p = 0.5 if kind == "dense" else 0.01
        masks = torch.from_numpy((rs.rand(B, M, P) < p).astype(np.uint8))
    empty = rs.rand(B, M) < rs.choice([0.0, 0.1, 0.6])
    masks[torch.from_numpy(empty)] = 0
    if rs.rand() < 0.1:
        masks[rs.randint(B)] = 0                                  # an image without any region
    cap = 0 if rs.rand() < 0.6 else int(rs.randint(1, M + 1))
    return feats, masks.to(dev), cap, kind


def reference(feats, masks, cap):
    B, M, P = masks.shape
    lim = M if cap <= 0 else min(M, cap)
    w = (masks[:, :lim] != 0).float()
    area = w.sum(-1)
    pooled = torch.einsum("bmp,bpd->bmd", w, feats.float()) / area.clamp(min=1).unsqueeze(-1)
    nrm = pooled.norm(dim=-1, keepdim=True)
    emb = torch.where(nrm > 0, pooled / nrm.clamp(min=1e-30), torch.zeros_like(pooled))
    keep = area > 0
    b_idx, m_idx = keep.nonzero(as_tuple=True)
    return emb[keep], keep.sum(1).to(torch.int32), (b_idx * M + m_idx).to(torch.int32)


bad = 0
for c in range(cases):
    feats, masks, cap, kind = make(c)
    B, M, P = masks.shape
    D = feats.shape[2]
    tag = f"case {c} [{kind}] B={B} M={M} P={P} D={D} {str(feats.dtype)[6:]} cap={cap}"
    try:
        emb, cnt, src = reference(feats, masks, cap)
        out, counts, gsrc, total = ops.mask_pool(feats, masks, cap)
        torch.cuda.synchronize()
        t = int(total.item())
        assert t == emb.shape[0], f"total {t} vs {emb.shape[0]}"
        assert torch.equal(counts, cnt), "counts differ"
        assert torch.equal(gsrc[:t], src), "source indices differ"
        if t:
            err = (out[:t] - emb).abs().max().item()
            assert err < 2e-5, f"max abs err {err}"
        if rs.rand() < 0.5:      # fused ingest: same rows, rounded to bf16, at row0 of a tiled DB
            row0 = int(rs.choice([0, 1, 127, 128, 1000]))
            db = ops.db_alloc(row0 + B * M + 256, D, dev)
            db.zero_()
            c2, s2, t2, f32 = ops.mask_pool_to_db(feats, masks, db, row0, cap, want_f32=True)
            torch.cuda.synchronize()
            assert int(t2.item()) == t and torch.equal(c2, cnt) and torch.equal(s2[:t], src), "fused ingest bookkeeping differs"
            if t:
                rows = ops.untile_rows(db, row0 + t, D)[row0:].float()
                err = (rows - emb).abs().max().item()
                assert err < 2e-5 + 2 ** -8 * emb.abs().max().item(), f"fused ingest rows differ by {err}"
                err = (f32[:t] - emb).abs().max().item()
                assert err < 2e-5, f"fused ingest f32 rows differ by {err}"
                if row0:
                    assert ops.untile_rows(db, row0, D).abs().max().item() == 0, "rows before row0 were written"
    except AssertionError as e:
        bad += 1
        print(f"MISMATCH {tag}: {e}")
    if c % 10 == 9:
        print(f"{c + 1} cases done, {bad} mismatches", flush=True)
    del feats, masks
print(f"fuzz: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
