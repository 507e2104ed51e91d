"""GPU parity tests for K1 (segmented mask pooling + L2 normalise) against the CPU oracle: pooled embeddings
within 1e-3 relative (north_star), identical kept-region counts and order."""
import numpy as np
import pytest
import torch

from oracle import reverso_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def _check(feats, masks, max_regions=0):
    from revers_o_b200 import ops
    out, counts, src, total = ops.mask_pool(feats, masks, max_regions)
    torch.cuda.synchronize()
    B, M, P = masks.shape
    emb, rc, rsrc = O.mask_pool(feats.float().cpu().numpy(), masks.cpu().numpy(), max_regions if max_regions > 0 else None)
    t = int(total.item())
    assert t == emb.shape[0] and np.array_equal(counts.cpu().numpy(), rc)
    got = out[:t].cpu().numpy()
    assert np.array_equal(src[:t].cpu().numpy(), rsrc[:, 0] * M + rsrc[:, 1])
    rel = np.abs(got - emb) / np.maximum(np.abs(emb), 1e-3)
    assert np.max(np.abs(got - emb)) < 1e-5 and rel.max() < 1e-3
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    return got


def test_golden_fixture(dev, golden):
    feats = torch.from_numpy(golden["pool_feats_bf16_bits"].astype(np.int16)).view(torch.bfloat16).to(dev)
    masks = torch.from_numpy(golden["pool_masks"]).to(dev)
    got = _check(feats, masks)
    assert np.allclose(got, golden["pool_emb"], atol=1e-5)
    from revers_o_b200 import ops
    out, counts, _, total = ops.mask_pool(feats, masks, 4)
    assert np.array_equal(counts.cpu().numpy(), golden["pool_counts_cap4"])
    assert np.allclose(out[: int(total.item())].cpu().numpy(), golden["pool_emb_cap4"], atol=1e-5)


@pytest.mark.parametrize("B,M,grid,D", [(2, 8, 24, 1024), (5, 64, 24, 1024), (3, 50, 16, 1280), (1, 1, 7, 32), (4, 33, 23, 96),
                                        (300, 3, 7, 128),     # > 148 images, masks converted from global memory (M*P % 16 != 0)
                                        (331, 16, 24, 256)])  # > 2 images per CTA on the bulk-prefetched raw-mask path
def test_random_rectangles(dev, B, M, grid, D):
    from revers_o_b200 import synth
    feats, masks = synth.make_maskpool_inputs(B, M, grid, D, seed=11 + B, device=dev, n_empty=min(2, M - 1))
    _check(feats, masks)


def test_known_answers(dev):
    """KA7 all-ones mask == token mean (core_system.py:345-346), KA8 single patch, KA9 empty dropped, cap (:363)."""
    from revers_o_b200 import ops
    g = torch.Generator().manual_seed(0)
    feats = torch.randn((1, 577, 1024), generator=g).to(torch.bfloat16).to(dev)       # 577 tokens: cls + 24*24
    masks = torch.zeros((1, 4, 577), dtype=torch.uint8, device=dev)
    masks[0, 0] = 1
    masks[0, 2, 100] = 1
    masks[0, 3, 10:20] = 1
    out, counts, src, total = ops.mask_pool(feats, masks)
    assert int(total.item()) == 3 and counts.tolist() == [3] and src[:3].tolist() == [0, 2, 3]
    f = feats.float()
    ref0 = f[0].mean(0); ref0 = ref0 / ref0.norm()
    ref1 = f[0, 100] / f[0, 100].norm()
    assert torch.allclose(out[0], ref0, atol=1e-5) and torch.allclose(out[1], ref1, atol=1e-6)
    masks = torch.ones((1, 64, 577), dtype=torch.uint8, device=dev)
    out, counts, _, total = ops.mask_pool(feats, masks, 50)
    assert counts.tolist() == [50] and int(total.item()) == 50


def test_dense_random_masks_and_all_empty_image(dev):
    g = torch.Generator().manual_seed(1)
    feats = torch.randn((3, 576, 256), generator=g).to(torch.bfloat16).to(dev)
    masks = (torch.rand((3, 20, 576), generator=g) < 0.4).to(torch.uint8)
    masks[1] = 0                                                                       # image with no region at all
    _check(feats, masks.to(dev))


def test_config2_full_shape(dev):
    """BASELINE config 2: 256 images x 64 masks x 24x24 patches x 1024-d; oracle parity on a sample of images,
    unit norms and counts on all."""
    from revers_o_b200 import ops, synth
    B, M, grid, D = 256, 64, 24, 1024
    feats, masks = synth.make_maskpool_inputs(B, M, grid, D, seed=11, device=dev)
    out, counts, src, total = ops.mask_pool(feats, masks)
    torch.cuda.synchronize()
    t = int(total.item())
    assert t == B * (M - 2) and torch.all(counts == M - 2)
    assert torch.allclose(out[:t].norm(dim=1), torch.ones(t, device=dev), atol=1e-5)
    sel = [0, 100, 255]
    emb, rc, _ = O.mask_pool(feats[sel].float().cpu().numpy(), masks[sel].cpu().numpy())
    offs = torch.cumsum(counts, 0).cpu().numpy()
    for j, b in enumerate(sel):
        lo = offs[b] - counts[b].item()
        got = out[lo: lo + rc[j]].cpu().numpy()
        ref = emb[sum(rc[:j]): sum(rc[:j]) + rc[j]]
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-3


# ---- K1 fused with ingest: embeddings straight into the tiled bf16 DB ------------------------------------------------
@pytest.mark.parametrize("B,M,grid,D,row0", [(7, 64, 24, 1024, 0), (200, 20, 24, 256, 1000), (3, 50, 16, 1280, 128),
                                             (9, 64, 24, 1280, 256), (5, 64, 24, 2048, 0)])
def test_mask_pool_to_db_matches_two_step_ingest(dev, B, M, grid, D, row0):
    """rvo_mask_pool_to_db == rvo_mask_pool followed by rvo_normalize_rows into the DB (what qdrant's COSINE upsert of the
    reference does with the region embeddings, core_system.py:608-621): same rows, at most one bf16 ulp apart."""
    from revers_o_b200 import ops, synth
    feats, masks = synth.make_maskpool_inputs(B, M, grid, D, seed=5 + B, device=dev, n_empty=2)
    cap = row0 + B * M
    db = ops.db_alloc(cap, D, dev)
    counts, src, total, f32 = ops.mask_pool_to_db(feats, masks, db, row0, want_f32=True)
    out, counts2, src2, total2 = ops.mask_pool(feats, masks)
    torch.cuda.synchronize()
    t = int(total.item())
    assert t == int(total2.item()) and torch.equal(counts, counts2) and torch.equal(src[:t], src2[:t])
    got = ops.untile_rows(db, cap, D)[row0: row0 + t].float()
    ref_db = ops.db_alloc(cap, D, dev)
    ops.normalize_rows(out[:t].contiguous(), db=ref_db, row0=row0)
    ref = ops.untile_rows(ref_db, cap, D)[row0: row0 + t].float()
    assert torch.max(torch.abs(got - ref)).item() <= 2 ** -8 * float(ref.abs().max())      # one bf16 ulp at most
    assert (got != ref).float().mean().item() < 1e-3
    assert torch.allclose(f32[:t], out[:t], atol=1e-6)
    emb, rc, _ = O.mask_pool(feats.float().cpu().numpy(), masks.cpu().numpy())
    assert np.max(np.abs(got.cpu().numpy() - O.round_to_bf16(emb))) <= 2 ** -8
    assert float(ops.untile_rows(db, cap, D)[:row0].abs().sum()) == 0.0                      # rows before row0 untouched


def test_ingest_regions_then_search(dev):
    """Batched ingest through the vector DB, then every stored region finds itself (score ~1) with its own payload."""
    from revers_o_b200 import ops, synth
    from revers_o_b200.vector_db import B200VectorDB, models
    B, M, grid, D = 12, 16, 24, 1024
    feats, masks = synth.make_maskpool_inputs(B, M, grid, D, seed=3, device=dev, n_empty=1)
    vdb = B200VectorDB(device=dev)
    vdb.recreate_collection("r", vectors_config=models.VectorParams(size=D, distance=models.Distance.COSINE))
    n1 = vdb.ingest_regions("r", feats[:5].contiguous(), masks[:5].contiguous())
    n2 = vdb.ingest_regions("r", feats[5:].contiguous(), masks[5:].contiguous(),
                            payload_fn=lambda b, m: (f"img{b + 5}-r{m}", {"image": b + 5, "region": m}))
    assert n1 == 5 * (M - 1) and n2 == 7 * (M - 1) and vdb.count("r") == n1 + n2
    out, counts, src, total = ops.mask_pool(feats, masks)
    q = out[: n1 + n2]
    ids, sc, cnt = vdb.search_batch("r", q, 1)
    assert np.all(sc[:, 0] > 0.999)
    same = ids[:, 0] == np.arange(n1 + n2)
    assert same.mean() > 0.9
    for j in np.nonzero(~same)[0]:      # two regions of an image may be the same rectangle: identical embeddings, lower id wins
        assert ids[j, 0] < j and torch.allclose(out[int(ids[j, 0])], out[int(j)], atol=1e-6)
    j = int(np.nonzero(same)[0][-1])
    hit = vdb.search("r", q[j].cpu().numpy(), limit=1)[0]
    b, m = divmod(int(src[j]), M)
    expect_id = f"img{b}-r{m}" if j >= n1 else None
    assert hit.payload == {"image": b, "region": m} and (expect_id is None or hit.id == expect_id)


@pytest.mark.parametrize("path", [0, 1])
def test_fp16_features_are_consumed_natively(dev, path):
    """SURVEY §8 H5 "bf16 or fp16": the reference's encoder is `.half()` on CUDA (core_system.py:195-196).  fp16 features go
    into the kernels unchanged (no lossy cast to bf16): parity against the oracle fed the same fp16 values, 1e-3 relative."""
    from revers_o_b200 import _lib, ops
    g = torch.Generator().manual_seed(3)
    B, M, P, D = 5, 20, 576, 1024
    feats = torch.randn((B, P, D), generator=g).to(dev).to(torch.float16)
    masks = (torch.rand((B, M, P), generator=g) < 0.2).to(torch.uint8).to(dev)
    masks[2, 7] = 0
    _lib.set_option("pool_path", path)
    try:
        out, counts, src, total = ops.mask_pool(feats, masks)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("pool_path", 0)
    emb, rc, _ = O.mask_pool(feats.float().cpu().numpy(), masks.cpu().numpy())
    t = int(total.item())
    assert t == emb.shape[0] and np.array_equal(counts.cpu().numpy(), rc)
    got = out[:t].cpu().numpy()
    assert np.max(np.abs(got - emb)) <= 1e-3 * np.max(np.abs(emb))
    # and they are NOT the bf16-rounded features' embeddings (the cast the round-1 drop-in applied)
    emb_bf16, _, _ = O.mask_pool(feats.to(torch.bfloat16).float().cpu().numpy(), masks.cpu().numpy())
    assert np.max(np.abs(got - emb)) < 0.2 * np.max(np.abs(emb_bf16 - emb))


@pytest.mark.parametrize("B,M,grid,D", [(6, 50, 24, 1280), (160, 64, 24, 1280), (4, 64, 24, 2048), (3, 33, 24, 4096)])
def test_wide_features_stay_on_the_tensor_path(dev, B, M, grid, D):
    """PE-Core-G14 width (D = 1280) at the reference's region cap (50, core_system.py:363) and at the
    benchmark's 64 masks — all D/128 slabs of all regions no longer fit the 512 TMEM columns, so an image's regions are split
    into groups (one work item each).  Same results as the oracle, ONE launch (not the four of the CUDA-core kernels), counts and
    compaction order across groups intact."""
    from revers_o_b200 import _lib, ops, synth
    feats, masks = synth.make_maskpool_inputs(B, M, grid, D, seed=B + M, device=dev, n_empty=3)
    masks[0, : M // 2] = 0                      # a whole first group empty: later groups' rows still land compacted
    l0 = _lib.kernel_launch_count()
    out, counts, src, total = ops.mask_pool(feats, masks, 50 if M == 50 else 0)
    torch.cuda.synchronize()
    assert _lib.kernel_launch_count() - l0 == 1
    sel = list(range(min(B, 6)))
    emb, rc, rsrc = O.mask_pool(feats[sel].float().cpu().numpy(), masks[sel].cpu().numpy())
    assert np.array_equal(counts[: len(sel)].cpu().numpy(), rc)
    t = int(rc.sum())
    got = out[:t].cpu().numpy()
    assert np.array_equal(src[:t].cpu().numpy(), rsrc[:, 0] * M + rsrc[:, 1])
    assert np.max(np.abs(got - emb)) < 1e-5 and np.max(np.abs(got - emb) / np.maximum(np.abs(emb), 1e-3)) < 1e-3
    tt = int(total.item())
    assert tt == int(counts.sum().item()) and torch.allclose(out[:tt].norm(dim=1), torch.ones(tt, device=dev), atol=1e-5)


def test_two_pooling_calls_on_two_streams_do_not_deadlock(dev):
    """the grid-wide rendezvous of the pooling kernel is launched cooperatively, so two calls from two Gradio worker
    threads (their own streams) cannot each hold part of the SMs and wait for the other forever."""
    from revers_o_b200 import ops, synth
    feats, masks = synth.make_maskpool_inputs(256, 64, 24, 1024, seed=2, device=dev)
    ref = ops.mask_pool(feats, masks)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    outs = []
    for rep in range(6):
        for st in (s1, s2):
            with torch.cuda.stream(st):
                # separate scratch per stream: the workspace is per calling thread, so give each stream its own call chain
                outs.append(ops.mask_pool(feats, masks, ws_key=id(st)))
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o[1], ref[1]) and int(o[3].item()) == int(ref[3].item())
        assert torch.allclose(o[0][: int(ref[3].item())], ref[0][: int(ref[3].item())], atol=1e-6)
