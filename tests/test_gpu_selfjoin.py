"""GPU parity for the near-duplicate self-join (BASELINE config 4) against the oracle's brute-force restatement."""
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


def _make(n, d, dup_frac, dev, seed=0):
    """Random unit rows with `dup_frac` of them replaced by near copies (cos ~0.95..0.999) of earlier rows."""
    from revers_o_b200 import ops
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
    return ops.tile_rows(x.to(torch.bfloat16))


@pytest.mark.parametrize("n,d,lo,hi", [(6000, 256, 0, 6000), (9000, 1024, 0, 9000), (9000, 1024, 2000, 5000)])
def test_selfjoin_matches_oracle(dev, n, d, lo, hi):
    from revers_o_b200 import ops
    thr = 0.95
    db = _make(n, d, 0.05, dev, seed=n)
    pairs, scores, count, over = ops.selfjoin_threshold(db, n, d, thr, lo, hi, out_cap=1 << 16)
    torch.cuda.synchronize()
    c = int(count.item())
    assert int(over.item()) == 0 and 0 < c < (1 << 16)
    got = {(int(a), int(b)): float(s) for (a, b), s in zip(pairs[:c].cpu().numpy(), scores[:c].cpu().numpy())}
    assert len(got) == c, "duplicate pairs emitted"
    dbf = ops.untile_rows(db, n, d).float().cpu().numpy()
    rp, rs = O.selfjoin_threshold(dbf, thr, db_is_normalized=True)
    ref = {(int(a), int(b)): float(s) for (a, b), s in zip(rp, rs) if lo <= a < hi}
    assert all(a < b for a, b in got)
    for key in set(got) ^ set(ref):   # may differ only within the tolerance of the threshold itself
        s = got.get(key, ref.get(key))
        assert abs(s - thr) <= 1e-3, (key, s)
    for key in set(got) & set(ref):
        assert abs(got[key] - ref[key]) <= 1e-3
    assert len(set(got) & set(ref)) >= 0.99 * len(ref)
