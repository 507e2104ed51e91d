"""Differential fuzz of K2: random shapes, thresholds, tie structures and paths against an fp32 restatement on the GPU (torch matmul,
checker code).  Tighter than the north_star tolerance: the product re-scores in fp32, so scores must agree to ~1e-5 and id sets may
differ only among ties within that.  `python scripts/dev/fuzz_search.py [cases] [seed]`; exits 1 on the first mismatch."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import assert_topk_match
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
TOL = 1e-5


def reference(db, n, d, q, k, thr):
    rows = ops.untile_rows(db, n, d).float()
    qn = q.float() / q.float().norm(dim=1, keepdim=True)
    out = []
    for lo in range(0, q.shape[0], 64):
        sc = qn[lo:lo + 64] @ rows.T
        kk = min(n, k + 8)
        v, i = torch.topk(sc, kk, dim=1)
        v, i = v.cpu().numpy(), i.cpu().numpy()
        for r in range(v.shape[0]):
            order = np.lexsort((i[r], -v[r]))            # (score desc, id asc)
            vv, ii = v[r][order], i[r][order]
            if thr is not None:
                keep = vv >= thr
                vv, ii = vv[keep], ii[keep]
            out.append((ii[:k], vv[:k]))
    return out


def make_case(c):
    kind = rs.choice(["random", "ties", "adjacent", "boundary"])
    d = int(rs.choice([64, 96, 128, 256, 1024, 1280, 2048]))
    nq = int(rs.choice([1, 2, 4, 5, 16, 37, 64, 130, 256, 300]))
    k = int(rs.choice([1, 7, 10, 100, 128, 129, 512]))
    if kind == "boundary":
        n = int(rs.choice([16384, 16385, 16640, 32768, 131072, 131073, 65535, 255, 256, 257, 128])) + int(rs.randint(0, 3))
    else:
        n = int(rs.randint(200, 250_000 if d <= 1024 else 120_000))
    k = min(k, 512)
    thr = None if rs.rand() < 0.5 else float(rs.choice([0.0, 0.03, 0.08, 0.5]))
    q = synth.make_queries(nq, d, seed=int(rs.randint(1 << 30)), device=dev)
    db = synth.make_db(n, d, q, n_plant=int(rs.choice([0, 8, 128])), seed=int(rs.randint(1 << 30)), device=dev)
    if kind in ("ties", "adjacent") and n >= 1000:
        rows = ops.untile_rows(db, n, d)
        if kind == "ties":          # many exact duplicates of a few rows scattered over the shard (equal scores, ids decide)
            src = torch.from_numpy(rs.randint(0, n, 8)).to(dev)
            dst = torch.from_numpy(rs.choice(n, size=min(n // 2, 4000), replace=False)).to(dev)
            rows[dst] = rows[src[torch.arange(dst.numel(), device=dev) % 8]]
        else:                       # a run of adjacent near-copies of the first query (video frames): stresses the grouped seed maxima
            a = int(rs.randint(0, n - 600))
            qn = (q[0] / q[0].norm()).to(torch.bfloat16)
            rows[a:a + 600] = qn
            rows[a:a + 600, : min(8, d)] += (torch.arange(600, device=dev).view(-1, 1) % 7 * 0.01).to(torch.bfloat16)
            rr = rows[a:a + 600].float()
            rows[a:a + 600] = (rr / rr.norm(dim=1, keepdim=True)).to(torch.bfloat16)
        db = ops.tile_rows(rows.contiguous())
    return kind, n, d, nq, k, thr, q, db


bad = 0
for c in range(cases):
    kind, n, d, nq, k, thr, q, db = make_case(c)
    ref = reference(db, n, d, q, k, thr)
    paths = [("auto", 0)]
    if nq <= 4 and ops.d_pad_of(d) <= 2048:
        paths += [("tensor", _lib.RVO_PATH_TENSOR), ("dense", _lib.RVO_PATH_DENSE)]
    elif nq <= 4:
        pass
    for pname, path in paths:
        try:
            if path == 0:
                ids, sc, cnt = ops.search_topk_exact(db, n, d, q, k, thr)
            else:
                ids, sc, cnt = ops.search_topk(db, n, d, q, k, thr, path=path)
            torch.cuda.synchronize()
            cn = cnt.cpu().numpy()
            if (cn < 0).any():
                if path == 0:
                    raise AssertionError("exact wrapper left an overflow flag")
                print(f"case {c} [{kind}] n={n} d={d} nq={nq} k={k} thr={thr} path={pname}: {int((cn < 0).sum())} queries flagged (allowed on a forced path)")
                continue
            assert_topk_match(ids.cpu().numpy(), sc.cpu().numpy(), cn, ref, k, tol=TOL, name=f"case {c}")
        except AssertionError as e:
            bad += 1
            print(f"MISMATCH case {c} [{kind}] n={n} d={d} nq={nq} k={k} thr={thr} path={pname}: {e}")
    if c % 10 == 9:
        print(f"{c + 1} cases done, {bad} mismatches", flush=True)
    del db, q
print(f"fuzz: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
