"""Differential fuzz of the near-duplicate self-join (configs[4] path, `ops.selfjoin_exact`): random sizes, widths, thresholds,
planted near-copies, static-scene cliques, row ranges, id offsets and deliberately small candidate / output capacities (to force the
exact fallbacks) against a brute-force fp32 restatement on the GPU.  `python scripts/dev/fuzz_selfjoin.py [cases] [seed]`."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from revers_o_b200 import ops, synth
dev = torch.device("cuda:0")
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
TOL = 1e-4
bad = 0
for c in range(cases):
    n = int(rs.choice([300, 1000, 4095, 4096, 4097, 9000, 20000, 30000]))
    d = int(rs.choice([64, 256, 1024, 1280]))
    thr = float(rs.choice([0.8, 0.9, 0.95, 0.99]))
    cluster = int(rs.choice([0, 0, 50, 400]))
    if cluster >= n // 2:
        cluster = 0
    db = synth.make_selfjoin_db(n, d, float(rs.choice([0.0, 0.05, 0.3])), dev, seed=int(rs.randint(1 << 30)), cluster=cluster)
    lo = 0 if rs.rand() < 0.5 else int(rs.randint(0, n - 1))
    hi = n if rs.rand() < 0.5 else int(rs.randint(lo + 1, n + 1))
    off = 0 if rs.rand() < 0.7 else int(rs.randint(1, 1 << 33))
    kw = {}
    if rs.rand() < 0.4:
        kw["cand_cap"] = int(rs.choice([1024, 2048]))
    if rs.rand() < 0.4:
        kw["max_pairs"] = int(rs.choice([16, 1000]))
    tag = f"case {c} n={n} d={d} thr={thr} cluster={cluster} rows [{lo},{hi}) offset={off} {kw}"
    try:
        p, s = ops.selfjoin_exact(db, n, d, thr, lo, hi, id_offset=off, **kw)
        rows = ops.untile_rows(db, n, d).float()
        got = {}
        for (a, b), sc in zip(p.tolist(), s.tolist()):
            assert (a - off, b - off) not in got, f"pair {(a, b)} emitted twice"
            got[(a - off, b - off)] = sc
        assert all(lo <= a < hi and a < b < n for a, b in got), "pair outside the requested range / not i < j"
        want, near = {}, set()
        for b0 in range(lo, hi, 2048):
            b1 = min(hi, b0 + 2048)
            sc = rows[b0:b1] @ rows.T
            ii = torch.arange(b0, b1, device=dev).view(-1, 1)
            jj = torch.arange(n, device=dev).view(1, -1)
            up = jj > ii
            m = (sc >= thr - TOL) & up
            a, b = m.nonzero(as_tuple=True)
            v = sc[m]
            for x, y, z in zip((a + b0).tolist(), b.tolist(), v.tolist()):
                if z >= thr + TOL:
                    want[(x, y)] = z
                elif z >= thr:
                    want[(x, y)] = z; near.add((x, y))
                else:
                    near.add((x, y))
        missing = [k for k in want if k not in got and k not in near]
        extra = [k for k in got if k not in want and k not in near]
        assert not missing, f"{len(missing)} pairs missing, e.g. {missing[:3]}"
        assert not extra, f"{len(extra)} pairs too many, e.g. {extra[:3]} score {got[extra[0]]}"
        worst = max((abs(got[k] - want[k]) for k in got if k in want), default=0.0)
        assert worst <= TOL, f"score differs by {worst}"
    except AssertionError as e:
        bad += 1
        print(f"MISMATCH {tag}: {e}")
    if c % 10 == 9:
        print(f"{c + 1} cases done, {bad} mismatches", flush=True)
    del db
print(f"fuzz: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
