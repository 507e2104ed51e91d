import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from revers_o_b200 import _lib, ops, synth
from oracle import reverso_oracle as O
dev = torch.device("cuda:0")
n, d, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 100
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
for opts in sys.argv[4:] or [""]:
    for kv in filter(None, opts.split(",")):
        a, b = kv.split("="); _lib.set_option(a, int(b))
    ids, sc, cnt = ops.search_topk(db, n, d, q, k)
    torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    bad = np.nonzero(c != k)[0]
    print(f"opts[{opts}] flagged={np.sum(c < 0)} short={np.sum((c >= 0) & (c < k))} first_bad={bad[:10].tolist()} vals={c[bad[:10]].tolist()}")
    sel = list(range(0, nq, max(1, nq // 8)))
    dbf = ops.untile_rows(db, n, d).float().cpu().numpy()
    ref = O.search_batch(dbf, q[sel].cpu().numpy(), k, None, db_is_normalized=True)
    for j, qi in enumerate(sel):
        if c[qi] == k:
            ok = set(ids[qi].tolist()) == set(ref[j][0].tolist())
            print("  q", qi, "match" if ok else "MISMATCH", float(np.max(np.abs(sc[qi].cpu().numpy() - ref[j][1]))))
