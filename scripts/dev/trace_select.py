"""Phase time stamps of the fused last-level select kernel on configs[1] (option select_trace)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
n, d, nq, k = int(os.environ.get("N", 1_000_000)), int(os.environ.get("D", 1024)), int(os.environ.get("Q", 256)), 100
plant = int(os.environ.get("PLANT", 128))
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=plant, seed=1000, device=dev)
for _ in range(3):
    ops.search_topk(db, n, d, q, k)
tr = torch.zeros((nq, 16), dtype=torch.int64, device=dev)
_lib.set_option("select_trace", tr.data_ptr())
ops.search_topk(db, n, d, q, k)
torch.cuda.synchronize()
_lib.set_option("select_trace", 0)
t = tr.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
hot = t[:, 1] == 0          # the hot path never stamps slots 1..4
print(f"hot-path CTAs: {int(hot.sum())} of {nq}")
if hot.any():
    h = t[hot]
    names = {0: "start", 5: "cut+compact", 6: "rescore", 7: "rank", 8: "emit"}
    prev = 0
    for sl in (5, 6, 7, 8):
        dur = (h[:, sl] - h[:, prev]) / 1e3
        print(f"  hot {names[sl]:12s} median {np.median(dur):6.2f} us   max {dur.max():6.2f}")
        prev = sl
    tot = (h[:, 8] - h[:, 0]) / 1e3
    print(f"  hot CTA total median {np.median(tot):.2f} us, max {tot.max():.2f}; kernel span {(h[:, 8].max() - t0) / 1e3:.2f} us")
    print("  n_hot median %d max %d, n_act median %d max %d" % (np.median(h[:, 9]), h[:, 9].max(), np.median(h[:, 10]), h[:, 10].max()))
if (~hot).any():
    g = t[~hot]
    names = ["start", "hist", "bound", "compact", "sort1", "prefix", "rescore", "sort2", "emit"]
    ph = (g[:, :9] - t0) / 1e3
    d_ = np.diff(ph, axis=1)
    print("general path durations (us), median:", " ".join(f"{names[i + 1]}={np.median(d_[:, i]):.2f}" for i in range(8)))
    print("  C median %d, n_act median %d, nnz median %d" % (np.median(g[:, 9]), np.median(g[:, 10]), np.median(g[:, 11])))
