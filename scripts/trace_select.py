"""Phase time stamps of the fused last-level select kernel on configs[1] (option select_trace)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
n, d, nq, k = int(os.environ.get("N", 1_000_000)), 1024, 256, 100
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
for _ in range(3):
    ops.search_topk(db, n, d, q, k)
tr = torch.zeros((nq, 16), dtype=torch.int64, device=dev)
_lib.set_option("select_trace", tr.data_ptr())
ops.search_topk(db, n, d, q, k)
torch.cuda.synchronize()
_lib.set_option("select_trace", 0)
t = tr.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
names = ["start", "hist", "bound", "compact", "sort1", "prefix", "rescore", "sort2", "emit"]
ph = (t[:, :9] - t0) / 1e3
print("phase ends (us since first CTA start), median over CTAs / max over CTAs")
for i, nme in enumerate(names):
    print(f"  {nme:8s} {np.median(ph[:, i]):7.2f} {ph[:, i].max():7.2f}")
d_ = np.diff(ph, axis=1)
print("phase durations (us), median:", " ".join(f"{names[i + 1]}={np.median(d_[:, i]):.2f}" for i in range(8)))
print("C (compacted) median %d, n_act median %d, nnz median %d" % (np.median(t[:, 9]), np.median(t[:, 10]), np.median(t[:, 11])))
