"""Phase time stamps of dense_topk_kernel on configs[0] (option select_trace)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
n, d, nq, k = int(os.environ.get("N", 10_000)), 1024, int(os.environ.get("Q", 1)), int(os.environ.get("K", 10))
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=32, seed=1000, device=dev)
for _ in range(5):
    ops.search_topk(db, n, d, q, k)
tr = torch.zeros((nq, 16), dtype=torch.int64, device=dev)
_lib.set_option("select_trace", tr.data_ptr())
for rep in range(3):
    tr.zero_()
    ops.search_topk(db, n, d, q, k)
    torch.cuda.synchronize()
    t = tr.cpu().numpy().astype(np.float64)
    names = ["load+range", "rounds", "compact", "sort", "emit"]
    print("phase us:", " ".join(f"{names[i]}={(t[0, i + 1] - t[0, i]) / 1e3:.2f}" for i in range(5)), "total", (t[0, 5] - t[0, 0]) / 1e3, "C", t[0, 9])
_lib.set_option("select_trace", 0)
