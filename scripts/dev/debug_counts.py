import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
n, d, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 100
nlev = int(sys.argv[4])  # number of FILTER levels expected (for offsets)
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
ids, sc, cnt = ops.search_topk(db, n, d, q, k)
torch.cuda.synchronize()
lib = _lib.load()
nbytes = lib.rvo_search_workspace_bytes(n, d, nq, k)
ws = ops.workspace(dev, nbytes)
nq_pad = lib.rvo_padded_queries(nq, d); d_pad = ops.d_pad_of(d)
al = lambda x, a: (x + a - 1) // a * a
off = 0
off_qb = al(off, 1024); off = off_qb + nq_pad * d_pad * 2
off_cnt = al(off, 256); off = off_cnt + (nlev + 1) * nq_pad * 16 * 4
off_qn = al(off, 1024); off = off_qn + nq * d_pad * 4
off_tau = al(off, 256); off = off_tau + (nlev + 1) * nq_pad * 4
w = ws.cpu().numpy()
c = w[off_cnt: off_cnt + (nlev + 1) * nq_pad * 16 * 4].view(np.int32).reshape(nlev + 1, nq_pad, 16)
t = w[off_tau: off_tau + (nlev + 1) * nq_pad * 4].view(np.float32).reshape(nlev + 1, nq_pad)
print("flagged", int((cnt < 0).sum()))
for L in range(nlev):
    tot = c[L].sum(1)
    print(f"level {L}: total/query mean {tot[:nq].mean():.0f} max {tot[:nq].max()} ; max sublist {c[L][:nq].max()} ; per-sublist mean {c[L][:nq].mean(0).round().tolist()}")
    print("   tau mean %.4f min %.4f max %.4f  pad %s" % (t[L][:nq].mean(), t[L][:nq].min(), t[L][:nq].max(), t[L][nq:nq+2].tolist()))
    for qb in range(0, nq, 256):
        print("   qblock", qb // 256, "tot mean", tot[qb:qb+256].mean().round(), "tau mean", t[L][qb:qb+256].mean().round(4))
