"""In-stream timeline of the search chain on configs[1]: every kernel records the earliest start / latest end of its CTAs on the
GPU's global timer (options chain_trace + select_trace), for several back-to-back steps — kernel durations AND the gaps between
them, which neither CUDA events around one launch nor ncu's isolated replays show."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
n, d, nq, k = int(os.environ.get("N", 1_000_000)), int(os.environ.get("D", 1024)), int(os.environ.get("Q", 256)), int(os.environ.get("K", 100))
steps = int(os.environ.get("STEPS", 6))
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
for kv in filter(None, os.environ.get("RVO_OPTS", "").split(",")):
    _lib.set_option(kv.split("=")[0].strip(), int(kv.split("=")[1]))
ps = ops.PreparedSearch(db, n, d, q, k)
for _ in range(5):
    ps()
torch.cuda.synchronize()
U64 = np.iinfo(np.uint64).max
init = torch.from_numpy(np.tile(np.array([U64, 0], dtype=np.uint64), 8).view(np.int64)).to(dev)
tr = [init.clone() for _ in range(steps)]
st = [torch.zeros((nq, 16), dtype=torch.int64, device=dev) for _ in range(steps)]
torch.cuda.synchronize()
for s in range(steps):
    _lib.set_option("chain_trace", tr[s].data_ptr())
    _lib.set_option("select_trace", st[s].data_ptr())
    ps()
_lib.set_option("chain_trace", 0)
_lib.set_option("select_trace", 0)
torch.cuda.synchronize()
names = ["normalise", "seed scan", "seed tau", "scan", "select"]
rows = []
for s in range(steps):
    t = tr[s].cpu().numpy().view(np.uint64).reshape(8, 2).astype(np.float64)
    sel = st[s].cpu().numpy().view(np.uint64).astype(np.float64)
    t[4, 0] = sel[:, 0].min()
    t[4, 1] = sel[:, 1:9].max()
    rows.append(t[:5])
t0 = rows[0][0, 0]
prev_end = None
for s, t in enumerate(rows):
    line = []
    for i, nm in enumerate(names):
        gap = (t[i, 0] - prev_end) / 1e3 if prev_end is not None else float("nan")
        line.append(f"{nm}: gap {gap:5.1f} run {(t[i, 1] - t[i, 0]) / 1e3:6.1f}")
        prev_end = t[i, 1]
    print(f"step {s}: start {(t[0, 0] - t0) / 1e3:8.1f} us | " + " | ".join(line))
tot = [(rows[s + 1][0, 0] - rows[s][0, 0]) / 1e3 for s in range(steps - 1)]
print("step period (us):", " ".join(f"{x:.1f}" for x in tot))
