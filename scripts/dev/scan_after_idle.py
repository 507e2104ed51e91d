"""configs[1] full-shard scan timed (library events around that launch) after idle periods of different length, and in back-to-back
steps: how much of the sustained scan time is the board's power cap."""
import os, sys, time, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
n, d, nq, k = 1_000_000, 1024, 256, 100
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(n, d, q, n_plant=128, seed=1000, device=dev)
lib = _lib.load()
ps = ops.PreparedSearch(db, n, d, q, k)
for _ in range(20):
    ps()
torch.cuda.synchronize()
_lib.set_option("time_scan", 1)
for idle_ms in (0, 0.3, 1, 3, 10, 50, 200):
    ts = []
    for _ in range(9):
        for _ in range(10):          # a sustained burst first, so every sample starts from the same thermal / power state
            ps()
        torch.cuda.synchronize()
        if idle_ms:
            time.sleep(idle_ms / 1e3)
        ps()
        ts.append(float(lib.rvo_last_scan_ms()))
    print(f"idle {idle_ms:6.1f} ms before the step: scan {statistics.median(ts) * 1e3:6.1f} us (min {min(ts) * 1e3:.1f}, max {max(ts) * 1e3:.1f})", flush=True)
_lib.set_option("time_scan", 0)
