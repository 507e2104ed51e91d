"""Per-step CUDA-event times of the north-star sequence (Q = 4096 -> 64 -> 16 on one 12.5M x 1280 shard)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revers_o_b200 import synth, _lib
from revers_o_b200.sharded import ShardedIndex
dev = torch.device("cuda:0")
n, d = int(os.environ.get("N", 12_500_000)), 1280
q_all = synth.make_queries(4096, d, seed=7, device=dev)
db = synth.make_db(n, d, q_all, n_plant=16, seed=2000, device=dev)
idx = ShardedIndex(db, n, d, 0)
for nq in [int(x) for x in os.environ.get("QS", "4096,64,16,16").split(",")]:
    q = q_all[:nq].contiguous()
    for _ in range(3):
        idx.search(q, 100)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    l0 = _lib.kernel_launch_count()
    evs[0].record()
    for i in range(10):
        out = idx.search(q, 100)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(10)]
    print(nq, "launches/step", (_lib.kernel_launch_count() - l0) / 10, "ms:", " ".join(f"{t:.2f}" for t in ts),
          "bad", int((out[2] != 100).sum()), flush=True)
