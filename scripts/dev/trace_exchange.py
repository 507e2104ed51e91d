"""Where the exchange + merge time of the row-sharded search goes (torchrun, one rank per GPU): per rank, per step
   K2 (scan .. select + push) duration, merge wait for the peers' flags, merge itself; and the step time with / without exchange."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from revers_o_b200 import _lib, ops, synth
from revers_o_b200.sharded import ShardedIndex, shard_bounds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, nq, k = 1_000_000, 1024, 256, 100
lo, hi = shard_bounds(n, world, rank)
q = synth.make_queries(nq, d, seed=7, device=dev)
db = synth.make_db(hi - lo, d, q, n_plant=max(1, 128 // world), seed=1000 + rank, device=dev)
idx = ShardedIndex(db, hi - lo, d, lo)
assert idx.enable_peer_exchange(nq, k)
tr = torch.zeros(3, dtype=torch.int64, device=dev)
for _ in range(10):
    idx.search(q, k)
torch.cuda.synchronize(); dist.barrier()
_lib.set_option("merge_trace", tr.data_ptr())
rows = []
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for step in range(30):
    ev[0].record()
    t = idx.submit(q, k)
    ev[1].record()
    idx.collect(t)
    ev[2].record()
    torch.cuda.synchronize()
    a = tr.cpu().numpy().astype(np.float64)
    rows.append((ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3, (a[1] - a[0]) / 1e3, (a[2] - a[1]) / 1e3))
_lib.set_option("merge_trace", 0)
r = np.array(rows[5:])
out = [None] * world
dist.all_gather_object(out, r.mean(0).tolist())
if rank == 0:
    print("per rank, us (sync after every step): K2+push | merge launch..end | of which waiting for flags | merge after flags")
    for g, v in enumerate(out):
        print(f"  rank {g}: {v[0]:7.1f} {v[1]:7.1f} {v[2]:7.1f} {v[3]:7.1f}")
# back-to-back steps (no host sync): with and without exchange
def loop(fn, steps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3], device=dev, dtype=torch.float64)
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return float(mx), float(mn)
a = loop(lambda: idx.search_local(q, k))
b = loop(lambda: idx.search(q, k))
lib = _lib.load()
import math
x = idx._xchg
def push_only():
    nbytes = lib.rvo_search_workspace_bytes(idx.n_local, d, nq, k); ws = ops.workspace(dev, nbytes)
    x["epoch"] += 1
    lib.rvo_search_topk_push(db.data_ptr(), idx.n_local, d, ops.d_pad_of(d), q.data_ptr(), nq, k, -math.inf, lo, x["table"], world, rank,
                             x["nq_max"], x["k_max"], x["epoch"], ws.data_ptr(), nbytes, torch.cuda.current_stream(dev).cuda_stream)
c = loop(push_only)
if rank == 0:
    print(f"back-to-back us/step (max, min over ranks): local K2 {a[0]:.1f} {a[1]:.1f} | K2 with push, no merge {c[0]:.1f} {c[1]:.1f} | full search {b[0]:.1f} {b[1]:.1f}")
idx.disable_peer_exchange()
dist.destroy_process_group()
