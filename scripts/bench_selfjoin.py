"""BASELINE configs[4]: near-duplicate self-join of N frame embeddings x 1024-d at cos >= 0.95 (5 % planted near-duplicates).
One GPU: `python scripts/bench_selfjoin.py [N]`.  N GPUs: launch with torchrun — the DB is replicated (4.1 GB), the 4096-row query
blocks are dealt in snake order (sharded.selfjoin_blocks), no collective on the data path; time = max over ranks.
Reports pairs found, seconds, and achieved TFLOP/s on the upper-triangle flops N^2 * D (each block scans only rows >= itself)."""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from revers_o_b200 import ops
from revers_o_b200.sharded import selfjoin_blocks
from revers_o_b200 import synth
_make = lambda n, d, f, dev, seed=0: synth.make_selfjoin_db(n, d, f, dev, seed=seed)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = 1024
db = _make(n, d, 0.05, dev, seed=5)                      # same seed on every rank: replicated DB
ops.selfjoin_threshold(db, n, d, 0.95, 0, min(n, 8192), out_cap=1 << 20)   # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
blocks = selfjoin_blocks(n, world, rank)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
ovf = torch.zeros(1, dtype=torch.int64, device=dev)
e0.record()
for lo, hi in blocks:
    pairs, scores, count, over = ops.selfjoin_threshold(db, n, d, 0.95, lo, hi, out_cap=1 << 22)
    cnt += count
    ovf += over
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt)
    dist.all_reduce(ovf)
s = float(t.item())
flops = float(n) * n * d   # 2 * (N^2/2) * D
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"bf16_tflops": 1590.0}
if rank == 0:
    print(json.dumps({"workload": f"self-join {n} x {d}, cos>=0.95, {world} GPU(s)", "seconds": s, "pairs": int(cnt.item()),
                      "overflowed": int(ovf.item()), "tflops_upper_triangle_total": flops / s / 1e12,
                      "frac_of_burst_bf16_per_gpu": flops / s / 1e12 / world / peaks["bf16_tflops"], "rows_per_s": n / s}))
if world > 1:
    dist.destroy_process_group()
