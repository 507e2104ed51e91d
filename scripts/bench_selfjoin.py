"""BASELINE config 4 on one GPU: self-join of N frame embeddings x 1024-d at cos >= 0.95 (5% planted near-duplicates).
Reports pairs found, seconds, and achieved TFLOP/s on the upper-triangle flops N^2 * D (the kernel scans only j >= block)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from revers_o_b200 import ops
from test_gpu_selfjoin import _make
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = 1024
db = _make(n, d, 0.05, dev, seed=5)
ops.selfjoin_threshold(db, n, d, 0.95, 0, min(n, 8192), out_cap=1 << 20)   # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pairs, scores, count, over = ops.selfjoin_threshold(db, n, d, 0.95, out_cap=1 << 22)
e1.record()
torch.cuda.synchronize()
s = e0.elapsed_time(e1) / 1e3
flops = float(n) * n * d   # 2 * (N^2/2) * D
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
print(json.dumps({"workload": f"self-join {n} x {d}, cos>=0.95", "seconds": s, "pairs": int(count.item()), "overflowed": int(over.item()),
                  "tflops_upper_triangle": flops / s / 1e12, "frac_of_sustained_bf16": flops / s / 1e12 / peaks["bf16_tflops_sustained"],
                  "rows_per_s": n / s}))
