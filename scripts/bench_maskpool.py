"""K1 timing on BASELINE config 2 (256 images x 64 masks x 24x24 patches x 1024-d): images/s and achieved
algorithmic GB/s (feats bf16 + masks u8 + out fp32 = 378.5 MB per batch) against the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from revers_o_b200 import _lib, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
B, M, G, D = 256, 64, 24, 1024
feats, masks = synth.make_maskpool_inputs(B, M, G, D, seed=11, device=dev)
for _ in range(5):
    ops.mask_pool(feats, masks)
torch.cuda.synchronize()
# inputs (311 MB) are larger than L2 (126 MB), so back-to-back calls stream from HBM every time; timing N calls between
# two events keeps the CPU's launch latency out of the measurement (a single 80 us call would expose it)
N = 20
l0 = _lib.kernel_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    out, counts, src, total = ops.mask_pool(feats, masks)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
launches = (_lib.kernel_launch_count() - l0) // N
alg = B * G * G * D * 2 + B * M * G * G + int(total.item()) * D * 4
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
print(json.dumps({"kernel": f"mask_pool ({launches} launches)", "ms": ms, "images_per_s": B / ms * 1e3, "regions": int(total.item()),
                  "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "hbm_frac": alg / ms / 1e6 / peaks["hbm_gbs"]}))
# PE-Core-G14 width at the benchmark's 64 masks: two region groups per image on the tensor path (was the CUDA-core kernels)
for (B2, M2, D2) in ((256, 64, 1280), (256, 50, 1280)):
    f2, m2 = synth.make_maskpool_inputs(B2, M2, G, D2, seed=12, device=dev)
    bufs = ops.mask_pool(f2, m2)
    for _ in range(3):
        ops.mask_pool(f2, m2, out=bufs)
    torch.cuda.synchronize()
    l0 = _lib.kernel_launch_count()
    e0.record()
    for _ in range(N):
        ops.mask_pool(f2, m2, out=bufs)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / N
    alg2 = B2 * G * G * D2 * 2 + B2 * M2 * G * G + int(bufs[3].item()) * D2 * 4
    print(json.dumps({"shape": f"{B2} x {M2} masks x {G * G} patches x {D2}", "launches": (_lib.kernel_launch_count() - l0) // N, "ms": ms2,
                      "images_per_s": B2 / ms2 * 1e3, "achieved_gbs": alg2 / ms2 / 1e6, "hbm_frac": alg2 / ms2 / 1e6 / peaks["hbm_gbs"]}))
