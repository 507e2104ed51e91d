"""Per-image pipeline time stamps of the tensor-core mask-pool kernel (option pool_trace): where does a CTA's time go?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revers_o_b200 import _lib, ops, synth
dev = torch.device("cuda:0")
B, M, G, D = 256, 64, 24, 1024
feats, masks = synth.make_maskpool_inputs(B, M, G, D, seed=11, device=dev)
for _ in range(3):
    ops.mask_pool(feats, masks)
tr = torch.zeros((B, 8), dtype=torch.int64, device=dev)
_lib.set_option("pool_trace", tr.data_ptr())
ops.mask_pool(feats, masks)
torch.cuda.synchronize()
_lib.set_option("pool_trace", 0)
t = tr.cpu().numpy().astype(np.float64)
t0 = t[t > 0].min()
t = (t - t0) / 1e3
names = ["bfull0", "tempty0", "mma_done", "tfull0", "tfull1", "pass1", "epi_done", "cvt_start"]
print("image  " + " ".join(f"{n:>9}" for n in names))
for b in (0, 1, 50, 107, 108, 147, 148, 149, 198, 255):
    print(f"{b:5d}  " + " ".join(f"{x:9.2f}" for x in t[b]))
print("end of kernel (max stamp): %.2f us" % t.max())
d = lambda a, b: np.median(t[:, b] - t[:, a])
print("median  mma phase %.2f  tfull1->pass1 done %.2f  pass2 %.2f" % (d(1, 2), d(4, 5), d(5, 6)))
print("round1 mma phase median %.2f, round2 %.2f" % (np.median((t[:, 2] - t[:, 1])[:148]), np.median((t[:, 2] - t[:, 1])[148:])))
