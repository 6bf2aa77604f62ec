"""Event trace of the tcgen05 filter kernel (CTA 0, first 256 units): where does a unit's time go?
    DSIR_TC_DEBUG=2 python tools/trace_filter.py [--batch 32]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import _lib as L, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--n", type=int, default=16384)
a = ap.parse_args()
dev = torch.device("cuda:0")
b = {k: v.to(dev) for k, v in synth.make_batch(a.batch, a.n, 64, "kitti", config=2).items()}
fs, fr = b["feat_src"], b["feat_ref"]
B, C, J = fs.shape
K = fr.shape[2]
lib = D.lib()
(f1, k1), (f2, k2) = L.feat_cn(fs), L.feat_cn(fr)
idx = torch.empty(B, J, dtype=torch.int64, device=dev)
ws = L.workspace(lib.dsir_match_argmin_workspace_bytes(B, C, J, K, D.MATCH_TC), dev)
for _ in range(3):
    L.check(lib.dsir_match_argmin(f1, f2, B, C, J, K, idx.data_ptr(), None, ws.data_ptr(), ws.numel(), D.MATCH_TC,
                                  L.stream_ptr(dev)), "match")
torch.cuda.synchronize()
out = (ctypes.c_uint32 * 4096)()
L.check(lib.dsir_match_argmin_filter_trace(ws.data_ptr(), ws.numel(), B, C, J, K, ctypes.addressof(out), L.stream_ptr(dev)), "trace")
t = np.frombuffer(out, dtype=np.uint32).astype(np.int64)
mma = t[:2048].reshape(256, 4, 2)      # [useq][a][free seen, issued]
epi = t[2048:].reshape(4, 256, 2)      # [a][useq][full seen, drained]
t0 = mma[0, 0, 0]
d = lambda x: (x - t0) & 0xffffffff
sl = slice(40, 200)                    # steady state, inside the first item (128 units) and into the second
per_unit = np.diff(d(mma[sl, 0, 0])).mean()
print(f"period per unit (MMA thread, accumulator 0): {per_unit:.0f} clk")
for acc in range(4):                   # acc = 2 * row block + column half
    r = acc >> 1
    free_seen, issued = d(mma[sl, r, 0]), d(mma[sl, r, 1])
    full_seen, drained = d(epi[acc, sl, 0]), d(epi[acc, sl, 1])
    print(f"rb {r} half {acc & 1}: issue {np.mean(issued - free_seen):6.0f} | issued->full seen {np.mean(full_seen - issued):6.0f} | "
          f"drain {np.mean(drained - full_seen):6.0f} | drained->free seen(same stage, 2 units later) "
          f"{np.mean(d(mma[sl, r, 0])[2:] - drained[:-2]):6.0f}")
print("first units of acc 0 (free seen, issued, full seen, drained), clk from start:")
for u in list(range(0, 3)) + list(range(100, 103)):
    print(u, d(mma[u, 0, 0]), d(mma[u, 0, 1]), d(epi[0, u, 0]), d(epi[0, u, 1]))
