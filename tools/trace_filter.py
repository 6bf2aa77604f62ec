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
mma = t[:1024].reshape(256, 2, 2)      # [useq][row block][free seen, issued]
epi = t[1024:1024 + 2048].reshape(256, 8)   # [useq][k]
t0 = mma[0, 0, 0]
d = lambda x: (x - t0) & 0xffffffff
sl = slice(40, 200)
print(f"period per unit (issuer of row block 0): {np.diff(d(mma[sl, 0, 0])).mean():.0f} clk")
for r in range(2):
    print(f"issuer {r}: free seen -> issued {np.mean(d(mma[sl, r, 1]) - d(mma[sl, r, 0])):6.0f}")
e = d(epi[sl])
names = ["full seen -> 32 cols in regs", "step 1 (tree, vote, slow path)", "wait second 32 cols", "step 2", "fence + arrive"]
print(f"epilogue warp 0: issued -> full seen {np.mean(e[:, 0] - d(mma[sl, 0, 1])):6.0f}")
for k, n in enumerate(names):
    dd = e[:, k + 1] - e[:, k]
    print(f"   {n:34s} mean {np.mean(dd):6.0f}  median {np.median(dd):6.0f}  p10 {np.percentile(dd, 10):6.0f}  p90 {np.percentile(dd, 90):6.0f}")
uu = np.diff(d(mma[sl, 0, 0]))
print(f"   unit period: median {np.median(uu):.0f}  p10 {np.percentile(uu, 10):.0f}  p90 {np.percentile(uu, 90):.0f}")
print(f"   drain total {np.mean(e[:, 5] - e[:, 0]):6.0f};  released -> issuer sees it free (2 units later) "
      f"{np.mean(d(mma[sl, 0, 0])[2:] - e[:-2, 5]):6.0f}")
