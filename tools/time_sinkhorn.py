"""Sinkhorn on the implicit affinity at the C3 shape: 5 iterations (11 fused sweeps).  python tools/time_sinkhorn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402

dev = "cuda:0"
B, N, C = 32, 5000, 32
b = synth.make_batch(B, N, C, "3dmatch", config=3)
ref = b["points_ref"][:, :, :3].contiguous().to(dev)
fs, fr = b["feat_src"].to(dev), b["feat_ref"].to(dev)
beta = torch.full((B,), 10.0, device=dev)
for _ in range(2):
    D.sinkhorn_implicit(fs, fr, ref, beta, 0.5, n_iters=5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    D.sinkhorn_implicit(fs, fr, ref, beta, 0.5, n_iters=5)
e1.record()
torch.cuda.synchronize()
print(f"sinkhorn_implicit, 32 pairs of 5000x5000, 5 iterations: {e0.elapsed_time(e1) / 5:.3f} ms")
