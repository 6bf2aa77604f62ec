"""In-situ per-launch-site timing of the C3 soft step (match_soft + kabsch_soft).  python tools/prof_soft.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402

dev = "cuda:0"
B, N, C = 32, 5000, 32
b = synth.make_batch(B, N, C, "3dmatch", config=3)
src, ref = b["points_src"][:, :, :3].contiguous().to(dev), b["points_ref"][:, :, :3].contiguous().to(dev)
fs, fr = b["feat_src"].to(dev), b["feat_ref"].to(dev)
beta = torch.full((B,), 10.0, device=dev)


def step():
    y, s, _ = D.match_soft(fs, fr, ref, beta, 0.5)
    return D.kabsch_soft(src, y, s)


for _ in range(3):
    step()
torch.cuda.synchronize()
lib = D.lib()
lib.dsir_profile_begin(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    step()
buf = ctypes.create_string_buffer(1 << 16)
lib.dsir_profile_report(buf, len(buf))
print("per 5 steps of 32 pairs:")
print(buf.value.decode())
