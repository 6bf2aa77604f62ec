"""Level-0 self-kNN (16384 points, k=16) alone, for ncu captures and quick timings.
    python tools/knn_l0.py [--batch 32] [--reps 5]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
p = synth.make_batch(a.batch, a.n, 8, "kitti", config=2)["points_src"].to("cuda:0")
for _ in range(2):
    D.knn(p, p, 16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    D.knn(p, p, 16)
e1.record()
torch.cuda.synchronize()
print(f"level-0 self-kNN B={a.batch}: {1e3 * e0.elapsed_time(e1) / a.reps / a.batch:.1f} us/cloud")
