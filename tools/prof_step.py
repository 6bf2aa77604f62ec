"""One warm + N profiled passes of the hot path at the C2 shape (reduced batch), for ncu launch lists.
    python tools/prof_step.py [--batch 8] [--iters 1] [--what all|match|knn]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--what", default="all")
a = ap.parse_args()
dev = "cuda:0"
b = {k: v.to(dev) for k, v in synth.make_batch(a.batch, a.n, 64, "kitti", config=2).items()}
xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
for it in range(1 + a.iters):
    if a.what in ("all", "knn"):
        D.nn_search_cloud(b["points_src"], 16, (4, 4, 4, 4))
        D.nn_search_cloud(b["points_ref"], 16, (4, 4, 4, 4))
    if a.what in ("all", "match"):
        D.align_loop(b["feat_src"], b["feat_ref"], xs, xr, b["weights"], 1)
    torch.cuda.synchronize()
print("ok")
