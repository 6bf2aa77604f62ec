"""Level-0 self-kNN (k=16) through the bucket-tree path alone, for ncu captures.
    python tools/knn_tree_l0.py [--batch 8] [--algo 3]"""
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
ap.add_argument("--algo", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
p = synth.make_batch(a.batch, a.n, 8, "kitti", config=2)["points_src"].to("cuda:0")
for _ in range(a.reps):
    D.knn(p, p, 16, algo=a.algo)
torch.cuda.synchronize()
print("ok")
