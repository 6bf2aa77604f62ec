"""Eager step vs CUDA-graph replay of the same step (C2 shape, device resident).
    python tools/graph_probe.py [--batch 32] [--iters 1]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from bench import make_inputs, KNN_K, RATIOS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
d = {k: v.to(dev) for k, v in make_inputs(a.batch, 0).items()}
xs = d["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
xr = d["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()


def eager():
    D.nn_search_pair(d["points_src"], d["points_ref"], KNN_K, RATIOS)
    return D.align_loop(d["feat_src"], d["feat_ref"], xs, xr, d["weights"], a.iters)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = D.GraphedRegistration(d, KNN_K, RATIOS, iters=a.iters)
print(f"eager step: {timeit(eager):.3f} ms   graph replay: {timeit(lambda: g.step()):.3f} ms   (B={a.batch}, iters={a.iters})")
