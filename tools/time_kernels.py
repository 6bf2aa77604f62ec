"""CUDA-event timings of the two dominant stages at the C2 shape (not under a profiler).
    python tools/time_kernels.py [--batch 32] [--reps 10]"""
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
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev = "cuda:0"
b = {k: v.to(dev) for k, v in synth.make_batch(a.batch, a.n, 64, "kitti", config=2).items()}


def timeit(fn, reps=a.reps):
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


flops = 2.0 * a.n * a.n * 64 * a.batch
for name, algo in (("tc", D.MATCH_TC), ("fp32", D.MATCH_FP32)):
    if name == "fp32" and a.batch > 8:
        continue
    ms = timeit(lambda: D.match_argmin(b["feat_src"], b["feat_ref"], algo=algo))
    print(f"match_argmin[{name}] B={a.batch}: {ms:8.3f} ms  {1e3 * ms / a.batch:7.1f} us/pair  {flops / ms / 1e9:7.1f} TFLOP/s")
tm = {}
_, n_resc = D.match_argmin(b["feat_src"], b["feat_ref"], algo=D.MATCH_TC, return_rescued=True, timing=tm)
print("rescued rows:", n_resc, "of", a.batch * a.n)
print("filter kernel (device timers, steady state): span %.1f us = %.1f us/pair, %.0f cycles/CTA (%.2f GHz), %.0f cycles per "
      "item-unit (tensor floor 320 per 128x128 tile)" % (tm["span_ns"] / 1e3, tm["span_ns"] / 1e3 / a.batch, tm["cycles_per_cta"],
                                            tm["cycles_per_cta"] / tm["span_ns"], tm["cycles_per_unit"]))
for name, algo in (("grid", D.KNN_AUTO), ("brute", D.KNN_BRUTE)):
    if name == "brute" and a.batch > 8:
        continue
    ms = timeit(lambda: D.nn_search_cloud(b["points_src"], 16, (4, 4, 4, 4), algo=algo))
    print(f"knn pyramid[{name}] B={a.batch} clouds: {ms:8.3f} ms  {1e3 * ms / a.batch:7.1f} us/cloud")
xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
idx = D.match_argmin(b["feat_src"], b["feat_ref"])
ms = timeit(lambda: D.kabsch_gather(xs, xr, idx, b["weights"]))
print(f"kabsch(gather) B={a.batch}: {ms:8.3f} ms")
