"""In-situ per-launch-site timing of the steady-state step (library event profiler, no ncu).
    python tools/prof_insitu.py [--batch 32] [--iters 5] [--what all|match|knn]"""
import argparse
import ctypes
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--what", default="all", help="all | match | knn | topk (config-3-shaped soft match + top-k)")
ap.add_argument("--topk", type=int, default=32)
a = ap.parse_args()
dev = "cuda:0"
b = {k: v.to(dev) for k, v in synth.make_batch(a.batch, a.n, 64, "kitti", config=2).items()}
xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()


if a.what == "topk":
    g = torch.Generator().manual_seed(1)
    tfs = torch.nn.functional.normalize(torch.randn(a.batch, 32, 5000, generator=g), dim=1).to(dev)
    tfr = torch.nn.functional.normalize(torch.randn(a.batch, 32, 5000, generator=g), dim=1).to(dev)
    tbeta, talpha = torch.full((a.batch,), 10.0, device=dev), torch.full((a.batch,), 0.5, device=dev)


def step():
    if a.what == "topk":
        D.match_soft(tfs, tfr, None, tbeta, talpha, topk=a.topk)
    if a.what in ("all", "knn"):
        D.nn_search_cloud(b["points_src"], 16, (4, 4, 4, 4))
        D.nn_search_cloud(b["points_ref"], 16, (4, 4, 4, 4))
    if a.what in ("all", "match"):
        D.align_loop(b["feat_src"], b["feat_ref"], xs, xr, b["weights"], 1)


for _ in range(3):
    step()
torch.cuda.synchronize()
lib = D.lib()
lib.dsir_profile_begin(torch.cuda.current_stream().cuda_stream)
for _ in range(a.iters):
    step()
buf = ctypes.create_string_buffer(1 << 16)
lib.dsir_profile_report(buf, len(buf))
txt = buf.value.decode()


def label(m):  # file:line -> the kernel launched on the line(s) just above
    f, ln = m.group(1), int(m.group(2))
    try:
        lines = open(os.path.join(ROOT, "deepsir_b200", "csrc", f)).read().split("\n")
        for i in range(ln - 1, max(ln - 8, 0), -1):
            k = re.search(r"(\w+)(<[^<>]*>)?<<<", lines[i])
            if k:
                return f"{k.group(1):32s}"
    except OSError:
        pass
    return f"{m.group(0):32s}"


print(f"per {a.iters} steps of {a.batch} pairs:")
print(re.sub(r"(\w+\.cu):(\d+)\s*", label, txt))
