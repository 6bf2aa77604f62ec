"""Where does the end-to-end step go?  Raw pinned->device copy rate of one C2 batch vs the pipelined step.
    python tools/e2e_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from bench import make_inputs, KNN_K, RATIOS  # noqa: E402

dev = torch.device("cuda:0")
host = make_inputs(32, 0)
pinned = {k: v.pin_memory() for k, v in host.items()}
nbytes = sum(v.numel() * v.element_size() for v in pinned.values())
dst = {k: torch.empty_like(v, device=dev) for k, v in pinned.items()}
ev = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(2):
    for k in pinned:
        dst[k].copy_(pinned[k], non_blocking=True)
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
for _ in range(5):
    for k in pinned:
        dst[k].copy_(pinned[k], non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"raw H2D of one batch ({nbytes / 1e6:.0f} MB) into preallocated buffers: {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s")
big = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dbig = torch.empty(nbytes, dtype=torch.uint8, device=dev)
dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    dbig.copy_(big, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"one contiguous copy of the same size: {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s")
pipe = D.RegistrationPipeline(dev, KNN_K, RATIOS, iters=1, depth=2)
for _ in pipe.run(pinned for _ in range(3)):
    pass
torch.cuda.synchronize()
for depth in (2, 3):
    pipe = D.RegistrationPipeline(dev, KNN_K, RATIOS, iters=1, depth=depth)
    for _ in pipe.run(pinned for _ in range(3)):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    n = 0
    for _ in pipe.run(pinned for _ in range(10)):
        n += 1
    e1.record()
    torch.cuda.synchronize()
    print(f"pipeline depth {depth}: {e0.elapsed_time(e1) / 10:.3f} ms per step (host wall {1e3 * (time.perf_counter() - t0) / 10:.3f})")
