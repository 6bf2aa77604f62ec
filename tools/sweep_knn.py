"""Sweep the grid-KNN tunables (read from the environment when the library loads) in sub-processes."""
import os
import subprocess
import sys

code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import deepsir_b200 as D
from deepsir_b200 import synth
b = synth.make_batch(32, 16384, 8, "kitti", config=2)
p = b["points_src"].to("cuda:0")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: D.nn_search_cloud(p, 16, (4, 4, 4, 4)))
ms0 = t(lambda: D.knn(p, p, 16))
print(f"cpp={os.environ.get('DSIR_GRID_CPP')} r0={os.environ.get('DSIR_GRID_R0')}: pyramid {1e3*ms/32:6.1f} us/cloud   level0 self-knn {1e3*ms0/32:6.1f} us/cloud")
'''
for cpp in ("1.0", "2.0", "4.0", "8.0"):
    for r0 in ("0.5", "0.75", "1.0", "1.5"):
        env = dict(os.environ, DSIR_GRID_CPP=cpp, DSIR_GRID_R0=r0)
        subprocess.run([sys.executable, "-c", code], env=env)
