"""K-split sweep of the argmin filter (needs a -DDSIR_TC_TRACE build: DSIR_TC_SPLIT forces S).  One process per setting:
    for s in 0 1 2 4 8; do DSIR_TC_SPLIT=$s DSIR_B200_LIB=build/libdeepsir_trace.so python tools/sweep_split.py; done"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402

dev = "cuda:0"
out = []
for (B, J, K) in [(1, 16384, 131072), (1, 32768, 131072), (1, 65536, 131072), (1, 131072, 131072), (8, 5000, 5000), (2, 20000, 20000)]:
    fs = synth.random_features(B, 64, J, 1).to(dev)
    fr = synth.random_features(B, 64, K, 2).to(dev)
    for _ in range(3):
        D.match_argmin(fs, fr, algo=D.MATCH_TC)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        D.match_argmin(fs, fr, algo=D.MATCH_TC)
    b.record()
    torch.cuda.synchronize()
    out.append("%dx%dx%d %.3f" % (B, J, K, a.elapsed_time(b) / 10))
print("S=%s  " % os.environ.get("DSIR_TC_SPLIT", "auto"), "  ".join(out))
