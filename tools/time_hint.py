import os, sys, torch
sys.path.insert(0, os.getcwd())
import deepsir_b200 as D
from deepsir_b200 import synth
b = {k: v.to("cuda:0") for k, v in synth.make_batch(32, 16384, 64, "kitti", config=2).items()}
fs, fr = b["feat_src"], b["feat_ref"]
base = D.match_argmin(fs, fr)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("match_argmin unhinted %.3f ms, hinted by the previous result %.3f ms" % (t(lambda: D.match_argmin(fs, fr)), t(lambda: D.match_argmin(fs, fr, prior=base))))
