"""Re-run one trial of tools/fuzz_parity.py::fuzz_loop and classify the rows that differ:  python tools/repro_loop_fuzz.py <seed>"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402
from deepsir_b200 import synth  # noqa: E402
from oracle import deepsir_oracle as O  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fuzz_parity as F  # noqa: E402

seed = int(sys.argv[1])
rng = random.Random(seed)
B, C, n, iters = rng.randint(1, 3), rng.choice([16, 32, 64]), max(64, F.size(rng, 3000)), rng.randint(1, 4)
kind = rng.choice(["kitti", "oxford"])
b = synth.make_batch(B, n, C, kind, config=5, first_pair=seed % 997)
print("trial", B, C, n, iters, kind)
fs, fr = b["feat_src"], b["feat_ref"]
io = O.match_argmin(fs, fr)
cu = lambda t: t.to("cuda:0")
for name, algo in (("tc", D.MATCH_TC), ("fp32", D.MATCH_FP32), ("auto", D.MATCH_AUTO)):
    ig = D.match_argmin(cu(fs), cu(fr), algo=algo).cpu()
    diff = (ig != io).nonzero()
    print(name, "rows differing from the oracle:", diff.tolist())
_, gap = O.match_top2_fp64(fs, fr)
d64 = ((fs.double().transpose(1, 2)[:, :, None, :] - fr.double().transpose(1, 2)[:, None, :, :]) ** 2).sum(-1)
ig = D.match_argmin(cu(fs), cu(fr)).cpu()
for bb, j in (ig != io).nonzero().tolist():
    ko, kg = int(io[bb, j]), int(ig[bb, j])
    d32 = O.match_features_V2(fs[bb:bb + 1], fr[bb:bb + 1])[0, j]
    print(f"  b{bb} row {j}: oracle {ko} lib {kg}  fp64 d {d64[bb, j, ko].item():.9f} vs {d64[bb, j, kg].item():.9f}  gap {gap[bb, j].item():.3e}"
          f"  reference fp32 d {d32[ko].item():.9f} vs {d32[kg].item():.9f}")
