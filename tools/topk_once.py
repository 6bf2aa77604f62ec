"""One fused top-k call per shape (for an ncu launch list):  python tools/topk_once.py [topk]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsir_b200 as D  # noqa: E402

topk = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
B, C, J, K = 32, 32, 5000, 5000
g = torch.Generator().manual_seed(1)
fs = torch.nn.functional.normalize(torch.randn(B, C, J, generator=g), dim=1).to(dev)
fr = torch.nn.functional.normalize(torch.randn(B, C, K, generator=g), dim=1).to(dev)
beta, alpha = torch.full((B,), 10.0, device=dev), torch.full((B,), 0.5, device=dev)
for _ in range(2):
    out = D.match_soft(fs, fr, None, beta, alpha, topk=topk)
torch.cuda.synchronize()
print("ok", out[3].shape)
