"""KNN development probe: tree / grid / brute agreement on the C2 cloud + timings (level-0 self-kNN, pyramid, pair).
    python tools/knn_dev.py [--batch 32] [--n 16384] [--reps 10]"""
import argparse
import ctypes
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
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
dev = "cuda:0"
b = synth.make_batch(a.batch, a.n, 8, "kitti", config=2)
p = b["points_src"].to(dev)
r = b["points_ref"].to(dev)


def t(fn, reps=a.reps):
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


if a.check:
    pc = p[:2].contiguous()
    for k in (1, 16):
        ib, db = D.knn(pc, pc, k, algo=D.KNN_BRUTE)
        for name, algo in (("tree", D.KNN_TREE), ("grid", D.KNN_GRID)):
            i, d = D.knn(pc, pc, k, algo=algo)
            bad = (i != ib).any(dim=2).sum().item()
            print(f"k={k} {name} vs brute: rows differing {bad} / {ib.shape[0] * ib.shape[1]}, dist equal {torch.equal(d, db)}", flush=True)
    gq = D.nn_search_cloud(pc, 16, (4, 4, 4, 4), algo=D.KNN_BRUTE)
    for name, algo in (("tree", D.KNN_TREE), ("grid", D.KNN_GRID), ("auto", D.KNN_AUTO)):
        g = D.nn_search_cloud(pc, 16, (4, 4, 4, 4), algo=algo)
        print(f"pyramid {name} vs brute:", {k2: bool(torch.equal(g[k2], gq[k2])) for k2 in g}, flush=True)

for name, algo in (("tree", D.KNN_TREE), ("grid", D.KNN_GRID)):
    ms0 = t(lambda: D.knn(p, p, 16, algo=algo))
    ms1 = t(lambda: D.nn_search_cloud(p, 16, (4, 4, 4, 4), algo=algo))
    ms2 = t(lambda: D.nn_search_pair(p, r, 16, (4, 4, 4, 4), algo=algo))
    print(f"{name}: level-0 self-kNN {1e3 * ms0:.1f} us / {a.batch} clouds | pyramid {1e3 * ms1:.1f} us | pair of pyramids {1e3 * ms2:.1f} us", flush=True)

# development counters (libraries built with -DDSIR_KNN_STATS only)
import ctypes as _ct
raw = _ct.CDLL(D._lib.LIB_PATH)
if hasattr(raw, "dsir_knn_tree_stats"):
    out = (_ct.c_ulonglong * 8)()
    for label, fn, nq in (("level-0 self-kNN k=16", lambda: D.knn(p, p, 16, algo=D.KNN_TREE), a.batch * a.n),
                          ("1-NN 16384 -> 4096", lambda: D.knn(p[:, :a.n // 4].contiguous(), p, 1, algo=D.KNN_TREE), a.batch * a.n)):
        torch.cuda.synchronize()
        raw.dsir_knn_tree_stats(None, 1)
        fn()
        torch.cuda.synchronize()
        raw.dsir_knn_tree_stats(out, 1)
        w = nq / 32
        print(f"stats {label}: per warp: leaves picked {out[0] / w:.1f}, scanned {out[1] / w:.1f}, lanes needing a scanned leaf {out[2] / max(out[1], 1):.1f}, "
              f"drains {out[3] / w:.1f}, drain iterations {out[4] / w:.1f}, appended per lane {out[5] / nq:.1f}", flush=True)

# per-launch-site times of one pyramid (in-situ profiler: events after every launch)
lib = D.lib()
for name, algo in (("tree", D.KNN_TREE),):
    torch.cuda.synchronize()
    lib.dsir_profile_begin(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    for _ in range(3):
        D.nn_search_cloud(p, 16, (4, 4, 4, 4), algo=algo)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.dsir_profile_report(buf, len(buf))
    print(f"--- in-situ ({name}, 3 pyramids; forked streams make neighbouring sites overlap) ---")
    print(buf.value.decode())
