"""Does the KNN pyramid overlap the persistent match filter?  Times the C2 step of bench.py (32 pairs, device resident) with
the KNN fork launched before / after the match call and with the match stream at default / high priority.

    DSIR_B200_LIB=build/libdeepsir_w2.so python tools/coresidency_probe.py     # a dev build (tools/build_variant.sh)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import deepsir_b200 as D  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    host = bench.make_inputs(bench.BATCH, 0)
    d = {k: v.to(dev) for k, v in host.items()}
    xs0 = d["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr0 = d["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    side = {0: torch.cuda.Stream(dev), -1: torch.cuda.Stream(dev, priority=-1)}

    def knn():
        return D.nn_search_pair(d["points_src"], d["points_ref"], bench.KNN_K, bench.RATIOS)

    def match():
        return D.align_loop(d["feat_src"], d["feat_ref"], xs0, xr0, d["weights"], 1)

    def step(order, knn_stream):
        cur = torch.cuda.current_stream(dev)
        knn_stream.wait_stream(cur)
        if order == "knn_first":
            with torch.cuda.stream(knn_stream):
                g = knn()
            out = match()
        else:
            out = match()
            with torch.cuda.stream(knn_stream):
                g = knn()
        cur.wait_stream(knn_stream)
        return out, g

    lag = {"ev": None}

    def step_lag(knn_stream):
        """KNN of step i runs on its own stream without waiting for the match of step i-1; the main stream joins the KNN of the
        PREVIOUS step (one step of lag), so consecutive steps overlap: prologue of one under the tail of the other."""
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(knn_stream):
            g = knn()
            e = torch.cuda.Event()
            e.record(knn_stream)
        out = match()
        if lag["ev"] is not None:
            cur.wait_event(lag["ev"])
        lag["ev"] = e
        return out, g

    def timed(fn, steps=20, blocks=9):
        for _ in range(5):
            fn()
        ms = []
        for _ in range(blocks):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b) / steps)
        return sorted(ms)[len(ms) // 2]

    print("lib", os.environ.get("DSIR_B200_LIB", "shipped"))
    print("  knn alone            %.3f ms" % timed(knn))
    print("  match+kabsch alone   %.3f ms" % timed(match))
    for main_prio in (0, -1):
        main_stream = torch.cuda.Stream(dev, priority=main_prio)
        with torch.cuda.stream(main_stream):
            for order in ("knn_first", "match_first"):
                print("  main prio %2d  %-12s %.3f ms" % (main_prio, order, timed(lambda: step(order, side[0]))))
            print("  main prio %2d  %-12s %.3f ms" % (main_prio, "lag1", timed(lambda: step_lag(side[0]))))


if __name__ == "__main__":
    main()
