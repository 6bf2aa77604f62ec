#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[2..4]) with the timing rules of bench.py (device events, barrier +
synchronize on both sides, max over ranks, >= 3 warm-up steps).  bench.py stays the C2 headline; this script is for the
other shapes:

    python tools/bench_configs.py --workload c3|c4|c5 [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_configs.py --workload c4 --gpus N

  c3  3DMatch-shaped fragments, 5000 keypoints, 32-d features: soft match (compute_affinity beta=10, alpha=0.5 + row softmax
      + soft targets) + soft weighted Kabsch, 32 pairs per GPU, pair-sharded (no collective)
  c4  one 131072 x 131072 D=64 pair, source rows split over the ranks: local argmin match + local fp64 moments +
      ONE all_reduce of 17 doubles (NCCL) + identical Kabsch solve on every rank  ("strong" scaling: total work fixed)
  c5  Oxford-shaped 20000-point pairs, 10 re-match / re-solve iterations, 8 pairs per GPU, pair-sharded
One JSON line per run on stdout (rank 0).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["c3", "c4", "c5"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=0, help="override the cloud size (tests)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    import deepsir_b200 as D
    from deepsir_b200 import dist as DD, synth
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert D.lib().dsir_device_check() == 0
    warm = max(args.warmup, 3)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = args.workload
    if w == "c3":
        B, N, C = 32, args.n or 5000, 32
        b = synth.make_batch(B, N, C, "3dmatch", config=3, first_pair=rank * B)
        host = dict(src=b["points_src"][:, :, :3].contiguous(), ref=b["points_ref"][:, :, :3].contiguous(),
                    fs=b["feat_src"], fr=b["feat_ref"])
        beta = torch.full((B,), 10.0, device=dev)

        def step(d):
            y, s, _ = D.match_soft(d["fs"], d["fr"], d["ref"], beta, 0.5)
            T, _ = D.kabsch_soft(d["src"], y, s)
            return T
        flops = 2.0 * N * N * C * B
        name = f"C3: synthetic 3DMatch-shaped fragments, {N} keypoints, 32-d, soft match + soft weighted Kabsch, batch 32/GPU"
        pairs_per_step, scaling, shard = B * world, "weak", "by pair, no collective"
    elif w == "c5":
        B, N, C, IT = 8, args.n or 20000, 64, 10
        b = synth.make_batch(B, N, C, "oxford", config=5, first_pair=rank * B)
        host = dict(xs=b["points_src"][:, :, :3].permute(0, 2, 1).contiguous(), xr=b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous(),
                    fs=b["feat_src"], fr=b["feat_ref"], w=b["weights"][:, :, 0].contiguous())

        def step(d):
            tr, pred, xyz, st = D.align_loop(d["fs"], d["fr"], d["xs"], d["xr"], d["w"], IT)
            return tr[-1]
        flops = 2.0 * N * N * C * B * IT
        name = f"C5: Oxford-shaped {N}-pt pairs, {IT} re-match/re-solve iterations, batch 8/GPU"
        pairs_per_step, scaling, shard = B * world, "weak", "by pair, no collective"
    else:
        B, N, C = 1, args.n or 131072, 64
        b = synth.make_batch(B, N, C, "kitti", config=4, first_pair=0)        # the SAME pair on every rank
        lo, hi = DD.row_block(N, world, rank)
        host = dict(xs=b["points_src"][:, lo:hi, :3].permute(0, 2, 1).contiguous(), xr=b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous(),
                    fs=b["feat_src"][:, :, lo:hi].contiguous(), fr=b["feat_ref"], w=b["weights"][:, lo:hi, 0].contiguous())

        def step(d):
            tr, pred, xyz, st = DD.align_rowblock(d["fs"], d["fr"], d["xs"], d["xr"], d["w"], 1)
            return tr[-1]
        flops = 2.0 * (hi - lo) * N * C
        name = f"C4: one {N} x {N} D=64 pair, source rows sharded over {world} rank(s), NCCL all_reduce of fp64 moments"
        pairs_per_step, scaling, shard = 1, "strong", "by source-row block; one all_reduce(SUM) of [B,17] fp64 per iteration"

    pinned = {k: v.pin_memory() for k, v in host.items()}
    devt = {k: v.to(dev) for k, v in host.items()}
    # ---- device resident ----
    for _ in range(warm):
        T = step(devt)
    barrier()
    e0, e1 = ev(), ev()
    l0 = D.lib().dsir_launch_count()
    e0.record()
    for _ in range(args.steps):
        T = step(devt)
    e1.record()
    barrier()
    launches = D.lib().dsir_launch_count() - l0
    ms = e0.elapsed_time(e1)
    # ---- end to end: pinned host inputs up, pose down, every step ----
    out_h = torch.empty(T.shape, dtype=torch.float32, pin_memory=True)
    for _ in range(2):
        d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        out_h.copy_(step(d), non_blocking=True)
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        out_h.copy_(step(d), non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the pose is on the host before the next step starts
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    T_all = T.clone()
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        if w == "c4":   # every rank must hold the same pose
            ref = T.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, T), "ranks disagree on the pose"
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for v in pinned.values())
        gt = b["transform_gt"][:B]
        from oracle import deepsir_oracle as O   # checker only: pose error of the timed result against the planted pose
        err = O.rotation_angle_deg(T_all.cpu()[:, :, :3], gt[:, :, :3]).max().item()
        out = {"metric": "pairs/sec", "value": pairs_per_step * args.steps / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
               "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": name, "sharding": shard, "l2": "no explicit flush; see bench.py for the headline"},
               "e2e": {"value": pairs_per_step * args.steps / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": out_h.numel() * 4},
               "gpu_launches": int(launches),
               "match_tflops_per_gpu": flops * args.steps / (ms / 1e3) / 1e12,
               "max_rotation_error_vs_planted_deg": err}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
