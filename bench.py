#!/usr/bin/env python
"""bench.py — pairs/sec of the correspondence-and-pose hot path (BASELINE.json metric).

Workload at every N (weak scaling, pair-sharded, no data-path collective): BASELINE.json configs[1] —
synthetic KITTI-shaped pairs, 16384 pts/cloud, batch 32 per GPU, k=16 KNN pyramid (4 levels + 1-NN upsample) on
both clouds + full 16384x16384 D=64 match (fused argmin) + gather + weighted Kabsch + transform + compose
(one registration iteration, the "KNN + full match + Kabsch" unit of SURVEY §8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` = pairs/s with inputs resident in HBM; `e2e` = the same through the
public host API with pinned HOST buffers (H2D of every input and D2H of transforms + correspondences inside the
timed region); `roofline` = the dominant kernel (feature match) against the measured tensor peak;
`cpu_baseline` = the CPU oracle (port of the reference path) timed on this box's host cores.
`--impl reference` times that CPU port alone with all host threads (the reference has no GPU kernels of its own
for this path and cannot be shipped to the box; the oracle restates it op for op — see DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PTS, FEAT_D, KNN_K, RATIOS, BATCH = 16384, 64, 16, (4, 4, 4, 4), 32
METRIC = "pairs/sec (16k pts/cloud)"
WORKLOAD = "C2: synthetic KITTI-shaped pairs, 16384 pts/cloud, batch 32/GPU, k=16 KNN pyramid x2 + 16384x16384 D=64 match + Kabsch"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        recs = []
        for r in self.rows:
            try:
                recs.append((float(r[1]), float(r[2]), float(r[3]), [v.lower().startswith("active") for v in r[4:8]]))
            except Exception:
                continue
        # the sampler also sees the idle gaps around the timed loops: "under load" = power above half of the maximum seen
        pmax = max((r[2] for r in recs), default=0.0)
        load = [r for r in recs if r[2] >= 0.5 * pmax] or recs
        reasons = set()
        for r in load:
            for name, on in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if on:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[0] for r in load) if load else None,
                "sm_max_mhz": max((r[1] for r in load), default=None), "power_w_max": pmax if recs else None,
                "reasons": sorted(reasons), "samples": len(load), "samples_total": len(recs)}


def filter_traffic_from_digest():
    """dram__bytes_read.sum + dram__bytes_write.sum of the filter kernel from the newest committed ncu digest
    (profiles/ncu_digest_filter_*.txt, one `ncu --set full` capture per round), per launch; (bytes, file) or (None, None)."""
    import glob
    import re
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_digest_filter_*.txt"))):
        best = f                                  # names sort by round: r1c < r1f < r2a ...
    if best is None:
        return None, None
    rd = wr = None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(best):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m:
            v = float(m.group(2)) * unit.get(m.group(3), 1.0)
            if m.group(1) == "read" and rd is None:
                rd = v
            if m.group(1) == "write" and wr is None:
                wr = v
    if rd is None or wr is None:
        return None, None
    return rd + wr, os.path.relpath(best, ROOT)


def make_inputs(batch, first_pair):
    from deepsir_b200 import synth
    b = synth.make_batch(batch, N_PTS, FEAT_D, "kitti", config=2, first_pair=first_pair)
    return dict(points_src=b["points_src"], points_ref=b["points_ref"], feat_src=b["feat_src"], feat_ref=b["feat_ref"],
                weights=b["weights"][:, :, 0].contiguous())


REF_COPY = os.path.join(ROOT, "baseline", "_ref")
_ref_mods = None


def reference_modules():
    """The reference's own modules from baseline/_ref (unmodified copy shipped by __graft_entry__.build()), or None."""
    global _ref_mods
    if _ref_mods is None:
        _ref_mods = False
        if os.path.isdir(os.path.join(REF_COPY, "network")):
            import warnings
            warnings.filterwarnings("ignore")
            if REF_COPY not in sys.path:
                sys.path.insert(0, REF_COPY)
            try:
                from network import model as M
                _ref_mods = M
            except Exception:
                _ref_mods = False
    return _ref_mods or None


def cpu_step(host, n_pairs):
    """One step of the C2 workload on the host CPU for the first n_pairs pairs.  With baseline/_ref present this is the
    REFERENCE's own code: match_features_V2 + .min in 6000-row chunks (network/model.py:558-569), gather_neighbour_V3
    (:571), compute_rigid_transform_2 incl. its fp64 LAPACK SVD (:588, :22-66), se3_torch.transform (:590) — the statements
    of one registration iteration, on the synthetic features; otherwise the oracle port of the same statements.  The KNN
    pyramids: the reference's torch_points_kernels.knn is not installable here, so both variants use the kd-tree stand-in
    (scipy cKDTree, all cores — the algorithm class of its nanoflann) through the oracle's nn_search restatement."""
    from oracle import deepsir_oracle as O
    s = slice(0, n_pairs)
    O.nn_search_kdtree(host["points_src"][s], KNN_K, RATIOS)
    O.nn_search_kdtree(host["points_ref"][s], KNN_K, RATIOS)
    xs = host["points_src"][s, :, :3].permute(0, 2, 1).contiguous()
    xr = host["points_ref"][s, :, :3].permute(0, 2, 1).contiguous()
    M = reference_modules()
    if M is None:
        O.align_loop(host["feat_src"][s], host["feat_ref"][s], xs, xr, host["weights"][s, :, None], 1)
        return
    with torch.no_grad():
        feat_src, feat_ref = host["feat_src"][s], host["feat_ref"][s]
        stride, N = 6000, feat_src.shape[2]
        indexs = []
        for n in range((N + stride - 1) // stride):
            mm = M.match_features_V2(feat_src[:, :, n * stride:(n + 1) * stride], feat_ref)
            indexs.append(mm.min(dim=2, keepdim=False)[1])
        indexs = torch.cat(indexs, dim=1)
        xyz_ref_new = M.gather_neighbour_V3(xr, indexs)
        a = xs.permute(0, 2, 1).contiguous()
        bnew = xyz_ref_new.permute(0, 2, 1).contiguous()
        R_t, _ = M.compute_rigid_transform_2(a, bnew, weights=host["weights"][s, :, None])
        M.se3_torch.transform(R_t.detach(), a)


def cpu_kind():
    return "reference" if reference_modules() is not None else "port"


def cpu_sample_text(reps, n_pairs, cores):
    if reference_modules() is not None:
        return (f"{reps} x {n_pairs} pair of the C2 workload on the host: the reference's own match_features_V2 + min "
                f"(6000-row chunks), gather_neighbour_V3, compute_rigid_transform_2 (fp64 LAPACK SVD), se3_torch.transform from "
                f"baseline/_ref; KNN pyramids by the kd-tree stand-in (scipy cKDTree; torch_points_kernels is not installable), "
                f"{cores} threads")
    return (f"{reps} x {n_pairs} pair of the C2 workload on the host: oracle port (torch-CPU MKL sgemm/LAPACK + scipy "
            f"cKDTree KNN), {cores} threads")


def c1_reference_forward():
    """BASELINE.json configs[0]: the reference's Network.forward_align_4 (network/model.py:520-607, test.py:399-402) on one
    synthetic 4096-point KITTI-shaped pair, batch 1, 5 iterations, on the CPU; seeded random-init weights (the shipped
    checkpoint is a missing blob).  Returns seconds per pair (median of 3), or None without baseline/_ref."""
    M = reference_modules()
    if M is None:
        return None
    import arguments
    from deepsir_b200 import synth
    from oracle import deepsir_oracle as O
    args = arguments.eval_arguments().parse_args([])
    torch.manual_seed(0)
    net = M.Network(args).eval()
    b = synth.make_batch(1, 4096, 64, "kitti", config=1)
    data = {"points_src": b["points_src"], "points_ref": b["points_ref"]}
    ts = []
    with torch.no_grad():
        for _ in range(4):
            t0 = time.perf_counter()
            d = dict(data)
            for key in ("points_src", "points_ref"):           # loader-side KNN (data_base.py:153-183): kd-tree stand-in
                g = O.nn_search_kdtree(d[key], KNN_K, RATIOS)
                for name in ("xyz", "neigh_idx", "sub_idx", "interp_idx"):
                    d[key + "_" + name] = g[name]
            net(d, (args.num_reg_iter, False))
            ts.append(time.perf_counter() - t0)
    return statistics.median(ts[1:])


def torch_gpu_step(devt, xs0, xr0, n_pairs):
    """The reference's own statements for match + gather + Kabsch + transform (model.py:551-601, restated in the oracle)
    executed by stock PyTorch on the GPU: cuBLAS sgemm + torch.min per 6000-row chunk, covariance by torch.matmul, the 3x3
    SVD in fp64 on the host exactly as the reference does it.  No KNN: the reference's KNN has no GPU implementation."""
    from oracle import deepsir_oracle as O
    s = slice(0, n_pairs)
    return O.align_loop(devt["feat_src"][s], devt["feat_ref"][s], xs0[s], xr0[s], devt["weights"][s, :, None], 1)


def measure_rowblock(D, dist, dev, rank, world, ev, barrier, n=131072, steps=10):
    """C4 strong scaling inside the N-rank run: the sharded step (graph replay) on all ranks, then the SAME pair unsharded on
    rank 0 alone (dsir_align_loop, one iteration) as the N=1 time of this very box."""
    from deepsir_b200 import dist as DD, synth
    torch.cuda.empty_cache()
    b = synth.make_batch(1, n, FEAT_D, "kitti", config=4, first_pair=0)        # the same pair on every rank
    lo, hi = DD.row_block(n, world, rank)
    xs_all = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous().to(dev)
    fr = b["feat_ref"].to(dev)
    g = DD.GraphedRowBlock(b["feat_src"][:, :, lo:hi].contiguous().to(dev), fr, xs_all[:, :, lo:hi].contiguous().to(dev), xr,
                           b["weights"][:, lo:hi, 0].contiguous().to(dev), num_iter=1)
    for _ in range(3):
        g.step()
    times = []
    for _ in range(3):                     # three blocks of `steps` steps, median
        barrier()
        a, c = ev(), ev()
        a.record()
        for _ in range(steps):
            T, idx, _ = g.step()
        c.record()
        barrier()
        times.append(a.elapsed_time(c) / steps)
    t = torch.tensor([statistics.median(times)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_n = t.item()
    Ts = [torch.empty_like(T) for _ in range(world)]
    dist.all_gather(Ts, T.contiguous())
    identical = all(torch.equal(Ts[0], x) for x in Ts)
    res = None
    if rank == 0:                           # the unsharded pair on one GPU of the same box
        fs_all, w_all = b["feat_src"].to(dev), b["weights"][:, :, 0].contiguous().to(dev)
        xs_d = xs_all.to(dev)
        for _ in range(2):
            tr, pred, _, _ = D.align_loop(fs_all, fr, xs_d, xr, w_all, 1)
        torch.cuda.synchronize()
        a, c = ev(), ev()
        a.record()
        for _ in range(5):
            tr, pred, _, _ = D.align_loop(fs_all, fr, xs_d, xr, w_all, 1)
        c.record()
        torch.cuda.synchronize()
        ms_1 = a.elapsed_time(c) / 5
        from oracle import deepsir_oracle as O   # the checker: angle between the sharded and the unsharded pose
        ang = O.rotation_angle_deg(Ts[0].cpu()[:, :, :3], tr[-1].cpu()[:, :, :3]).max().item()
        dtr = (Ts[0].cpu()[:, :, 3] - tr[-1].cpu()[:, :, 3]).norm(dim=1).max().item()
        res = {"workload": f"C4: one {n} x {n} D={FEAT_D} pair, source rows sharded over {world} ranks, NCCL all_reduce of fp64 "
                           "moments, the rank's kernels of an iteration captured in two CUDA graphs around the eager all_reduce" + ("" if g.graphed else " (capture refused: eager)"),
               "ms_per_pair": ms_n, "pairs_per_s": 1e3 / ms_n, "ms_per_pair_n1": ms_1, "eff_vs_n1": ms_1 / (world * ms_n),
               "T_identical_across_ranks": bool(identical), "vs_unsharded_deg": ang, "vs_unsharded_m": dtr,
               "graphed": bool(g.graphed), "scaling": "strong"}
    barrier()
    del g
    torch.cuda.empty_cache()
    return res


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n_pairs = 1
    host = make_inputs(n_pairs, 0)
    for _ in range(args.warmup):
        cpu_step(host, n_pairs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(host, n_pairs)
    dt = time.perf_counter() - t0
    val = n_pairs * args.steps / dt
    cores = os.cpu_count() or 1
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": f"{n_pairs} pair per step on the host CPU"},
           "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": cpu_kind(),
                            "sample": cpu_sample_text(args.steps, n_pairs, cores)},
           "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    import deepsir_b200 as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deepsir_b200 has no CPU path")
    # N < visible GPUs: consecutive GPUs of these boxes share one PCIe switch uplink (measured r1: 28 GB/s per rank with
    # ranks on GPUs 0-3 against 54 GB/s alone), so the ranks are spread over the visible devices instead of packed
    n_vis = torch.cuda.device_count()
    dev_index = local * (n_vis // world) if (world > 1 and n_vis >= 2 * world and local < world) else local
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    torch.cuda.set_stream(torch.cuda.Stream(dev, priority=-1))     # the main stream of this process (see step_resident)
    local = dev_index
    if True:
        # one process per GPU: keep the rank (and the pinned host buffers it is about to allocate: first touch) on the CPUs /
        # NUMA node closest to its GPU, otherwise eight uploads fight over one socket's memory and root complex
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + bit for w, m in enumerate(words) for bit in range(64) if (m >> bit) & 1]
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert D.lib().dsir_device_check() == 0
    B = args.batch
    warm = max(args.warmup, 3)

    host = make_inputs(B, first_pair=rank * B)                      # pair-sharded: every rank owns its own 32 pairs
    pinned = {k: v.pin_memory() for k, v in host.items()}
    devt = {k: v.to(dev) for k, v in host.items()}
    xs0 = devt["points_src"][:, :, :3].permute(0, 2, 1).contiguous()   # the loop's [B,3,N] layout (model.py:541-549)
    xr0 = devt["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    match_events, knn_events = [], []

    knn_stream = torch.cuda.Stream(dev)
    knn_stream.wait_stream(torch.cuda.current_stream(dev))          # the inputs above were produced on the main stream
    knn_pending = {"ev": None}

    def join_knn():
        """The main stream waits for the KNN pyramids still in flight on the forked stream."""
        if knn_pending["ev"] is not None:
            torch.cuda.current_stream(dev).wait_event(knn_pending["ev"])
            knn_pending["ev"] = None

    def step_resident(record=False, d=None):
        d = d or devt
        cur = torch.cuda.current_stream(dev)
        if record:      # attribution passes: everything in sequence on one stream, bracketed by events
            k0, k1 = ev(), ev()
            k0.record()
            D.nn_search_pair(devt["points_src"], devt["points_ref"], KNN_K, RATIOS)
            k1.record()
            knn_events.append((k0, k1))
            e0, e1 = ev(), ev()
            e0.record()
            D.match_argmin(devt["feat_src"], devt["feat_ref"])        # the dominant kernel, timed on its own stream
            e1.record()
            match_events.append((e0, e1))
            return D.align_loop(devt["feat_src"], devt["feat_ref"], xs0, xr0, devt["weights"], 1)
        # The KNN pyramids of a batch do not depend on its match / Kabsch (in the reference they run in the DataLoader
        # workers while the GPU is busy with the previous batch): they go to a forked stream and are joined at the end of
        # the step, so that their latency-bound kernels fill the gaps around the persistent match kernel.
        # Nor do they depend on the PREVIOUS batch's match: the forked stream runs pyramid after pyramid in its own order and
        # the main stream joins the pyramids of step i at the end of step i+1 (one step of lag; `join_knn` before every end
        # event), so the pack / tree-build prologue of one step runs under the tail of the previous one.  The main stream
        # has the higher priority: the persistent match kernel takes its SMs first, the KNN CTAs fill in around it.
        with torch.cuda.stream(knn_stream):
            g = D.nn_search_pair(d["points_src"], d["points_ref"], KNN_K, RATIOS)
            e = torch.cuda.Event()
            e.record(knn_stream)
        out = D.align_loop(d["feat_src"], d["feat_ref"], xs0, xr0, d["weights"], 1)
        join_knn()
        knn_pending["ev"] = e
        return out, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput (`value`)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.1 s to come up: started before the warm-up so that it is sampling
        time.sleep(0.3)          # by the time the timed loops run; it keeps sampling through both timed regions
    for _ in range(warm):
        step_resident()

    def timed_block(fn):
        """EXACTLY args.steps steps between two events, barrier + synchronize on both sides; ms."""
        barrier()
        a, b_ = ev(), ev()
        a.record()
        for _ in range(args.steps):
            fn()
        join_knn()               # every pyramid of the block has finished before the end event
        b_.record()
        barrier()
        return a.elapsed_time(b_)

    def agree_max(v):
        if world > 1:
            t_ = torch.tensor([float(v)], device=dev, dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return t_.item()
        return float(v)

    l0 = D.lib().dsir_launch_count()
    blocks = [timed_block(step_resident)]
    launches = D.lib().dsir_launch_count() - l0
    # one block of 20 steps lasts ~50 ms - too short for the clock sampler to see the GPU under load.  The K-step block is
    # repeated until ~1 s has been timed (same count on every rank); the reported time is the MEDIAN block.
    n_blocks = int(min(40, max(1, -(-1000.0 // max(agree_max(blocks[0]), 1e-3)))))
    for _ in range(n_blocks - 1):
        blocks.append(timed_block(step_resident))
    ms = statistics.median(blocks)
    # the same step on UN-PLANTED unit features (small top-2 gaps, the regime of learned descriptors): the candidate lists of
    # the filter stay longer, more rows go through the exact re-scoring; reported beside the headline (planted matches)
    from deepsir_b200 import synth
    d_rand = dict(devt, feat_src=synth.random_features(B, FEAT_D, N_PTS, 7001 + rank).to(dev),
                  feat_ref=synth.random_features(B, FEAT_D, N_PTS, 9001 + rank).to(dev))
    for _ in range(2):
        step_resident(d=d_rand)
    join_knn()
    ms_rand = timed_block(lambda: step_resident(d=d_rand))
    _, n_rescued = D.match_argmin(d_rand["feat_src"], d_rand["feat_ref"], algo=D.MATCH_TC, return_rescued=True)
    join_knn()
    torch.cuda.synchronize()
    del d_rand
    # the same step with the reference's default of 5 registration iterations (arguments.py:69), reported beside the headline
    def step_r5():
        cur = torch.cuda.current_stream(dev)
        knn_stream.wait_stream(cur)
        with torch.cuda.stream(knn_stream):
            g = D.nn_search_pair(devt["points_src"], devt["points_ref"], KNN_K, RATIOS)
        out = D.align_loop(devt["feat_src"], devt["feat_ref"], xs0, xr0, devt["weights"], 5)
        cur.wait_stream(knn_stream)
        return out, g
    for _ in range(2):
        step_r5()
    barrier()
    r0, r1 = ev(), ev()
    r0.record()
    for _ in range(max(args.steps // 2, 1)):
        step_r5()
    r1.record()
    barrier()
    ms_r5 = r0.elapsed_time(r1) / max(args.steps // 2, 1)
    # dominant-kernel timing: separate passes (so the headline loop above carries no extra events) with the library's
    # in-situ profiler: a CUDA event recorded on the launching stream right after every kernel launch
    import ctypes
    import re
    lib = D.lib()
    n_prof = 3
    torch.cuda.empty_cache()          # the timed loop used the forked KNN stream's pool: start the sequential passes clean
    step_resident(record=True)        # (absorbs the allocations of the first sequential pass)
    torch.cuda.synchronize()
    match_events.clear(); knn_events.clear()
    lib.dsir_profile_begin(torch.cuda.current_stream().cuda_stream)
    for _ in range(n_prof):
        step_resident(record=True)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.dsir_profile_report(buf, len(buf))
    sites = {}
    for line in buf.value.decode().splitlines():
        m = re.match(r"(\S+)\s+launches\s+(\d+)\s+total\s+([0-9.]+) us", line)
        if m and m.group(1) != "TOTAL":
            sites[m.group(1)] = (int(m.group(2)), float(m.group(3)))
    tc_sites = {k: v for k, v in sites.items() if k.startswith("match_tc.cu")}
    filt_site = max(tc_sites, key=lambda k: tc_sites[k][1])          # the tcgen05 filter is the largest match_tc.cu site
    filter_ms = tc_sites[filt_site][1] / tc_sites[filt_site][0] / 1e3  # per launch (the profiled passes run match twice)
    knn_ms = statistics.mean(a.elapsed_time(b) for a, b in knn_events)   # both pyramids of one step (fork/join bracketed)
    match_ms = statistics.mean(a.elapsed_time(b) for a, b in match_events)

    # ---------------------------------------------------------------- end to end through the host API (`e2e`)
    # RegistrationPipeline.run: every step uploads its inputs from pinned host memory (copy stream, overlapped with
    # the previous step's kernels) and downloads transforms + int32 correspondences (model.py:599-601) to the host.
    torch.cuda.empty_cache()
    pipe = D.RegistrationPipeline(dev, KNN_K, RATIOS, iters=1, depth=2)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = 0
    # warm-up: allocator pools of both streams AND the PCIe link (after a second of compute-only blocks the first uploads
    # run far below the link rate: 18 GB/s was measured on a cold link against 54 GB/s warm)
    for out_h in pipe.run(pinned for _ in range(max(20, 3))):
        d2h = sum(out_h[k].numel() * out_h[k].element_size() for k in ("T", "pred", "status"))

    def e2e_block(batch):
        barrier()
        a, b_ = ev(), ev()
        a.record()
        n_out = 0
        for _ in pipe.run(batch for _ in range(args.steps)):
            n_out += 1                                              # results are on the host here (event-synchronised)
        b_.record()
        barrier()
        assert n_out == args.steps
        return a.elapsed_time(b_)

    e2e_blocks = [e2e_block(pinned)]
    n_eb = int(min(10, max(1, -(-500.0 // max(agree_max(e2e_blocks[0]), 1e-3)))))
    for _ in range(n_eb - 1):
        e2e_blocks.append(e2e_block(pinned))
    ms_e2e = statistics.median(e2e_blocks)
    # the same pipeline with the FEATURES already on the device (behind Network.forward they are produced there): only
    # points and weights cross PCIe
    pts_only = dict(points_src=pinned["points_src"], points_ref=pinned["points_ref"], weights=pinned["weights"],
                    feat_src=devt["feat_src"], feat_ref=devt["feat_ref"])
    h2d_pts = sum(pinned[k].numel() * pinned[k].element_size() for k in ("points_src", "points_ref", "weights"))
    for _ in pipe.run(pts_only for _ in range(3)):
        pass
    ms_e2e_pts = statistics.median([e2e_block(pts_only) for _ in range(3)])

    clocks = sampler.stop() if rank == 0 else None   # sampled across the resident and the end-to-end timed loops

    # ---------------------------------------------------------------- N > 1: the row-block path (BASELINE configs[3])
    # ONE 131072 x 131072 D=64 pair, source rows sharded over the ranks, reference side replicated, one NCCL all_reduce of
    # the [B,17] fp64 moments per iteration (deepsir_b200/dist.py), the rank's iteration captured in a CUDA graph.
    rowblock = None
    if world > 1:
        rowblock = measure_rowblock(D, dist, dev, rank, world, ev, barrier)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, match_ms, filter_ms, knn_ms, ms_r5, ms_rand, ms_e2e_pts], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, match_ms, filter_ms, knn_ms, ms_r5, ms_rand, ms_e2e_pts = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    pairs = B * world * args.steps
    value = pairs / (ms / 1e3)
    flops = 2.0 * N_PTS * N_PTS * FEAT_D * B                      # algorithmic: 2*J*K*D per pair, B pairs/launch
    achieved = flops / (filter_ms / 1e3) / 1e12
    tc_peak = pk["bf16"]                                           # kind::f16 tcgen05 MMAs; 1.3 ms launches at full clocks: the burst figure
    traffic, traffic_src = filter_traffic_from_digest()
    knn_bytes = 6.31e6 * B                                         # 3.16 MB per cloud, two clouds per pair
    out = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "timed_blocks": {"blocks": len(blocks), "steps_per_block": args.steps, "ms_min": min(blocks), "ms_median": ms,
                            "ms_max": max(blocks), "note": "every block times exactly `steps` steps; value uses the median block"},
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "pairs_per_gpu": B, "registration_iters": 1,
                      "l2": "inputs larger than L2 (268 MB of features per step vs 126 MB L2), no explicit flush",
                      "sharding": "by pair, no collective",
                      "streams": "match + Kabsch on a high-priority stream, the KNN pyramids of a step on a forked stream joined one "
                                 "step later (every pyramid of a timed block finishes before its end event)"},
           "clocks": clocks,
           "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "h2d_gbs": h2d * args.steps / (ms_e2e / 1e3) / 1e9, "blocks": len(e2e_blocks),
                   "ms_blocks": [round(x, 2) for x in e2e_blocks],
                   "note": "upload-bound: the fp32 feature tensors of a step (268 MB) cross PCIe at the rate shown; a raw "
                           "pinned->device copy of the same bytes measures 55.3 GB/s on this pool (tools/e2e_probe.py)"},
           "e2e_points_only": {"value": pairs / (ms_e2e_pts / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_pts,
                               "d2h_bytes_per_step": d2h,
                               "note": "same host API, features already device resident (as behind Network.forward); points + "
                                       "weights up, transforms + correspondences down every step"},
           "gpu_launches": int(launches),
           "random_features": {"value": pairs / (ms_rand / 1e3), "unit": "pairs/s", "ms_per_step": ms_rand / args.steps,
                               "rescued_rows_per_step": int(n_rescued), "rows_per_step": B * N_PTS,
                               "note": "same step, un-planted random unit features (median top-2 gap ~0.04 instead of ~1): the "
                                       "data-dependent slow path / exact re-scoring of the filter is exercised"},
           "five_iterations": {"value": B * world / (ms_r5 / 1e3), "unit": "pairs/s", "ms_per_step": ms_r5,
                               "note": "same step with 5 registration iterations (the reference's default, arguments.py:69), "
                                       "device resident; iterations 2-5 hint the match filter with the previous correspondences"},
           "roofline": {"bound": "tensor", "kernel": "match_tc_filter_kernel (tcgen05 fp16 distance + row-argmin filter)",
                        "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak,
                        "frac_of_sustained": achieved / pk["bf16_sus"], "peak_sustained": pk["bf16_sus"],
                        "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": f"{pk['src']} bf16_tflops (burst: the timed region runs at full clocks, not under the "
                                       "power cap; 16-bit tensor-core inputs, fp32 accumulate; kernel timed inside the step by "
                                       "CUDA events on its stream); frac_of_sustained uses bf16_tflops_sustained",
                        "ms_per_launch": filter_ms, "match_call_ms": match_ms,
                        "match_call_frac": flops / (match_ms / 1e3) / 1e12 / tc_peak},
           "roofline_knn": {"bound": "hbm", "kernel": "knn pyramid (kd-ordered bucket-tree build + warp-cooperative queries, both clouds of a step)",
                            "achieved": knn_bytes / (knn_ms / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                            "frac": knn_bytes / (knn_ms / 1e3) / 1e9 / pk["hbm"], "ms_per_step": knn_ms,
                            "note": "HBM-bound by the scan/graph rule, but instruction bound in practice: ncu on the level-0 "
                                    "query kernel shows issue slots 54 % busy, DRAM < 2 % (profiles/ncu_digest_knn_tree_r2d.txt, "
                                    "profiles/knn_tree_r2.md); brute-force equivalent: 5.73 GFLOP per pair"}}
    if rowblock is not None:
        out["rowblock"] = rowblock
    if not args.no_cpu_baseline and world == 1:      # the CPU port is timed beside the N=1 run only
        torch.set_num_threads(os.cpu_count() or 1)
        n_pairs = 1
        cpu_step(host, n_pairs)                                    # warm (builds/loads the oracle lib, MKL init)
        t_0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t_0 < 10.0 and reps < 8):
            cpu_step(host, n_pairs)
            reps += 1
        dt = time.perf_counter() - t_0
        cores = os.cpu_count() or 1
        out["cpu_baseline"] = {"value": n_pairs * reps / dt, "unit": "pairs/s", "cores": cores, "kind": cpu_kind(),
                               "sample": cpu_sample_text(reps, n_pairs, cores)}
        c1 = c1_reference_forward()
        if c1 is not None:
            out["cpu_baseline"]["c1_forward_align_4"] = {
                "value": 1.0 / c1, "unit": "pairs/s", "s_per_pair": c1,
                "sample": "BASELINE configs[0]: the reference's Network.forward_align_4 (5 iterations, RandLA-Net + MLPs "
                          "included, seeded random-init weights) + kd-tree KNN pyramids on one synthetic 4096-point pair, CPU"}
        # "stock PyTorch on the same B200" (third column): reported beside the CPU port, never the product path
        torch.cuda.empty_cache()
        n_t = 8
        torch_gpu_step(devt, xs0, xr0, n_t)
        torch.cuda.synchronize()
        t_0 = time.perf_counter()
        for _ in range(3):
            torch_gpu_step(devt, xs0, xr0, n_t)
        torch.cuda.synchronize()
        dt_t = (time.perf_counter() - t_0) / 3
        torch.cuda.empty_cache()
        for _ in range(2):
            D.align_loop(devt["feat_src"], devt["feat_ref"], xs0, xr0, devt["weights"], 1)
        a0, a1 = ev(), ev()
        a0.record()
        for _ in range(5):
            D.align_loop(devt["feat_src"], devt["feat_ref"], xs0, xr0, devt["weights"], 1)
        a1.record()
        torch.cuda.synchronize()
        out["torch_gpu_baseline"] = {
            "value": n_t / dt_t, "unit": "pairs/s", "ours_same_scope": B * 5 / (a0.elapsed_time(a1) / 1e3),
            "scope": "match + gather + Kabsch + transform only (model.py:551-601): the reference has no GPU KNN",
            "sample": f"3 x {n_t} pairs of the C2 workload, the reference's statements (oracle port) on cuda tensors: cuBLAS "
                      "sgemm + torch.min in 6000-row chunks, fp64 SVD on the host as in the reference"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
