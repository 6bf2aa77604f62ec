"""Row-block sharding of ONE large pair across GPUs (BASELINE config 4).

The path shards by pair with no communication at all (bench.py --gpus N).  When a single pair is too large for that
(131072 x 131072), the SOURCE rows are split into contiguous blocks, one per rank; every rank holds a replica of the
reference cloud (features + xyz: 33.5 MB + 1.5 MB at C4).  Per iteration of the loop (network/model.py:551-601):

    idx_local = argmin match(feat_src[rows], feat_ref)                  local, no exchange
    mom_local = raw fp64 moments {S|w|, Sw, Swx, Swy, Swxy} of my rows  local  (dsir_kabsch_moments, gather fused)
    mom       = all_reduce(SUM, mom_local)                              17 doubles per pair: the only collective
    T         = Kabsch from moments                                     identical on every rank (deterministic)
    xyz_src[rows] <- T xyz_src[rows];  T_total <- T o T_total           local

Raw (uncentred) moments are additive, and H = S_xy/S - (2 - S_w/S) c_s c_t^T reproduces the reference's centred
covariance algebraically; fp64 keeps the cancellation far below the fp32 inputs' resolution.

`ops` abstracts the five device operations so that the exchange logic is exercised on CPU with `gloo` and the oracle
standing in for the kernels (tests/test_dist_cpu.py); the default is the CUDA library.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def row_block(J: int, world: int, rank: int):
    """Contiguous, balanced split of J source rows: rows [lo, hi) belong to `rank`."""
    return (rank * J) // world, ((rank + 1) * J) // world


class LibraryOps:
    """The CUDA path (libdeepsir_b200.so)."""

    @staticmethod
    def match_argmin(feat_src, feat_ref):
        from .match import match_argmin
        return match_argmin(feat_src, feat_ref)

    @staticmethod
    def moments(xyz_src, xyz_ref, idx, weights):
        """xyz_* [B,3,*]; idx [B,Jl] int64; weights [B,Jl] -> [B,17] fp64 with tgt_j = xyz_ref[:, :, idx_j]."""
        from .kabsch import kabsch_moments
        return kabsch_moments(xyz_src, xyz_ref, weights, gather=idx, layout="b3m")

    @staticmethod
    def solve(moments):
        from .kabsch import kabsch_from_moments
        return kabsch_from_moments(moments)

    @staticmethod
    def transform(T, xyz):
        from .se3 import transform_V2
        return transform_V2(T, xyz)

    @staticmethod
    def compose(a, b):
        from .se3 import concatenate
        return concatenate(a, b)


def all_reduce_moments(mom: torch.Tensor, group=None) -> torch.Tensor:
    """SUM over ranks of the [B,17] fp64 partial moments (136 bytes per pair; latency bound).  NCCL on GPUs, gloo on CPU."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mom, op=dist.ReduceOp.SUM, group=group)
    return mom


def gather_rows(idx_local: torch.Tensor, J: int, group=None) -> torch.Tensor:
    """All-gather of the per-rank correspondence blocks [B,Jl] into [B,J] (only when the caller wants pred_pairs)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return idx_local
    world = dist.get_world_size(group)
    B = idx_local.shape[0]
    width = max(row_block(J, world, r)[1] - row_block(J, world, r)[0] for r in range(world))
    pad = torch.zeros(B, width, dtype=idx_local.dtype, device=idx_local.device)
    pad[:, :idx_local.shape[1]] = idx_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty(B, J, dtype=idx_local.dtype, device=idx_local.device)
    for r in range(world):
        lo, hi = row_block(J, world, r)
        out[:, lo:hi] = parts[r][:, :hi - lo]
    return out


def align_rowblock(feat_src_local, feat_ref, xyz_src_local, xyz_ref, weights_local, num_iter, group=None, ops=None,
                   gather_pred_rows: int = 0):
    """The loop of forward_align_4 for this rank's source rows.  feat_src_local [B,C,Jl], feat_ref [B,C,K] (replica),
    xyz_src_local [B,3,Jl], xyz_ref [B,3,K] (replica), weights_local [B,Jl].
    Returns (transforms: list of cumulative [B,3,4], identical on every rank; pred: list of [B,Jl] local indices, or
    [B,J] gathered when gather_pred_rows=J; xyz_src_local after the last iteration; status list)."""
    ops = ops or LibraryOps
    transforms, preds, stats = [], [], []
    xyz = xyz_src_local
    w = weights_local.reshape(weights_local.shape[0], -1)
    for it in range(num_iter):
        idx = ops.match_argmin(feat_src_local, feat_ref)                        # model.py:558-569, rows of this rank
        mom = all_reduce_moments(ops.moments(xyz, xyz_ref, idx, w), group)      # the only exchange
        T, st = ops.solve(mom)                                                  # model.py:588, same T everywhere
        xyz = ops.transform(T, xyz)                                             # :590
        transforms.append(T if it == 0 else ops.compose(T, transforms[-1]))     # :595
        preds.append(gather_rows(idx, gather_pred_rows, group) if gather_pred_rows else idx)
        stats.append(st)
    return transforms, preds, xyz, stats


class GraphedRowBlock:
    """`num_iter` iterations of align_rowblock for FIXED shapes with the rank's kernels captured in CUDA graphs: per
    iteration ONE graph for match + gather + moments and ONE for solve + transform + compose, the NCCL all_reduce of the
    [B,17] moments issued eagerly between the two (collectives stay outside the graphs: a graph that holds NCCL kernels
    was measured to stall the process-group teardown for minutes).  A 16384-row block is 0.3 ms of match work per rank at
    C4 on 8 GPUs - the ~20 small launches of an eager iteration no longer fit beside it.  Inputs are static copies;
    `step()` returns (T [B,3,4] cumulative, idx [B,Jl] of the last iteration, xyz_src_local) as static tensors.  Falls
    back to eager execution when the capture is refused (`self.graphed` says which)."""

    def __init__(self, feat_src_local, feat_ref, xyz_src_local, xyz_ref, weights_local, num_iter=1, group=None, ops=None):
        self.fs, self.fr, self.xyz0, self.xr, self.w = [t.detach().clone().contiguous() for t in
                                                        (feat_src_local, feat_ref, xyz_src_local, xyz_ref, weights_local)]
        self.w = self.w.reshape(self.w.shape[0], -1)
        self.num_iter, self.group, self.ops = num_iter, group, ops or LibraryOps
        dev = self.fs.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):            # warm-up outside the capture (NCCL channels, stream pools, attributes)
            for _ in range(3):
                self.out = self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graphed, self.segA, self.segB, self.mom = False, [], [], []
        try:
            pool = torch.cuda.graph_pool_handle()
            xyz, Tprev = None, None
            for it in range(num_iter):
                gA = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gA, pool=pool):
                    x_in = self.xyz0.clone() if it == 0 else xyz
                    idx = self.ops.match_argmin(self.fs, self.fr)
                    mom = self.ops.moments(x_in, self.xr, idx, self.w)
                gB = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gB, pool=pool):
                    T, st = self.ops.solve(mom)
                    xyz = self.ops.transform(T, x_in)
                    Tprev = T if it == 0 else self.ops.compose(T, Tprev)
                self.segA.append(gA); self.segB.append(gB); self.mom.append(mom)
            self._static = (Tprev, idx, xyz)
            self.graphed = True
        except Exception as e:                   # noqa: BLE001 - report and run eagerly
            self.error = repr(e)
            self.graphed = False
            torch.cuda.synchronize(dev)

    def _eager(self):
        tr, pred, xyz, st = align_rowblock(self.fs, self.fr, self.xyz0, self.xr, self.w, self.num_iter, group=self.group,
                                           ops=self.ops)
        return tr[-1], pred[-1], xyz

    def step(self):
        if not self.graphed:
            self.out = self._eager()
            return self.out
        for gA, gB, mom in zip(self.segA, self.segB, self.mom):
            gA.replay()
            all_reduce_moments(mom, self.group)   # in place on the static tensor, ordered on the current stream
            gB.replay()
        self.out = self._static
        return self.out
