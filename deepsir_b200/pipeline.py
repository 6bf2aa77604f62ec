"""Host-side pipeline over the C-ABI path: pinned host batches in, poses and correspondences out.

The reference moves a collated batch to the device (`dict_all_to_device`, test.py:391), runs the forward pass and reads
the pose back, strictly in sequence.  Here the upload of batch i+1 (copy stream) overlaps the kernels of batch i
(compute stream) and the small result download, so a stream of batches runs at max(PCIe time, kernel time) per batch
instead of their sum.  The KNN pyramids of a batch depend on nothing but its points (in the reference they are built in
the DataLoader workers): they run on a third stream, next to the match / Kabsch loop of the same and of the previous
batch, and are joined on the host when the batch is handed out.  Everything on the device goes through libdeepsir_b200.so; torch only provides streams, events
and memory.
"""
from __future__ import annotations

from collections import deque

import torch

from . import _lib as L
from .knn import nn_search_pair
from .loop import align_loop


class RegistrationPipeline:
    """KNN pyramid of both clouds + `iters` x (match -> Kabsch -> transform) per batch, double buffered.

    A batch is a dict of PINNED host tensors: points_src/points_ref [B,N,>=3], feat_src/feat_ref [B,C,N],
    weights [B,N].  `run(batches)` yields, in order, dicts with host tensors
    T [B,3,4] (final cumulative transform), pred [B,N] int32 (last correspondences, the `pred_pairs` of
    network/model.py:599-601) and status [iters,B]; with keep_graph=True also the device KNN tensors of the batch.
    The host tensors of a yielded dict come from a ring of pinned buffers and are re-used depth + 2 batches later.
    A batch may carry features that are ALREADY on the device (feat_src / feat_ref as CUDA tensors - behind
    Network.forward they are produced there): only the host tensors are uploaded.
    """

    def __init__(self, device=None, num_knn=16, sub_sampling_ratio=(4, 4, 4, 4), iters=1, depth=2, keep_graph=False):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.dev.type != "cuda":
            raise L.DeepSIRError("RegistrationPipeline runs on a CUDA device only")
        self.k, self.ratios, self.iters, self.depth, self.keep_graph = num_knn, tuple(sub_sampling_ratio), iters, depth, keep_graph
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.knn_stream = torch.cuda.Stream(self.dev)
        # pinned result buffers: a ring of depth + 2 sets, allocated once per shape (cudaHostAlloc is far too slow for the
        # steady-state loop).  A yielded set is overwritten depth + 2 batches later: copy what must outlive that.
        self._ring, self._ring_shape, self._ring_pos = [], None, 0

    def _host_out(self, B, N):
        if self._ring_shape != (B, N):
            self._ring = [dict(T=torch.empty(B, 3, 4, dtype=torch.float32, pin_memory=True),
                               pred=torch.empty(B, N, dtype=torch.int32, pin_memory=True),
                               status=torch.empty(self.iters, B, dtype=torch.int32, pin_memory=True))
                          for _ in range(self.depth + 2)]
            self._ring_shape, self._ring_pos = (B, N), 0
        out = dict(self._ring[self._ring_pos])
        self._ring_pos = (self._ring_pos + 1) % len(self._ring)
        return out

    def _upload(self, host, compute):
        with torch.cuda.stream(self.copy_stream):
            d = {k: (v if v.is_cuda else v.to(self.dev, non_blocking=True)) for k, v in host.items()}
            up = torch.cuda.Event()
            up.record(self.copy_stream)
        for k, v in d.items():
            if not host[k].is_cuda:
                v.record_stream(compute)   # the compute stream reads them: keep the allocator from recycling early
                if k.startswith("points"):
                    v.record_stream(self.knn_stream)
        return d, up

    def _compute(self, d, up, compute):
        compute.wait_event(up)
        self.knn_stream.wait_event(up)
        with torch.cuda.stream(self.knn_stream):
            g_src, g_ref = nn_search_pair(d["points_src"], d["points_ref"], self.k, self.ratios)
            knn_done = torch.cuda.Event()
            knn_done.record(self.knn_stream)
        xs = d["points_src"][:, :, :3].permute(0, 2, 1).contiguous()   # the loop's [B,3,N] layout (model.py:541-549)
        xr = d["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
        tr, pred, _, status = align_loop(d["feat_src"], d["feat_ref"], xs, xr, d["weights"], self.iters)
        B, N = pred[-1].shape
        out = self._host_out(B, N)
        out["T"].copy_(tr[-1], non_blocking=True)
        out["pred"].copy_(pred[-1].to(torch.int32), non_blocking=True)
        out["status"].copy_(status, non_blocking=True)
        if self.keep_graph:
            out["graph_src"], out["graph_ref"] = g_src, g_ref
            # produced on the KNN stream, consumed by the caller on the compute stream: tell the caching allocator, so that a
            # tensor dropped while the caller's kernels still read it is not recycled under them
            for g in (g_src, g_ref):
                for lst in g.values():
                    for t in (lst if isinstance(lst, (list, tuple)) else [lst]):
                        if torch.is_tensor(t):
                            t.record_stream(compute)
        done = torch.cuda.Event()
        done.record(compute)
        return out, (done, knn_done)

    def run(self, batches):
        compute = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(compute)        # uploads start after whatever the caller enqueued before
        self.knn_stream.wait_stream(compute)         # (device-resident inputs were produced there)
        uploaded, inflight = deque(), deque()
        it = iter(batches)

        def feed():
            try:
                uploaded.append(self._upload(next(it), compute))
                return True
            except StopIteration:
                return False

        for _ in range(self.depth):
            if not feed():
                break
        while uploaded:
            d, up = uploaded.popleft()
            inflight.append(self._compute(d, up, compute))
            del d
            feed()                                   # the next upload is enqueued while this batch computes
            if len(inflight) >= self.depth:
                out, events = inflight.popleft()
                for e in events:
                    e.synchronize()
                yield out
        while inflight:
            out, events = inflight.popleft()
            for e in events:
                e.synchronize()
            yield out
