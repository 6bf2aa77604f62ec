"""Weighted Kabsch, mirroring compute_rigid_transform_2 / compute_rigid_transform of network/model.py."""
from __future__ import annotations

import torch

from . import _lib as L


def _solve(src_p, tgt_p, w, w_bs, gather, B, M, dev, want_moments=False, n_tgt=0):
    T = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    mom = torch.empty(B, 17, dtype=torch.float64, device=dev) if want_moments else None
    lib = L.lib()
    ws = L.workspace(lib.dsir_kabsch_workspace_bytes(B, M), dev)
    L.check(lib.dsir_kabsch(src_p, tgt_p, L.ptr(w), w_bs, L.ptr(gather), B, M, int(n_tgt), T.data_ptr(), status.data_ptr(),
                            L.ptr(mom), ws.data_ptr(), ws.numel(), L.stream_ptr(dev)), "dsir_kabsch")
    return T, status, mom


class LazyFlag:
    """`invalid_gradient` of the reference (network/model.py:61-64) without forcing a host sync: the device
    status is only read when the flag is actually tested."""

    def __init__(self, status):
        self.status = status

    def __bool__(self):
        return bool(self.status.any().item())

    def __or__(self, other):
        return bool(self) or bool(other)

    __ror__ = __or__


def compute_rigid_transform_2(src, tgt, weights, return_status=False):
    """network/model.py:22-66.  src [B,M,3], tgt [B,M,3], weights [B,M,1] -> (T [B,3,4], invalid_gradient).
    No host round trip: moments, fp64 3x3 SVD and the determinant fix run on the device."""
    dev = L.require_cuda(src, tgt, weights)
    B, M, _ = src.shape
    w = weights.reshape(B, M)
    if w.stride(1) != 1:
        w = w.contiguous()
    T, status, _ = _solve(L.points_bm3(src), L.points_bm3(tgt), w, w.stride(0), None, B, M, dev)
    return (T, status) if return_status else (T, LazyFlag(status))


def kabsch_gather(xyz_src, xyz_ref, indexs, weights):
    """Fused network/model.py:571 + :586-588: xyz_src [B,3,J], xyz_ref [B,3,K], indexs [B,J] int64, weights [B,J(,1)]
    -> (T [B,3,4], status int32 [B]) with tgt_j = xyz_ref[:, :, indexs_j] gathered inside the reduction."""
    dev = L.require_cuda(xyz_src, xyz_ref, indexs, weights)
    B, _, J = xyz_src.shape
    w = weights.reshape(B, J)
    if w.stride(1) != 1:
        w = w.contiguous()
    T, status, _ = _solve(L.points_b3m(xyz_src), L.points_b3m(xyz_ref), w, w.stride(0), indexs.contiguous(), B, J, dev,
                          n_tgt=xyz_ref.shape[2])
    return T, status


def kabsch_moments(src, tgt, weights, gather=None, layout="bm3"):
    """Additive fp64 raw moments [B,17] of this rank's rows (row-block sharding)."""
    dev = L.require_cuda(src, tgt, weights)
    pts = L.points_bm3 if layout == "bm3" else L.points_b3m
    B = src.shape[0]
    M = src.shape[1] if layout == "bm3" else src.shape[2]
    w = weights.reshape(B, M)
    if w.stride(1) != 1:
        w = w.contiguous()
    mom = torch.empty(B, 17, dtype=torch.float64, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_kabsch_workspace_bytes(B, M), dev)
    g = gather.contiguous() if gather is not None else None
    n_tgt = (tgt.shape[1] if layout == "bm3" else tgt.shape[2]) if g is not None else 0
    L.check(lib.dsir_kabsch_moments(pts(src), pts(tgt), w.data_ptr(), w.stride(0), L.ptr(g), B, M, n_tgt, mom.data_ptr(),
                                    ws.data_ptr(), ws.numel(), L.stream_ptr(dev)), "dsir_kabsch_moments")
    return mom


def kabsch_from_moments(moments):
    """Second half of the solve from (all-reduced) moments [B,17] fp64 -> (T [B,3,4], status [B])."""
    dev = L.require_cuda(moments)
    moments = moments.contiguous()
    B = moments.shape[0]
    T = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(L.lib().dsir_kabsch_from_moments(moments.data_ptr(), B, T.data_ptr(), status.data_ptr(), L.stream_ptr(dev)),
            "dsir_kabsch_from_moments")
    return T, status


def compute_rigid_transform(src, tgt, weights):
    """network/model.py:68-116 (soft).  src [B,M,3], tgt [B,N,3], weights [B,M,N] -> (T, invalid_gradient).
    The [B,M,N] weights are an input of this signature: one pass over them (dsir_soft_targets) gives the row masses and
    the soft targets; the fused path that never forms [M,N] is match_soft + kabsch_soft."""
    dev = L.require_cuda(src, tgt, weights)
    B, M, N = weights.shape
    if weights.dtype != torch.float32:
        raise L.DeepSIRError("compute_rigid_transform expects float32 weights")
    w = weights if weights.stride(2) == 1 else weights.contiguous()
    tg = tgt[:, :, :3].contiguous()
    y = torch.empty(B, M, 3, dtype=torch.float32, device=dev)
    mass = torch.empty(B, M, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_soft_targets(w.data_ptr(), w.stride(0), w.stride(1), tg.data_ptr(), B, M, N, y.data_ptr(),
                                      mass.data_ptr(), L.stream_ptr(dev)), "dsir_soft_targets")
    return kabsch_soft(src, y, mass)


def kabsch_soft(src, y_soft, rowmass):
    """Soft Kabsch from fused soft targets: src [B,M,3], y_soft [B,M,3], rowmass [B,M] -> (T, invalid)."""
    dev = L.require_cuda(src, y_soft, rowmass)
    B, M, _ = src.shape
    y_soft, rowmass = y_soft.contiguous(), rowmass.contiguous()
    T = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_kabsch_workspace_bytes(B, M), dev)
    L.check(lib.dsir_kabsch_soft(L.points_bm3(src), y_soft.data_ptr(), rowmass.data_ptr(), B, M, T.data_ptr(),
                                 status.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(dev)), "dsir_kabsch_soft")
    return T, LazyFlag(status)
