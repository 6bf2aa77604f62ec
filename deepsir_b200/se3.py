"""SE(3) helpers with the signatures of common/math/se3_torch.py, on the device."""
from __future__ import annotations

import torch

from . import _lib as L


def identity(batch_size, device=None):
    """se3_torch.py:6-7 (the reference builds it on the CPU; `device` is an extension)."""
    return torch.eye(3, 4, device=device)[None, ...].repeat(batch_size, 1, 1)


def _t34(Rt):
    if Rt.dtype != torch.float32:
        raise L.DeepSIRError("transforms must be float32")
    if Rt.dim() == 2:
        Rt = Rt[None]
    if Rt.stride(-1) != 1 or Rt.stride(-2) != 4:
        Rt = Rt[..., :3, :].contiguous()
    return Rt


def inverse(Rt):
    """se3_torch.py:10-25."""
    dev = L.require_cuda(Rt)
    Rt = _t34(Rt)
    B = Rt.shape[0]
    out = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_se3_inverse(Rt.data_ptr(), Rt.stride(0), B, out.data_ptr(), L.stream_ptr(dev)), "dsir_se3_inverse")
    return out


def concatenate(a, b):
    """se3_torch.py:28-48: a o b."""
    dev = L.require_cuda(a, b)
    a, b = _t34(a), _t34(b)
    B = a.shape[0]
    out = torch.empty(B, 3, 4, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_se3_compose(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), B, out.data_ptr(),
                                     L.stream_ptr(dev)), "dsir_se3_compose")
    return out


def _apply(Rt, pts_desc, B, N, out, o_bs, o_ps, o_cs, rotate_only, dev):
    tbs = Rt.stride(0) if Rt.shape[0] == B else 0  # ([1,] 3/4, 4) broadcasts over the batch
    L.check(L.lib().dsir_se3_apply(Rt.data_ptr(), tbs, pts_desc, B, N, out.data_ptr(), o_bs, o_ps, o_cs,
                                   int(rotate_only), L.stream_ptr(dev)), "dsir_se3_apply")


def transform(Rt, a, normals=None):
    """se3_torch.py:51-77: a [B,N,3] (or [N,3]) -> a R^T + t."""
    dev = L.require_cuda(Rt, a)
    if len(Rt.size()) != len(a.size()):
        raise NotImplementedError
    Rt = _t34(Rt)
    squeeze = a.dim() == 2
    a3 = a[None] if squeeze else a
    B, N, _ = a3.shape
    out = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    _apply(Rt, L.points_bm3(a3), B, N, out, N * 3, 3, 1, False, dev)
    if normals is not None:
        n3 = normals[None] if squeeze else normals
        on = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
        _apply(Rt, L.points_bm3(n3), B, N, on, N * 3, 3, 1, True, dev)
        return (out[0], on[0]) if squeeze else (out, on)
    return out[0] if squeeze else out


def transform_V2(Rt, a, normals=None):
    """se3_torch.py:80-100: a [B,3,N] -> R a + t."""
    dev = L.require_cuda(Rt, a)
    Rt = _t34(Rt)
    B, _, N = a.shape
    out = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
    _apply(Rt, L.points_b3m(a), B, N, out, 3 * N, 1, N, False, dev)
    if normals is not None:
        on = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        _apply(Rt, L.points_b3m(normals), B, N, on, 3 * N, 1, N, True, dev)
        return out, on
    return out
