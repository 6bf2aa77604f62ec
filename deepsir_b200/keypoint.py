"""Key-point scoring and top-k selection, mirroring Network.score_fun / Network.feat_score (network/model.py:668-757):
same argument order, shapes and return values, executed by libdeepsir_b200.so."""
from __future__ import annotations

import torch

from . import _lib as L
from .match import gather_neighbour_V3

# network/model.py:146-149 (semantic-KITTI label weights of the shipped configuration)
KITTI_LABEL_WEIGHTS = (3, 1, 1, 3, 2, 0, 0, 0, 6, 5, 6, 4, 7, 7, 6, 8, 4, 9, 9)


def score_fun(feat, xyz, prob, label, neigh_idx, label_weights=KITTI_LABEL_WEIGHTS, k_neighbors=16, ball_r=2.0):
    """network/model.py:700-757.  feat [B,C,N], xyz [B,3,N], prob [B,1,N], label [B,1,N], neigh_idx [B,N,k>=16] -> [B,N]."""
    dev = L.require_cuda(feat, xyz, neigh_idx)
    B, C, N = feat.shape
    if neigh_idx.shape[1] != N or neigh_idx.shape[2] < k_neighbors:
        raise L.DeepSIRError("score_fun: neigh_idx must be [B, N, >= k_neighbors]")
    feat, xyz = feat.contiguous(), xyz.contiguous()
    idx = neigh_idx if (neigh_idx.dtype == torch.int64 and neigh_idx.is_contiguous()) else neigh_idx.to(torch.int64).contiguous()
    p = None if prob is None else prob.reshape(B, N).to(torch.float32).contiguous()
    lab = None if label is None else label.reshape(B, N).to(torch.int64).contiguous()
    lw = torch.as_tensor(label_weights, dtype=torch.float32, device=dev).contiguous()
    score = torch.empty(B, N, dtype=torch.float32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_keypoint_score_workspace_bytes(B, C, N), dev)
    L.check(lib.dsir_keypoint_score(feat.data_ptr(), xyz.data_ptr(), L.ptr(p), L.ptr(lab), lw.data_ptr(), lw.numel(),
                                    idx.data_ptr(), idx.shape[2], k_neighbors, ball_r, B, C, N, score.data_ptr(), ws.data_ptr(),
                                    ws.numel(), L.stream_ptr(dev)), "dsir_keypoint_score")
    return score


def topk(score, k):
    """torch.topk(score, k, dim=-1, largest=True) (network/model.py:692) with ties resolved to the lower index."""
    dev = L.require_cuda(score)
    B, N = score.shape
    score = score.to(torch.float32).contiguous()
    values = torch.empty(B, k, dtype=torch.float32, device=dev)
    index = torch.empty(B, k, dtype=torch.int64, device=dev)
    L.check(L.lib().dsir_topk_rows(score.data_ptr(), B, N, k, values.data_ptr(), index.data_ptr(), L.stream_ptr(dev)),
            "dsir_topk_rows")
    return values, index


def feat_score(feat, xyz, prob, label, neigh_idx, num_sub=0, sub_selection=None, label_weights=KITTI_LABEL_WEIGHTS):
    """network/model.py:668-698: score every point, optionally keep the num_sub best and gather their xyz / features /
    labels.  Returns (feat, xyz, label, score) like the reference."""
    num_points = xyz.shape[2]
    neigh_idx = neigh_idx[:, 0:num_points, :]
    score = score_fun(feat, xyz, prob, label, neigh_idx, label_weights)
    if sub_selection is None:
        sub_selection = num_sub > 0
    if sub_selection:
        if not 0 < num_sub <= num_points:
            raise AssertionError("0 < num_sub <= num_points")
        score, index = topk(score, num_sub)
        xyz = gather_neighbour_V3(xyz, index)
        feat = gather_neighbour_V3(feat, index)
        label = gather_neighbour_V3(label.to(torch.float32), index).to(label.dtype)
    return feat, xyz, label, score
