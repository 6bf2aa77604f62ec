"""On-device evaluation of a registration result: the host-side steps that follow the forward pass in the reference
(network/loss.py:723-749, common/metrics_util.py:27-85, test.py:432-441), executed by libdeepsir_b200.so without a
device->host round trip per batch."""
from __future__ import annotations

import torch

from . import _lib as L
from . import se3


def pose_errors(pred_transforms, gt_transforms, rte_thresh=2.0, rre_thresh=5.0):
    """T_pred, T_gt [B,3,4] -> dict(err_r_deg, err_t, succ) of compute_metrics (metrics_util.py:55-63) and (rre, rte) of
    rte_rre (metrics_util.py:27-33)."""
    dev = L.require_cuda(pred_transforms, gt_transforms)
    B = pred_transforms.shape[0]
    tp = pred_transforms[:, :3, :].to(torch.float32).contiguous()
    tg = gt_transforms[:, :3, :].to(torch.float32).contiguous()
    out = torch.empty(B, 4, dtype=torch.float32, device=dev)
    succ = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(L.lib().dsir_pose_errors(tp.data_ptr(), tg.data_ptr(), B, rte_thresh, rre_thresh, out.data_ptr(), succ.data_ptr(),
                                     L.stream_ptr(dev)), "dsir_pose_errors")
    return dict(err_r_deg=out[:, 0], err_t=out[:, 1], rre=out[:, 2], rte=out[:, 3], succ=succ.bool())


def find_correct_correspondence(pos_pairs, pred_pairs, hash_seed=None, len_batch=None):
    """Loss.find_correct_correspondence (network/loss.py:723-749).  pos_pairs: list of [N_i',2] integer tensors,
    pred_pairs [B,N,2] -> bool [B,N] on the device of pred_pairs."""
    assert len(pos_pairs) == len(pred_pairs)
    dev = L.require_cuda(pred_pairs)
    B, N, _ = pred_pairs.shape
    if hash_seed is None:
        assert len(len_batch) == len(pos_pairs)
        seeds = [int(max(n0, n1)) for n0, n1 in len_batch]
    else:
        seeds = [int(hash_seed)] * B
    pos = [torch.as_tensor(p).to(device=dev, dtype=torch.int32).reshape(-1, 2) for p in pos_pairs]
    sizes = [int(p.shape[0]) for p in pos]
    offsets = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0).tolist()), dtype=torch.int64, device=dev)
    total = int(sum(sizes))
    cat = torch.cat(pos, 0).contiguous() if total > 0 else torch.zeros(1, 2, dtype=torch.int32, device=dev)
    pred = pred_pairs.to(torch.int32).contiguous()
    seed_t = torch.tensor(seeds, dtype=torch.int64, device=dev)
    correct = torch.empty(B, N, dtype=torch.uint8, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_correspondence_check_workspace_bytes(total), dev)
    L.check(lib.dsir_correspondence_check(cat.data_ptr(), offsets.data_ptr(), total, pred.data_ptr(), B, N, seed_t.data_ptr(),
                                          correct.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(dev)),
            "dsir_correspondence_check")
    return correct.bool()


def nn_sqdist(a, b):
    """a [B,N,3], b [B,M,3] -> (min_k |a_j - b_k|^2 [B,N], its mean over j [B]) by direct differences
    (metrics_util.py:38-40, 72-74)."""
    dev = L.require_cuda(a, b)
    B, N, _ = a.shape
    M = b.shape[1]
    a, b = a[:, :, :3].to(torch.float32).contiguous(), b[:, :, :3].to(torch.float32).contiguous()
    min_d = torch.empty(B, N, dtype=torch.float32, device=dev)
    mean = torch.empty(B, dtype=torch.float32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_nn_sqdist_workspace_bytes(B), dev)
    L.check(lib.dsir_nn_sqdist_mean(a.data_ptr(), b.data_ptr(), B, N, M, min_d.data_ptr(), mean.data_ptr(), ws.data_ptr(),
                                    ws.numel(), L.stream_ptr(dev)), "dsir_nn_sqdist_mean")
    return min_d, mean


def compute_metrics(data, pred_transforms, rte_thresh, rre_thresh):
    """The device part of compute_metrics (common/metrics_util.py:36-85): err_r_deg, err_t, succ, chamfer_dist.
    (The Euler-angle r_mse / r_mae of the reference go through scipy on the host and are not part of the path.)"""
    gt = data["transform_gt"]
    points_src = data["points_src"][:, :2048, :3]
    points_ref = data["points_ref"][:, :2048, :3]
    if "points_raw" in data:
        points_raw = data["points_raw"][..., :3]
    else:
        points_raw = torch.cat([se3.transform(gt, points_src), points_ref], dim=1)
    pe = pose_errors(pred_transforms, gt, rte_thresh, rre_thresh)
    src_transformed = se3.transform(pred_transforms, points_src)
    inter = se3.concatenate(pred_transforms, se3.inverse(gt))
    src_clean = se3.transform(inter, points_raw)
    _, m_src = nn_sqdist(src_transformed, points_raw)
    _, m_ref = nn_sqdist(points_ref, src_clean)
    return dict(err_r_deg=pe["err_r_deg"], err_t=pe["err_t"], succ=pe["succ"], chamfer_dist=m_src + m_ref)
