"""The iterative re-match / re-solve loop of Network.forward_align_4 (network/model.py:551-601)."""
from __future__ import annotations

import torch

from . import _lib as L
from . import kabsch as K
from . import match as M
from . import se3


def align_loop(feat_src, feat_ref, xyz_src, xyz_ref, weights, num_iter, feature_fn=None, weight_fn=None,
               want_pred=True, algo=L.MATCH_AUTO):
    """feat_* [B,C,N]; xyz_* [B,3,N] (the loop's layout, model.py:541-549); weights [B,J,1] or [B,J].
    Without callbacks the whole loop is ONE library call that enqueues every iteration on the current stream with
    no host synchronisation.  feature_fn(xyz_src) / weight_fn(xyz_src, xyz_ref_new) stand in for the two neural
    stages (self.aggregation :552, inlier_model :574-577) and switch to the per-iteration form.
    Returns (transforms: list of cumulative [B,3,4], pred_idx: list of int64 [B,J] or None, xyz_src_final [B,3,J],
    status int32 [iters,B])."""
    dev = L.require_cuda(feat_src, feat_ref, xyz_src, xyz_ref)
    B, C, J = feat_src.shape
    Kn = feat_ref.shape[2]
    if feature_fn is None and weight_fn is None:
        (fs, a), (fr, b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
        xs = xyz_src.contiguous().clone()
        xr = xyz_ref.contiguous()
        w = weights.reshape(B, J).contiguous()
        T = torch.empty(num_iter, B, 3, 4, dtype=torch.float32, device=dev)
        pred = torch.empty(num_iter, B, J, dtype=torch.int64, device=dev) if want_pred else None
        status = torch.empty(num_iter, B, dtype=torch.int32, device=dev)
        lib = L.lib()
        ws = L.workspace(lib.dsir_align_loop_workspace_bytes(B, C, J, Kn, algo), dev)
        L.check(lib.dsir_align_loop(fs, fr, B, C, J, Kn, xs.data_ptr(), xr.data_ptr(), w.data_ptr(), num_iter,
                                    T.data_ptr(), L.ptr(pred), status.data_ptr(), ws.data_ptr(), ws.numel(), algo,
                                    L.stream_ptr(dev)), "dsir_align_loop")
        return list(T.unbind(0)), (list(pred.unbind(0)) if want_pred else None), xs, status
    transforms, preds, stats = [], [], []
    for it in range(num_iter):
        fs = feature_fn(xyz_src) if feature_fn is not None else feat_src
        idx = M.match_argmin(fs, feat_ref, algo=algo, prior=preds[-1] if preds else None)   # :558-569 (hinted by the last match)
        if weight_fn is not None:
            w = weight_fn(xyz_src, M.gather_neighbour_V3(xyz_ref, idx))             # :571-577
        else:
            w = weights
        T, st = K.kabsch_gather(xyz_src, xyz_ref, idx, w)                           # :571,:586-588
        xyz_src = se3.transform_V2(T, xyz_src)                                      # :590-591
        transforms.append(T if it == 0 else se3.concatenate(T, transforms[-1]))     # :595
        preds.append(idx)
        stats.append(st)
    return transforms, preds, xyz_src, torch.stack(stats)


def pred_pairs(indexs):
    """network/model.py:599-601: [B,J] int64 device indices -> [B,J,2] int32 on the CPU (forces the only sync)."""
    B, J = indexs.shape
    i0 = torch.arange(J)[None, :].expand(B, J).int()[:, :, None]
    return torch.cat([i0, indexs.int().cpu()[:, :, None]], dim=2)
