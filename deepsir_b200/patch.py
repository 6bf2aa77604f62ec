"""Drop the library into an importable copy of the reference (LeoQLi/DeepSIR) without editing its files.

`patch()` rebinds, inside the already imported reference modules, exactly the names the hot path goes through
(SURVEY 8 b) and nothing else; the RandLA-Net bodies, MLPs, losses and loaders stay the reference's own code:

  network.model.match_features_V2            -> deepsir_b200.match_features_V2          (network/matchnet.py:116-144)
  network.model.gather_neighbour_V3          -> deepsir_b200.gather_neighbour_V3        (network/tools.py:211-221)
  network.model.compute_rigid_transform_2    -> deepsir_b200.compute_rigid_transform_2  (network/model.py:22-66)
  network.model.compute_rigid_transform      -> deepsir_b200.compute_rigid_transform    (network/model.py:68-116)
  network.model.se3_torch                    -> deepsir_b200.se3_torch                  (common/math/se3_torch.py)
  network.model.Network.forward_align_4      -> forward_align_4 below (level="loop", default): the loop of
                                                network/model.py:551-601 with the chunked matrix + argmin of :558-569
                                                replaced by ONE fused call; same inputs, outputs and endpoints
  <data_base module>.Util.knn                -> deepsir_b200.knn  (`patch_knn`, dataloader/data_base.py:13,165,170)

level="leaf" rebinds only the leaf functions (the reference's own loop then materialises every 6000-row block through
`match_features_V2` and runs `torch.min` on it) — the slower, signature-for-signature route.  `unpatch()` restores
everything.  CUDA tensors only: there is no CPU fallback, on CPU inputs the rebound functions raise DeepSIRError.
"""
from __future__ import annotations

import importlib
import sys

import torch

from . import kabsch as _K
from .knn import knn as _knn_fn
from . import match as _M
from . import se3 as _se3

_saved = {}


def _import_reference(ref_root=None):
    if ref_root is not None and ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    return importlib.import_module("network.model")


def forward_align_4(self, data, opt=None):
    """Network.forward_align_4 (network/model.py:520-607) over the fused kernels.  `self` is the reference's Network:
    forward_pair, aggregation and inlier_model are its own modules; returns the same (transforms, endpoints)."""
    num_reg_iter, clip_weight = opt
    src_xyz_multi = data['points_src_xyz']
    src_neigh_idx = data['points_src_neigh_idx']
    src_sub_idx = data['points_src_sub_idx']
    src_interp_idx = data['points_src_interp_idx']

    feat_src_0, xyz_src, label_src, score_src, feat_ref_0, xyz_ref, label_ref, score_ref = self.forward_pair(data)

    endpoints = {}
    endpoints['pt_src'] = xyz_src.permute(0, 2, 1).contiguous()
    endpoints['pt_ref'] = xyz_ref.permute(0, 2, 1).contiguous()

    transforms, all_matrices, all_pred_pairs, flags = [], [], [], []
    indexs = None
    xyz_ref_new = None
    for it in range(num_reg_iter):
        feat_src, feat_ref = self.aggregation(xyz_src, xyz_ref, feat_src_0, feat_ref_0,
                                              label_src, label_ref, score_src, score_ref)
        with torch.no_grad():
            # model.py:558-569 — the 6000-row chunks, the [B, stride, K] matrix and its argmin are one fused call; the
            # previous iteration's correspondences only speed the filter up, the result does not depend on them
            indexs = _M.match_argmin(feat_src, feat_ref, prior=indexs)
        xyz_ref_new = _M.gather_neighbour_V3(xyz_ref, indexs)                                   # :571
        cat_xyz = torch.cat((xyz_src, xyz_ref_new), dim=1).permute(0, 2, 1).contiguous()        # :574
        _, _, logit = self.inlier_model(cat_xyz, src_xyz_multi, src_neigh_idx, src_sub_idx, src_interp_idx)
        logit = logit.squeeze(dim=1)
        weights = logit.sigmoid()[:, :, None]                                                  # :577
        xyz_src = xyz_src.permute(0, 2, 1).contiguous()                                         # :586
        xyz_ref_new = xyz_ref_new.permute(0, 2, 1).contiguous()                                 # :587
        R_t, flag = _K.compute_rigid_transform_2(xyz_src, xyz_ref_new, weights=weights)         # :588
        xyz_src = _se3.transform(R_t.detach(), xyz_src)                                         # :590
        xyz_src = xyz_src.permute(0, 2, 1).contiguous()                                         # :591
        transforms.append(R_t if it == 0 else _se3.concatenate(R_t, transforms[-1]))            # :595
        all_matrices.append(logit)
        flags.append(flag)
        all_pred_pairs.append(indexs)

    # model.py:599-601: [B, J, 2] int32 on the CPU — the device->host copies happen here, once, after the loop
    B, J = all_pred_pairs[0].shape if all_pred_pairs else (0, 0)
    i0 = torch.arange(J)[None, :].expand(B, J).int()[:, :, None]
    endpoints['perm_matrices'] = all_matrices
    endpoints['pred_pairs'] = [torch.cat([i0, ix.int().cpu()[:, :, None]], dim=2) for ix in all_pred_pairs]
    endpoints['invalid_gradient'] = any(bool(f) for f in flags)
    endpoints['pt_ref_new'] = xyz_ref_new
    return transforms, endpoints


def _compute_rigid_transform_2(src, tgt, weights):
    """Leaf-level binding: the reference's loop does `invalid_gradient or cur_invalid_gradient` on a real bool
    (model.py:596); the lazy device flag is resolved here (one sync per iteration, where the reference has its own:
    the fp64 SVD on the host, model.py:47)."""
    T, flag = _K.compute_rigid_transform_2(src, tgt, weights)
    return T, bool(flag)


def _compute_rigid_transform(src, tgt, weights):
    T, flag = _K.compute_rigid_transform(src, tgt, weights)
    return T, bool(flag)


def patch(ref_root=None, level="loop"):
    """Rebind the hot-path names inside the reference's `network.model` (imported from sys.path or `ref_root`).
    Returns the module.  Idempotent."""
    if level not in ("loop", "leaf"):
        raise ValueError("level must be 'loop' or 'leaf'")
    mod = _import_reference(ref_root)
    if "model" not in _saved:
        _saved["model"] = {k: getattr(mod, k) for k in ("match_features_V2", "gather_neighbour_V3",
                                                         "compute_rigid_transform_2", "compute_rigid_transform", "se3_torch")}
        _saved["forward_align_4"] = mod.Network.forward_align_4
    mod.match_features_V2 = _M.match_features_V2
    mod.gather_neighbour_V3 = _M.gather_neighbour_V3
    mod.compute_rigid_transform_2 = _compute_rigid_transform_2
    mod.compute_rigid_transform = _compute_rigid_transform
    mod.se3_torch = _se3
    mod.Network.forward_align_4 = forward_align_4 if level == "loop" else _saved["forward_align_4"]
    return mod


def unpatch():
    """Undo patch() / patch_knn()."""
    if "model" in _saved:
        mod = importlib.import_module("network.model")
        for k, v in _saved.pop("model").items():
            setattr(mod, k, v)
        mod.Network.forward_align_4 = _saved.pop("forward_align_4")
    if "knn" in _saved:
        util, fn = _saved.pop("knn")
        util.knn = fn


class _KnnNamespace:
    """Stand-in for `import torch_points_kernels as Util` when that package is absent."""
    knn = staticmethod(_knn_fn)


def patch_knn(data_base_module):
    """dataloader/data_base.py:13 binds `torch_points_kernels` as `Util` and calls `Util.knn(support, query, k)` at
    :165,:170.  Rebind that one attribute; the points must already be on the device (CUDA contexts do not survive the
    DataLoader's fork: call nn_search after dict_all_to_device)."""
    util = getattr(data_base_module, "Util", None)
    if util is None:
        data_base_module.Util = _KnnNamespace
        return data_base_module
    if "knn" not in _saved:
        _saved["knn"] = (util, util.knn)
    util.knn = _knn_fn
    return data_base_module
