"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (torch-CPU / numpy, fp32 with the reference's op order; fp64 where the
reference uses fp64) of DeepSIR's correspondence-and-pose hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; ``deepsir_b200`` never does.

Pinning status (see DESIGN.md "Oracle"):
  * match / argmin / gather / affinity / sinkhorn / Kabsch_2 / SE(3): PINNED — checked
    bit-for-bit (integer) or to fp32 round-off against the reference's own functions
    imported from /root/reference in the build container; the outputs of the reference
    are committed as fixtures in tests/golden/ (generator: oracle/make_golden.py).
  * soft Kabsch ``compute_rigid_transform``: the reference function returns identity under
    torch>=2 (dtype bug at network/model.py:107).  The restatement keeps network/model.py:81-108
    with the centroids promoted to fp64 — PARITY UNPINNED for that one function.
  * KNN: PARITY UNPINNED — the arithmetic lives in torch_points_kernels (unvendored, unpinned
    version); semantics follow the call sites dataloader/data_base.py:165,170 and the tie rule
    defined in oracle/knn_oracle.c; cross-checked against scipy.spatial.cKDTree.

Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_EPS = 1e-16  # network/model.py:19, network/matchnet.py (same constant)
_HERE = os.path.dirname(os.path.abspath(__file__))
_KNN_SO = os.path.join(_HERE, "liboracle_knn.so")


# --------------------------------------------------------------------------------------
# feature distance (network/matchnet.py)
# --------------------------------------------------------------------------------------
def square_distance_V2(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """network/matchnet.py:96-113.  src [B,C,N], dst [B,C,M] -> [B,N,M] fp32.
    Op order kept: (-2 * (src^T dst)) + |src|^2, then + |dst|^2."""
    d = torch.matmul(src.permute(0, 2, 1).contiguous(), dst)
    d = -2 * d
    d += (src * src).sum(dim=1)[:, :, None]
    d += (dst * dst).sum(dim=1)[:, None, :]
    return d


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """network/matchnet.py:49-66.  src [B,N,C], dst [B,M,C] -> [B,N,M]."""
    d = -2 * torch.matmul(src, dst.permute(0, 2, 1).contiguous())
    d += (src * src).sum(dim=-1)[:, :, None]
    d += (dst * dst).sum(dim=-1)[:, None, :]
    return d


def match_features_V2(feat_src, feat_ref, metric="l2"):
    """network/matchnet.py:116-144.  channel-major features [B,C,J],[B,C,K] -> [B,J,K]."""
    assert feat_src.shape[1] == feat_ref.shape[1]
    if metric == "l2":
        return square_distance_V2(feat_src, feat_ref)
    if metric == "euclidean":
        return torch.sqrt(square_distance_V2(feat_src, feat_ref) + _EPS)
    if metric == "angle":
        a = feat_src / (torch.norm(feat_src, dim=1, keepdim=True) + _EPS)
        b = feat_ref / (torch.norm(feat_ref, dim=1, keepdim=True) + _EPS)
        return torch.acos(torch.matmul(a.permute(0, 2, 1).contiguous(), b))
    raise NotImplementedError(metric)


def match_features(feat_src, feat_ref, metric="l2"):
    """network/matchnet.py:69-93.  point-major features [B,J,C],[B,K,C] -> [B,J,K]."""
    assert feat_src.shape[-1] == feat_ref.shape[-1]
    if metric == "l2":
        return square_distance(feat_src, feat_ref)
    if metric == "angle":
        a = feat_src / (torch.norm(feat_src, dim=-1, keepdim=True) + _EPS)
        b = feat_ref / (torch.norm(feat_ref, dim=-1, keepdim=True) + _EPS)
        return torch.acos(torch.matmul(a, b.permute(0, 2, 1).contiguous()))
    raise NotImplementedError(metric)


def feat_dist(feat_src, feat_ref, metric="sqeuclidean"):
    """network/matchnet.py:147-192.  Broadcast-difference form, [B,C,J],[B,C,K] -> [B,J,K]."""
    diff = feat_src[:, :, :, None] - feat_ref[:, :, None, :]
    if metric == "sqeuclidean":
        return (diff * diff).sum(dim=1)
    if metric == "euclidean":
        return torch.sqrt((diff * diff).sum(dim=1) + _EPS)
    if metric == "cityblock":
        return diff.abs().sum(dim=1)
    if metric == "angle":
        return match_features_V2(feat_src, feat_ref, "angle")
    raise NotImplementedError(metric)


def match_argmin(feat_src, feat_ref, stride=6000):
    """network/model.py:558-569: source rows in chunks of `stride`, row-wise min index
    (torch.min returns the first minimal index).  [B,C,J],[B,C,K] -> int64 [B,J]."""
    J = feat_src.shape[2]
    out = []
    for lo in range(0, J, stride):
        m = match_features_V2(feat_src[:, :, lo:lo + stride], feat_ref)
        out.append(m.min(dim=2)[1])
    return torch.cat(out, dim=1)


def match_top2_fp64(feat_src, feat_ref, chunk=2048):
    """fp64 truth used to CLASSIFY rows (not part of the reference): returns
    (argmin int64 [B,J], gap = second-smallest - smallest distance, fp64 [B,J]).  A row whose gap is
    below fp32 round-off is 'tie-ambiguous': the reference's own sgemm order decides it."""
    fs, fr = feat_src.double(), feat_ref.double()
    nr = (fr * fr).sum(1)[:, None, :]
    idx, gap = [], []
    for lo in range(0, fs.shape[2], chunk):
        s = fs[:, :, lo:lo + chunk]
        d = -2 * torch.matmul(s.permute(0, 2, 1), fr) + (s * s).sum(1)[:, :, None] + nr
        if d.shape[2] < 2:
            idx.append(torch.zeros(d.shape[:2], dtype=torch.int64))
            gap.append(torch.full(d.shape[:2], float("inf"), dtype=torch.float64))
            continue
        v, i = torch.topk(d, 2, dim=2, largest=False)
        idx.append(i[:, :, 0])
        gap.append(v[:, :, 1] - v[:, :, 0])
    return torch.cat(idx, 1), torch.cat(gap, 1)


def gather_neighbour_V3(inputs, idx):
    """network/tools.py:211-221.  inputs [B,C,N], idx [B,M] -> [B,C,M]."""
    return torch.gather(inputs, 2, idx[:, None, :].expand(-1, inputs.shape[1], -1))


# --------------------------------------------------------------------------------------
# soft correspondence (network/matchnet.py)
# --------------------------------------------------------------------------------------
def gather_neighbour_V2(inputs, idx):
    """network/tools.py:197-209.  inputs [B,C,N], idx [B,M,k] -> [B,C,M,k]."""
    B, C, N = inputs.shape
    k = idx.shape[-1]
    flat = idx.reshape(B, -1)[:, None, :].expand(B, C, -1)
    return torch.gather(inputs, 2, flat).reshape(B, C, -1, k)


def relative_pos_encoding(xyz, idx):
    """network/RandLANet.py:197-212.  xyz [B,3,N], idx [B,N,k] -> [B,10,N,k] = {|rel|, rel, centre, neighbour}."""
    nb = gather_neighbour_V2(xyz, idx)
    tile = xyz[:, :, :, None].expand(-1, -1, -1, idx.shape[-1])
    rel = nb - tile
    dis = torch.sqrt(torch.sum(torch.pow(rel, 2), dim=1, keepdim=True))
    return torch.cat([dis, rel, tile, nb], dim=1)


def random_sample(feature, pool_idx):
    """network/RandLANet.py:374-391.  feature [B,C,N,1], pool_idx [B,M,k] -> [B,C,M,1]."""
    return gather_neighbour_V2(feature.squeeze(3), pool_idx).max(dim=3, keepdim=True)[0]


def nearest_interpolation(feature, interp_idx):
    """network/RandLANet.py:393-408.  feature [B,C,N,1], interp_idx [B,M,1] -> [B,C,M,1]."""
    return gather_neighbour_V3(feature.squeeze(3), interp_idx.reshape(interp_idx.shape[0], -1)).unsqueeze(3)


def compute_affinity(beta, feat_distance, alpha=0.5):
    """network/matchnet.py:195-208:  -beta_b * (d - alpha_b)."""
    if isinstance(alpha, float):
        return -beta[:, None, None] * (feat_distance - alpha)
    return -beta[:, None, None] * (feat_distance - alpha[:, None, None])


def soft_correspondence(feat_src, feat_ref, xyz_ref, beta, alpha):
    """Row-normalised affinity == one Sinkhorn row pass without slack (matchnet.py:259) ==
    SVDHead's softmax (matchnet.py:460-463), followed by the soft target of
    compute_rigid_transform (network/model.py:81-84).
    feat [B,C,J],[B,C,K]; xyz_ref [B,K,3] -> weights [B,J,K], y_soft [B,J,3], rowmass [B,J], lse [B,J]."""
    a = compute_affinity(beta, match_features_V2(feat_src, feat_ref), alpha)
    lse = torch.logsumexp(a, dim=2, keepdim=True)
    w = torch.exp(a - lse)
    s = w.sum(dim=2, keepdim=True)
    y = (w @ xyz_ref) / (s + _EPS)
    return w, y, s[:, :, 0], lse[:, :, 0]


def sinkhorn(log_alpha, n_iters=5, slack=True, eps=-1):
    """network/matchnet.py:211-271 including the eps early exit (:246-251, :262-267)."""
    prev = None
    if slack:
        la = torch.nn.functional.pad(log_alpha, (0, 1, 0, 1))
        for _ in range(n_iters):
            top = la[:, :-1, :] - torch.logsumexp(la[:, :-1, :], dim=2, keepdim=True)
            la = torch.cat((top, la[:, -1:, :]), dim=1)
            left = la[:, :, :-1] - torch.logsumexp(la[:, :, :-1], dim=1, keepdim=True)
            la = torch.cat((left, la[:, :, -1:]), dim=2)
            if eps > 0:
                cur = torch.exp(la[:, :-1, :-1])
                if prev is not None and torch.max(torch.sum(torch.abs(cur - prev), dim=[1, 2])) < eps:
                    break
                prev = cur.clone()
        return la[:, :-1, :-1]
    la = log_alpha
    for _ in range(n_iters):
        la = la - torch.logsumexp(la, dim=2, keepdim=True)
        la = la - torch.logsumexp(la, dim=1, keepdim=True)
        if eps > 0:
            cur = torch.exp(la)
            if prev is not None and torch.max(torch.sum(torch.abs(cur - prev), dim=[1, 2])) < eps:
                break
            prev = cur.clone()
    return la


def log_sinkhorn_iterations(Z, log_mu, log_nu, iters):
    """network/matchnet.py:827-833."""
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1)


def log_optimal_transport(scores, alpha, iters):
    """network/matchnet.py:836-856: dustbin row/column with score alpha, marginals (1 x m, n) / (1 x n, m) over m + n,
    result multiplied by m + n (Z - norm)."""
    b, m, n = scores.shape
    alpha = torch.as_tensor(alpha, dtype=scores.dtype)
    bins0 = alpha.expand(b, m, 1)
    bins1 = alpha.expand(b, 1, n)
    corner = alpha.expand(b, 1, 1)
    couplings = torch.cat([torch.cat([scores, bins0], -1), torch.cat([bins1, corner], -1)], 1)
    ms, ns = torch.tensor(float(m), dtype=scores.dtype), torch.tensor(float(n), dtype=scores.dtype)
    norm = -(ms + ns).log()
    log_mu = torch.cat([norm.expand(m), ns.log()[None] + norm])[None].expand(b, -1)
    log_nu = torch.cat([norm.expand(n), ms.log()[None] + norm])[None].expand(b, -1)
    return log_sinkhorn_iterations(couplings, log_mu, log_nu, iters) - norm


# --------------------------------------------------------------------------------------
# SE(3) (common/math/se3_torch.py)
# --------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------------
# key-point scoring / selection (network/model.py:668-757) and evaluation (loss.py:723-749, metrics_util.py:27-85)
# ---------------------------------------------------------------------------------------------------------
_EPS = 1e-16   # network/model.py:18


def score_fun(feat, xyz, prob, label, neigh_idx, label_weights, k_neighbors=16, ball_r=2.0):
    """network/model.py:700-757, statement by statement."""
    batch = feat.shape[0]
    neigh_idx = neigh_idx[:, :, :k_neighbors]
    max_per_sample = torch.max(feat.reshape(batch, -1), dim=1, keepdim=True)[0]
    feat_norm = feat / (max_per_sample.view(batch, 1, 1) + _EPS)
    neighbor_feat = torch.mean(gather_neighbour_V2(feat_norm, neigh_idx), dim=3)
    local_max_score = torch.nn.functional.softplus(feat_norm - neighbor_feat)
    neighbor_xyz = gather_neighbour_V2(xyz, neigh_idx)
    relative_xyz = torch.norm(neighbor_xyz - xyz.unsqueeze(-1), dim=1, keepdim=True)
    relative_xyz = torch.mean(relative_xyz, dim=-1)
    aggregation_score = (relative_xyz < ball_r).float()
    depth_wise_max = torch.max(feat_norm, dim=1, keepdim=True)[0]
    depth_wise_max_score = feat_norm / (depth_wise_max + _EPS)
    lw = torch.as_tensor(label_weights, dtype=torch.float32)
    label_score = lw[label.reshape(-1).long()].view(batch, 1, xyz.shape[-1])
    label_score = label_score / (torch.max(label_score, dim=-1, keepdim=True)[0] + _EPS)
    prob = prob / (torch.max(prob, dim=-1, keepdim=True)[0] + _EPS)
    label_score = label_score * torch.gt(prob, 0.2)
    score = local_max_score * aggregation_score * depth_wise_max_score * label_score
    return torch.max(score, dim=1)[0]


def topk_lower_index(score, k):
    """torch.topk(score, k, largest=True) (model.py:692) with the tie order fixed: equal values by ascending index."""
    B, N = score.shape
    vals, idxs = [], []
    for b in range(B):
        s = score[b].numpy()
        order = np.lexsort((np.arange(N), -s))     # primary: value descending, secondary: index ascending
        idxs.append(torch.from_numpy(order[:k].astype(np.int64)))
        vals.append(score[b][idxs[-1]])
    return torch.stack(vals), torch.stack(idxs)


def pair_hash(arr, M):
    """network/loss.py:280-294 for an [N,2] array."""
    arr = np.asarray(arr).astype(np.int64)
    return arr[:, 0] + arr[:, 1] * int(M)


def find_correct_correspondence(pos_pairs, pred_pairs, hash_seed=None, len_batch=None):
    """network/loss.py:723-749."""
    out = []
    for i in range(len(pos_pairs)):
        seed = max(len_batch[i]) if hash_seed is None else hash_seed
        out.append(np.isin(pair_hash(pred_pairs[i], seed), pair_hash(pos_pairs[i], seed), assume_unique=False))
    return np.stack(out, axis=0)


def rte_rre(T_pred, T_gt, eps=1e-16):
    """common/metrics_util.py:27-33 (numpy, in the dtype of the inputs like the reference: test.py:432-433 passes fp32)."""
    T_pred, T_gt = np.asarray(T_pred), np.asarray(T_gt)
    rte = np.linalg.norm(T_pred[:3, 3] - T_gt[:3, 3])
    rre = np.arccos(np.clip((np.trace(T_pred[:3, :3].T @ T_gt[:3, :3]) - 1) / 2, -1 + eps, 1 - eps)) * 180 / np.pi
    return rte, rre


def pose_residuals(pred, gt, eps=1e-16):
    """common/metrics_util.py:55-61."""
    c = se3_concatenate(se3_inverse(gt), pred)
    tr = c[:, 0, 0] + c[:, 1, 1] + c[:, 2, 2]
    deg = torch.acos(torch.clamp(0.5 * (tr - 1), min=-1 + eps, max=1 - eps)) * 180.0 / np.pi
    return deg, c[:, :, 3].norm(dim=-1)


def nn_sqdist(a, b):
    """common/metrics_util.py:38-40 + the row minimum of :72-73."""
    d = torch.sum((a[:, :, None, :] - b[:, None, :, :]) ** 2, dim=-1)
    return torch.min(d, dim=-1)[0]


def chamfer(points_src, points_ref, points_raw, pred, gt):
    """common/metrics_util.py:66-74."""
    src_transformed = se3_transform(pred, points_src)
    inter = se3_concatenate(pred, se3_inverse(gt))
    src_clean = se3_transform(inter, points_raw)
    return torch.mean(nn_sqdist(src_transformed, points_raw), dim=1) + torch.mean(nn_sqdist(points_ref, src_clean), dim=1)


def se3_identity(batch):
    """se3_torch.py:6-7."""
    return torch.eye(3, 4)[None].repeat(batch, 1, 1)


def se3_inverse(Rt):
    """se3_torch.py:10-25:  [R^T | -R^T t]."""
    R, t = Rt[..., :3, :3], Rt[..., :3, 3]
    Rt_ = R.transpose(-1, -2)
    return torch.cat([Rt_, Rt_ @ -t[..., None]], dim=-1)


def se3_concatenate(a, b):
    """se3_torch.py:28-48:  a o b = [Ra Rb | Ra tb + ta]."""
    Ra, ta, Rb, tb = a[..., :3, :3], a[..., :3, 3], b[..., :3, :3], b[..., :3, 3]
    return torch.cat([Ra @ Rb, Ra @ tb[..., None] + ta[..., None]], dim=-1)


def se3_transform(Rt, pts):
    """se3_torch.py:51-77:  pts [B,N,3] -> pts R^T + t."""
    return torch.matmul(pts, Rt[..., :3, :3].transpose(-1, -2)) + Rt[..., :3, 3][..., None, :]


def se3_transform_V2(Rt, pts):
    """se3_torch.py:80-100:  pts [B,3,N] -> R pts + t."""
    return torch.matmul(Rt[:, :3, :3], pts) + Rt[:, :3, 3][:, :, None]


# --------------------------------------------------------------------------------------
# weighted Kabsch (network/model.py)
# --------------------------------------------------------------------------------------
def _kabsch_from_cov(cov, c_src, c_tgt, keep_double):
    """network/model.py:45-58 (and :95-108): fp64 SVD on the host, R = V U^T, flip V[:, :, 2] when
    det <= 0, t = -R c_src + c_tgt."""
    u, s, v = torch.svd(cov.cpu().double(), some=False, compute_uv=True)
    r_pos = v @ u.transpose(-1, -2)
    v_neg = v.clone()
    v_neg[:, :, 2] *= -1
    r_neg = v_neg @ u.transpose(-1, -2)
    R = torch.where(torch.det(r_pos)[:, None, None] > 0, r_pos, r_neg)
    dev = c_src.device
    c_src, c_tgt = c_src.cpu(), c_tgt.cpu()        # :57 / :107: the solve stays on the host, the pose goes back to the device
    if not keep_double:
        R = R.float()
        t = -R @ c_src[:, :, None] + c_tgt[:, :, None]
    else:
        t = -R @ c_src.double()[:, :, None] + c_tgt.double()[:, :, None]
    return torch.cat((R, t), dim=2).float().to(dev), s


def compute_rigid_transform_2(src, tgt, weights, return_sv=False):
    """network/model.py:22-66.  src,tgt [B,M,3]; weights [B,M,1] -> (T [B,3,4] fp32, invalid)."""
    wn = weights / (weights.abs().sum(dim=1, keepdim=True) + _EPS)
    c_src = (src * wn).sum(dim=1)
    c_tgt = (tgt * wn).sum(dim=1)
    cov = (src - c_src[:, None, :]).transpose(-2, -1).contiguous() @ ((tgt - c_tgt[:, None, :]) * wn)
    T, s = _kabsch_from_cov(cov, c_src, c_tgt, keep_double=False)
    return (T, False, s) if return_sv else (T, False)


def compute_rigid_transform(src, tgt, weights):
    """network/model.py:68-116 with the torch>=2 dtype bug at :107 repaired by promoting the
    centroids to fp64 (the reference keeps R in fp64 until the final .float()).
    src [B,M,3], tgt [B,N,3], weights [B,M,N] -> (T [B,3,4], invalid)."""
    ws = weights.sum(dim=2, keepdim=True)
    wn = ws / (ws.sum(dim=1, keepdim=True) + _EPS)
    y = weights @ tgt / (ws + _EPS)
    c_src = (src * wn).sum(dim=1)
    c_tgt = (y * wn).sum(dim=1)
    cov = (src - c_src[:, None, :]).transpose(-2, -1).contiguous() @ ((y - c_tgt[:, None, :]) * wn)
    T, _ = _kabsch_from_cov(cov, c_src, c_tgt, keep_double=True)
    return T, False


def kabsch_moments_fp64(src, tgt, weights):
    """Additive raw moments in fp64: [Sabs, Sw, Sx(3), Sy(3), Sxy(9)] -> [B,17].
    Not in the reference; used to check the row-block sharded path: moments of blocks add up."""
    w = weights.double().reshape(weights.shape[0], -1)
    x, y = src.double(), tgt.double()
    out = torch.zeros(src.shape[0], 17, dtype=torch.float64)
    out[:, 0] = w.abs().sum(1)
    out[:, 1] = w.sum(1)
    out[:, 2:5] = (w[:, :, None] * x).sum(1)
    out[:, 5:8] = (w[:, :, None] * y).sum(1)
    out[:, 8:17] = torch.einsum("bm,bmi,bmj->bij", w, x, y).reshape(-1, 9)
    return out


def kabsch_from_moments_fp64(mom):
    """Kabsch from (summed) raw moments [B,17] fp64: the centred covariance of network/model.py:33-45 rewritten as
    H = Sxy/S - (2 - Sw/S) c_s c_t^T with S = S|w| + eps, c_s = Sx/S, c_t = Sy/S; then :47-58 (fp64 SVD, det fix)."""
    S = mom[:, 0] + _EPS
    c_s, c_t = mom[:, 2:5] / S[:, None], mom[:, 5:8] / S[:, None]
    H = mom[:, 8:17].reshape(-1, 3, 3) / S[:, None, None] - (2.0 - mom[:, 1] / S)[:, None, None] * c_s[:, :, None] * c_t[:, None, :]
    T, _ = _kabsch_from_cov(H, c_s, c_t, keep_double=True)
    return T, torch.zeros(mom.shape[0], dtype=torch.int32)


def rotation_angle_deg(Ra, Rb):
    """geodesic distance between rotations in degrees (same formula as network/loss.py:266 pose_error)."""
    R = Ra.double() @ Rb.double().transpose(-1, -2)
    tr = (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1.0) / 2.0
    # asin of the skew part is better conditioned than acos near 0
    sk = 0.5 * torch.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    ang = torch.atan2(sk.norm(dim=-1), tr)
    return torch.rad2deg(ang)


# --------------------------------------------------------------------------------------
# iterative re-match / re-solve loop (network/model.py:551-601), NN stages replaced by callables
# --------------------------------------------------------------------------------------
def align_loop(feat_src, feat_ref, xyz_src, xyz_ref, weights, num_iter, feature_fn=None, weight_fn=None):
    """network/model.py:551-601 without the two neural stages (they are not on the hot path):
    feat_* [B,C,N] are either fixed or re-derived by feature_fn(xyz_src [B,3,J]) each iteration
    (stands in for self.aggregation :552); weights [B,J,1] fixed or weight_fn(xyz_src, xyz_ref_new)
    (stands in for inlier_model + sigmoid :574-577).  xyz_* are [B,3,N] as in the reference loop.
    Returns (transforms list of [B,3,4] cumulative, pred_idx list of int64 [B,J], xyz_src_final [B,3,J])."""
    transforms, pred = [], []
    for it in range(num_iter):
        fs = feature_fn(xyz_src) if feature_fn is not None else feat_src
        idx = match_argmin(fs, feat_ref)                                  # :558-569
        ref_new = gather_neighbour_V3(xyz_ref, idx)                        # :571
        w = weight_fn(xyz_src, ref_new) if weight_fn is not None else weights
        src_p = xyz_src.permute(0, 2, 1).contiguous()                      # :586
        ref_p = ref_new.permute(0, 2, 1).contiguous()                      # :587
        T, _ = compute_rigid_transform_2(src_p, ref_p, w)                  # :588
        xyz_src = se3_transform(T, src_p).permute(0, 2, 1).contiguous()    # :590-591
        transforms.append(T if it == 0 else se3_concatenate(T, transforms[-1]))  # :595
        pred.append(idx)
    return transforms, pred, xyz_src


# --------------------------------------------------------------------------------------
# KNN (oracle/knn_oracle.c)
# --------------------------------------------------------------------------------------
def build_knn_lib(force=False):
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(_KNN_SO) or os.path.getmtime(_KNN_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               src, "-o", _KNN_SO, "-lm"])
    return _KNN_SO


_lib = None


def _knn_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_knn_lib())
        _lib.oracle_knn.restype = ctypes.c_int
        _lib.oracle_knn_pyramid.restype = ctypes.c_int
    return _lib


def knn(support, query, k):
    """Contract of torch_points_kernels.knn as used at dataloader/data_base.py:165,170:
    support [B,Ns,3], query [B,Nq,3] fp32 CPU -> (idx int64 [B,Nq,k], dist2 fp32 [B,Nq,k]) ascending."""
    s = np.ascontiguousarray(support.detach().cpu().numpy(), dtype=np.float32)
    q = np.ascontiguousarray(query.detach().cpu().numpy(), dtype=np.float32)
    B, Ns, _ = s.shape
    Nq = q.shape[1]
    idx = np.empty((B, Nq, k), dtype=np.int64)
    d2 = np.empty((B, Nq, k), dtype=np.float32)
    rc = _knn_lib().oracle_knn(s.ctypes.data_as(ctypes.c_void_p), q.ctypes.data_as(ctypes.c_void_p),
                               B, Ns, Nq, k, idx.ctypes.data_as(ctypes.c_void_p), d2.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("knn: fewer support points than k")
    return torch.from_numpy(idx), torch.from_numpy(d2)


def knn_numpy(support, query, k):
    """Independent numpy restatement of the same rule (small cases only; checks the C file)."""
    s = support.numpy().astype(np.float32)
    q = query.numpy().astype(np.float32)
    dx = q[:, :, None, 0] - s[:, None, :, 0]
    dy = q[:, :, None, 1] - s[:, None, :, 1]
    dz = q[:, :, None, 2] - s[:, None, :, 2]
    # fma emulated in fp64: products of fp32 are exact in fp64, one rounding per fma step
    d = (dx * dx).astype(np.float32)
    d = (dy.astype(np.float64) * dy.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
    d = (dz.astype(np.float64) * dz.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
    order = np.argsort(d, axis=2, kind="stable")[:, :, :k]
    return torch.from_numpy(order.astype(np.int64)), torch.from_numpy(np.take_along_axis(d, order, 2))


def nn_search(points, num_knn=16, ratios=(4, 4, 4, 4)):
    """dataloader/data_base.py:153-183 for one cloud tensor points [B,N,>=3] (CPU).
    Returns dict(xyz [B,sumN,3], neigh_idx [B,sumN,k], sub_idx [B,sumSub,k], interp_idx [B,sumN,1])."""
    pc = points[:, :, :3].contiguous()
    xs, nb, pool, up = [], [], [], []
    for r in ratios:
        idx, _ = knn(pc, pc, num_knn)
        m = pc.shape[1] // r
        sub = pc[:, :m, :].contiguous()
        u, _ = knn(sub, pc, 1)
        xs.append(pc); nb.append(idx); pool.append(idx[:, :m, :]); up.append(u)
        pc = sub
    return dict(xyz=torch.cat(xs, 1), neigh_idx=torch.cat(nb, 1), sub_idx=torch.cat(pool, 1), interp_idx=torch.cat(up, 1))


def nn_search_kdtree(points, num_knn=16, ratios=(4, 4, 4, 4), workers=-1):
    """DataBase.nn_search (dataloader/data_base.py:153-183) with scipy's cKDTree standing in for torch_points_kernels.knn
    (C++/nanoflann kd-tree, absent here): the SAME algorithm class as the reference's CPU path, used for the TIMED CPU
    baseline (the brute-force C oracle above defines the exact answer but is ~7x slower than a kd-tree at 16k points).
    Results agree with the exact oracle except at fp32 distance ties (tests/test_oracle_golden.py)."""
    from scipy.spatial import cKDTree
    pc_all = points[:, :, :3].contiguous().numpy()
    out = dict(xyz=[], neigh_idx=[], sub_idx=[], interp_idx=[])
    for b in range(pc_all.shape[0]):
        pc = pc_all[b]
        xs, nb, pool, up = [], [], [], []
        for r in ratios:
            _, idx = cKDTree(pc).query(pc, k=num_knn, workers=workers)
            m = pc.shape[0] // r
            sub = pc[:m]
            _, u = cKDTree(sub).query(pc, k=1, workers=workers)
            xs.append(pc); nb.append(idx.reshape(pc.shape[0], num_knn)); pool.append(idx.reshape(pc.shape[0], num_knn)[:m]); up.append(u.reshape(-1, 1))
            pc = sub
        out["xyz"].append(np.concatenate(xs)); out["neigh_idx"].append(np.concatenate(nb))
        out["sub_idx"].append(np.concatenate(pool)); out["interp_idx"].append(np.concatenate(up))
    return {k: torch.from_numpy(np.stack(v).astype(np.int64 if k != "xyz" else np.float32)) for k, v in out.items()}


def nn_search_c(points, num_knn=16, ratios=(4, 4, 4, 4)):
    """Same as nn_search but through the C pyramid driver (used as the timed cpu_baseline leg)."""
    p = np.ascontiguousarray(points.detach().cpu().numpy(), dtype=np.float32)
    B, N, S = p.shape
    L = len(ratios)
    sumN, sumSub, n = 0, 0, N
    for r in ratios:
        sumN += n; sumSub += n // r; n //= r
    xyz = np.empty((B, sumN, 3), np.float32)
    neigh = np.empty((B, sumN, num_knn), np.int64)
    sub = np.empty((B, sumSub, num_knn), np.int64)
    interp = np.empty((B, sumN, 1), np.int64)
    rat = (ctypes.c_int * L)(*ratios)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = _knn_lib().oracle_knn_pyramid(vp(p), B, N, S, rat, L, num_knn, vp(xyz), vp(neigh), vp(sub), vp(interp))
    if rc != 0:
        raise RuntimeError("knn pyramid: level smaller than k")
    return dict(xyz=torch.from_numpy(xyz), neigh_idx=torch.from_numpy(neigh), sub_idx=torch.from_numpy(sub),
                interp_idx=torch.from_numpy(interp))
