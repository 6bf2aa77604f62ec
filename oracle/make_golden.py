"""Generate tests/golden/*.npz by running the REFERENCE's own functions (imported from
/root/reference, build container only) on seeded synthetic inputs.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # rewrites tests/golden/

The fixtures pin the oracle (oracle/deepsir_oracle.py) — tests/test_oracle_golden.py replays them
without the reference being present (the GPU box has no /root/reference).
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DEEPSIR_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from network import matchnet as R_match  # noqa: E402  (reference)
from network import model as R_model  # noqa: E402
from network import tools as R_tools  # noqa: E402
from common.math import se3_torch as R_se3  # noqa: E402

from deepsir_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)  # single-thread MKL: the fixtures do not depend on the host's core count


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()[:16]


def save(name, **kw):
    arrs = {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in kw.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(name, {k: v.shape for k, v in arrs.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(7)

    # ---- feature distance (matchnet.py:49-192) -------------------------------------------------
    fs = torch.nn.functional.normalize(torch.randn(2, 64, 150, generator=g), dim=1)
    fr = torch.nn.functional.normalize(torch.randn(2, 64, 170, generator=g), dim=1)
    save("match_dense",
         feat_src=fs, feat_ref=fr,
         l2=R_match.match_features_V2(fs, fr, "l2"),
         euclidean=R_match.match_features_V2(fs, fr, "euclidean"),
         angle=R_match.match_features_V2(fs, fr, "angle"),
         nc_l2=R_match.match_features(fs.permute(0, 2, 1).contiguous(), fr.permute(0, 2, 1).contiguous(), "l2"),
         fd_sq=R_match.feat_dist(fs, fr, "sqeuclidean"),
         fd_city=R_match.feat_dist(fs, fr, "cityblock"),
         fd_euc=R_match.feat_dist(fs, fr, "euclidean"))

    # ---- chunked argmin (model.py:558-569) on planted features, N=1500 with stride forced to 600 and
    #      at the shipped stride 6000 on N=7000 (two chunks) -----------------------------------------
    def ref_argmin(fs_, fr_, stride):
        out = []
        for n in range(int(np.ceil(fs_.shape[2] / stride))):
            mm = R_match.match_features_V2(fs_[:, :, n * stride:(n + 1) * stride], fr_)
            out.append(mm.min(dim=2, keepdim=False)[1])
        return torch.cat(out, dim=1)

    b = synth.make_batch(2, 1500, 64, "kitti", config=1)
    save("match_argmin_1500", seed_config=1, n=1500, d=64, batch=2, sha_src=sha(b["feat_src"]), sha_ref=sha(b["feat_ref"]),
         idx=ref_argmin(b["feat_src"], b["feat_ref"], 600), idx_full=ref_argmin(b["feat_src"], b["feat_ref"], 6000))
    b = synth.make_batch(1, 7000, 32, "3dmatch", config=3)
    save("match_argmin_7000", seed_config=3, n=7000, d=32, batch=1, sha_src=sha(b["feat_src"]), sha_ref=sha(b["feat_ref"]),
         idx=ref_argmin(b["feat_src"], b["feat_ref"], 6000))
    # ties: duplicated reference columns -> torch.min returns the first index
    fr_dup = torch.cat([fr[:, :, :40], fr[:, :, :40], fr[:, :, 40:]], dim=2)
    save("match_argmin_ties", feat_src=fs, feat_ref=fr_dup, idx=ref_argmin(fs, fr_dup, 6000))

    # ---- gather (tools.py:211-221) -----------------------------------------------------------------
    xyz = torch.randn(2, 3, 170, generator=g)
    gi = torch.randint(0, 170, (2, 150), generator=g)
    save("gather_v3", inputs=xyz, idx=gi, out=R_tools.gather_neighbour_V3(xyz, gi))

    # ---- affinity + sinkhorn (matchnet.py:195-271) ----------------------------------------------
    d = R_match.match_features_V2(fs[:, :, :40], fr[:, :, :50])
    beta = torch.tensor([10.0, 4.0])
    alpha = torch.tensor([0.5, 0.3])
    aff = R_match.compute_affinity(beta, d, alpha)
    save("affinity_sinkhorn", dist=d, beta=beta, alpha=alpha, affinity=aff,
         affinity_scalar_alpha=R_match.compute_affinity(beta, d),
         sinkhorn_slack=R_match.sinkhorn(aff, n_iters=5, slack=True),
         sinkhorn_noslack=R_match.sinkhorn(aff, n_iters=5, slack=False),
         sinkhorn_1row=aff - torch.logsumexp(aff, dim=2, keepdim=True))

    # ---- weighted Kabsch (model.py:22-66) ---------------------------------------------------------
    cases = {}
    b = synth.make_batch(3, 400, 64, "kitti", config=1, first_pair=10)
    src = b["points_src"][:, :, :3].contiguous()
    tgt_full = torch.stack([b["points_ref"][i, b["perm"][i], :3] for i in range(3)])
    cases["planted"] = (src, tgt_full, b["weights"])
    cases["uniform_w"] = (src, tgt_full, torch.ones(3, 400, 1))
    planar = src.clone(); planar[:, :, 2] = 0.0
    Rz = synth._rot_zyx(0.3, 0.0, 0.0).float()
    cases["planar"] = (planar, planar @ Rz.t() + torch.tensor([1.0, -2.0, 0.5]), torch.ones(3, 400, 1))
    refl = src[:, :50].clone()
    cases["reflection"] = (refl, refl * torch.tensor([1.0, 1.0, -1.0]), torch.ones(3, 50, 1))
    cases["neg_w"] = (src, tgt_full, b["weights"] - 0.2)
    cases["single_heavy"] = (src, tgt_full, torch.cat([torch.full((3, 1, 1), 1e3), torch.ones(3, 399, 1)], 1))
    kw = {}
    for name, (s, t, w) in cases.items():
        T, inv = R_model.compute_rigid_transform_2(s, t, w)
        assert inv is False
        kw[name + "_src"], kw[name + "_tgt"], kw[name + "_w"], kw[name + "_T"] = s, t, w, T
    save("kabsch2", **kw)

    # ---- SE(3) (se3_torch.py) ------------------------------------------------------------------
    Ta = torch.stack([synth.random_pose(g) for _ in range(4)])
    Tb = torch.stack([synth.random_pose(g, any_axis=True, yaw_deg=90, tilt_scale=1.0) for _ in range(4)])
    pts = torch.randn(4, 33, 3, generator=g) * 20
    save("se3", Ta=Ta, Tb=Tb, pts=pts, identity=R_se3.identity(4), inverse=R_se3.inverse(Ta),
         concat=R_se3.concatenate(Ta, Tb), transform=R_se3.transform(Ta, pts),
         transform_v2=R_se3.transform_V2(Ta, pts.permute(0, 2, 1).contiguous()))

    # ---- one full loop of model.py:551-601 with fixed features/weights (NN stages removed) -------------
    b = synth.make_batch(2, 1200, 64, "oxford", config=5)
    xyz_src = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xyz_ref = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    transforms, preds = [], []
    for it in range(3):
        idx = ref_argmin(b["feat_src"], b["feat_ref"], 6000)
        ref_new = R_tools.gather_neighbour_V3(xyz_ref, idx)
        sp = xyz_src.permute(0, 2, 1).contiguous()
        T, _ = R_model.compute_rigid_transform_2(sp, ref_new.permute(0, 2, 1).contiguous(), b["weights"])
        xyz_src = R_se3.transform(T, sp).permute(0, 2, 1).contiguous()
        transforms.append(T if it == 0 else R_se3.concatenate(T, transforms[-1]))
        preds.append(idx)
    save("loop_oxford_1200", seed_config=5, n=1200, d=64, batch=2, sha_src=sha(b["feat_src"]),
         transforms=torch.stack(transforms), pred=torch.stack(preds), xyz_src_final=xyz_src,
         transform_gt=b["transform_gt"])


def graph():
    """KNN consumers (network/tools.py, network/RandLANet.py) and larger Sinkhorn cases from the reference itself."""
    from network import RandLANet as R_rla  # noqa: E402  (reference)
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(11)
    B, C, N, k = 2, 5, 300, 16
    xyz = torch.randn(B, 3, N, generator=g) * 10
    feat = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, N, k), generator=g)
    pool = idx[:, :N // 4, :].contiguous()
    interp = torch.randint(0, N // 4, (B, N, 1), generator=g)
    sub_feat = torch.randn(B, C, N // 4, generator=g)
    la = torch.randn(2, 37, 53, generator=g) * 3
    save("graph_ops", xyz=xyz, feat=feat, idx=idx, pool=pool, interp=interp, sub_feat=sub_feat,
         gather_v2=R_tools.gather_neighbour_V2(feat, idx),
         gather_v1=R_tools.gather_neighbour(feat.permute(0, 2, 1).contiguous(), idx),
         gather_v4=R_tools.gather_neighbour_V4(feat.permute(0, 2, 1).contiguous(), idx[:, :, 0].contiguous()),
         rel_pos=R_rla.Building_block.relative_pos_encoding(xyz, idx),
         pooled=R_rla.RandLA.random_sample(feat[:, :, :, None], pool),
         interpolated=R_rla.RandLA.nearest_interpolation(sub_feat[:, :, :, None], interp),
         log_alpha=la,
         sinkhorn_slack_5=R_match.sinkhorn(la, n_iters=5, slack=True),
         sinkhorn_noslack_3=R_match.sinkhorn(la, n_iters=3, slack=False),
         sinkhorn_slack_eps=R_match.sinkhorn(la, n_iters=50, slack=True, eps=1e-2))


def keypoint_eval():
    """score_fun / feat_score (model.py:668-757), find_correct_correspondence (loss.py:723-749), compute_metrics and rte_rre
    (metrics_util.py:27-85) run from the reference on seeded inputs."""
    import types
    from network import loss as R_loss
    from common import metrics_util as R_met
    from oracle import deepsir_oracle as O
    from scipy.spatial.transform import Rotation
    if not hasattr(Rotation, "from_dcm"):      # removed from scipy >= 1.6; only the Euler-angle r_mse/r_mae (not on the
        Rotation.from_dcm = Rotation.from_matrix   # path) of compute_metrics go through it (common/math/so3.py:23)
    g = torch.Generator().manual_seed(23)
    B, C, N, k = 2, 32, 1024, 16
    b = synth.make_batch(B, N, C, "kitti", config=1, first_pair=40)
    xyz = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    feat = torch.rand(B, C, N, generator=g) * 3 - 0.5
    prob = torch.rand(B, 1, N, generator=g)
    label = torch.randint(0, 19, (B, 1, N), generator=g)
    neigh = O.nn_search_c(b["points_src"], k, (4, 4, 4, 4))["neigh_idx"][:, :N]
    lw = torch.tensor(R_model_label_weights()).float()
    fake = types.SimpleNamespace(num_knn=16, label_weights=lw)
    score = R_model.Network.score_fun(fake, feat, xyz, prob, label, neigh)
    fake2 = types.SimpleNamespace(num_knn=16, label_weights=lw, sub_selection=True,
                                  score_fun=lambda *a: R_model.Network.score_fun(fake, *a))
    f_sub, x_sub, l_sub, s_sub = R_model.Network.feat_score(fake2, feat, xyz, prob, label, neigh, num_sub=200)
    # correspondence check
    pos = [torch.stack([torch.arange(N), torch.randperm(N, generator=g)], 1)[: N - 100 * i].int() for i in range(B)]
    pred = torch.stack([torch.stack([torch.arange(N), torch.where(torch.rand(N, generator=g) < 0.6, pos[i][:, 1].long()[torch.arange(N) % pos[i].shape[0]],
                                                                   torch.randint(0, N, (N,), generator=g))], 1) for i in range(B)]).int()
    corr_seed = R_loss.ScanAlignmentLoss.find_correct_correspondence(None, [p.numpy() for p in pos], pred.numpy(), hash_seed=N)
    # pose metrics
    gt = b["transform_gt"][:, :3, :]
    gn = torch.Generator().manual_seed(5)
    noise = torch.stack([synth.random_pose(gn, 3.0 * (i + 1), 1.0, 0.3 * (i + 1), any_axis=True) for i in range(B)])
    pred_T = R_se3.concatenate(noise, gt)
    data = dict(transform_gt=gt, points_src=b["points_src"], points_ref=b["points_ref"])
    m = R_met.compute_metrics(data, pred_T, 2.0, 5.0)
    rr = np.stack([R_met.rte_rre(pred_T[i].numpy(), gt[i].numpy(), 2.0, 5.0) for i in range(B)])
    save("keypoint_eval", feat=feat, xyz=xyz, prob=prob, label=label, neigh=neigh, label_weights=lw, score=score,
         sub_feat=f_sub, sub_xyz=x_sub, sub_label=l_sub, sub_score=s_sub,
         pos0=pos[0], pos1=pos[1], pred_pairs=pred, correct=corr_seed,
         transform_gt=gt, transform_pred=pred_T, points_src=b["points_src"], points_ref=b["points_ref"],
         err_r_deg=m["err_r_deg"], err_t=m["err_t"], succ=m["succ"], chamfer_dist=m["chamfer_dist"], rte_rre=rr)


def log_ot():
    """log_optimal_transport / log_sinkhorn_iterations (network/matchnet.py:827-856) from the reference itself."""
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(13)
    sc = torch.randn(2, 41, 57, generator=g) * 2.5
    fs = torch.nn.functional.normalize(torch.randn(2, 32, 120, generator=g), dim=1)
    fr = torch.nn.functional.normalize(torch.randn(2, 32, 140, generator=g), dim=1)
    beta, alpha = torch.tensor([10.0, 6.0]), torch.tensor([0.5, 0.4])
    aff = R_match.compute_affinity(beta, R_match.match_features_V2(fs, fr), alpha)
    save("log_ot", scores=sc, bin1=torch.tensor(1.0), bin2=torch.tensor(-0.7),
         out_bin1_it20=R_match.log_optimal_transport(sc, torch.tensor(1.0), 20),
         out_bin2_it3=R_match.log_optimal_transport(sc, torch.tensor(-0.7), 3),
         out_it0=R_match.log_optimal_transport(sc, torch.tensor(0.25), 0),
         feat_src=fs, feat_ref=fr, beta=beta, alpha=alpha,
         out_affinity_it5=R_match.log_optimal_transport(aff, torch.tensor(0.3), 5))


def R_model_label_weights():
    return [3, 1, 1, 3, 2, 0, 0, 0, 6, 5, 6, 4, 7, 7, 6, 8, 4, 9, 9]   # network/model.py:146-149


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "graph":
        graph()
    elif len(sys.argv) > 1 and sys.argv[1] == "keypoint":
        keypoint_eval()
    elif len(sys.argv) > 1 and sys.argv[1] == "log_ot":
        log_ot()
    else:
        main()
        graph()
        keypoint_eval()
        log_ot()
