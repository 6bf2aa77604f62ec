"""The oracle restatement replayed against fixtures produced by the REFERENCE's own functions
(oracle/make_golden.py, run in the build container).  CPU only."""
import hashlib

import numpy as np
import torch

from deepsir_b200 import synth
from oracle import deepsir_oracle as O

torch.set_num_threads(1)


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()[:16]


def test_match_dense_matches_reference(golden):
    g = golden("match_dense")
    fs, fr = g["feat_src"], g["feat_ref"]
    # same torch ops in the same order -> bitwise equal on the same BLAS; tolerance covers other hosts
    assert torch.allclose(O.match_features_V2(fs, fr, "l2"), g["l2"], atol=2e-6, rtol=0)
    assert torch.allclose(O.match_features_V2(fs, fr, "euclidean"), g["euclidean"], atol=2e-6, rtol=0)
    assert torch.allclose(O.match_features_V2(fs, fr, "angle"), g["angle"], atol=2e-6, rtol=0)
    assert torch.allclose(O.match_features(fs.permute(0, 2, 1).contiguous(), fr.permute(0, 2, 1).contiguous()),
                          g["nc_l2"], atol=2e-6, rtol=0)
    assert torch.allclose(O.feat_dist(fs, fr, "sqeuclidean"), g["fd_sq"], atol=2e-6, rtol=0)
    assert torch.allclose(O.feat_dist(fs, fr, "cityblock"), g["fd_city"], atol=1e-5, rtol=0)
    assert torch.allclose(O.feat_dist(fs, fr, "euclidean"), g["fd_euc"], atol=2e-6, rtol=0)


def test_match_argmin_matches_reference(golden):
    g = golden("match_argmin_1500")
    b = synth.make_batch(int(g["batch"]), int(g["n"]), int(g["d"]), "kitti", config=int(g["seed_config"]))
    assert _sha(b["feat_src"]) == str(g["sha_src"]) and _sha(b["feat_ref"]) == str(g["sha_ref"])
    assert torch.equal(O.match_argmin(b["feat_src"], b["feat_ref"], 600), g["idx"])
    assert torch.equal(O.match_argmin(b["feat_src"], b["feat_ref"], 6000), g["idx_full"])
    # planted matches: inlier rows recover the planted permutation
    inl = b["inlier"]
    assert torch.equal(g["idx"][inl], b["perm"][inl])
    g = golden("match_argmin_7000")
    b = synth.make_batch(1, 7000, 32, "3dmatch", config=3)
    assert _sha(b["feat_src"]) == str(g["sha_src"])
    assert torch.equal(O.match_argmin(b["feat_src"], b["feat_ref"]), g["idx"])


def test_match_argmin_ties_first_index(golden):
    g = golden("match_argmin_ties")
    idx = O.match_argmin(g["feat_src"], g["feat_ref"])
    assert torch.equal(idx, g["idx"])
    i64, gap = O.match_top2_fp64(g["feat_src"], g["feat_ref"])
    dup = (i64 < 80)
    assert dup.any() and (gap[dup] < 1e-12).all()


def test_gather_affinity_sinkhorn(golden):
    g = golden("gather_v3")
    assert torch.equal(O.gather_neighbour_V3(g["inputs"], g["idx"]), g["out"])
    g = golden("affinity_sinkhorn")
    aff = O.compute_affinity(g["beta"], g["dist"], g["alpha"])
    assert torch.equal(aff, g["affinity"])
    assert torch.equal(O.compute_affinity(g["beta"], g["dist"]), g["affinity_scalar_alpha"])
    assert torch.allclose(O.sinkhorn(aff, 5, True), g["sinkhorn_slack"], atol=1e-6, rtol=0)
    assert torch.allclose(O.sinkhorn(aff, 5, False), g["sinkhorn_noslack"], atol=1e-6, rtol=0)


def test_kabsch2_matches_reference(golden):
    g = golden("kabsch2")
    for name in ["planted", "uniform_w", "planar", "reflection", "neg_w", "single_heavy"]:
        T, inv = O.compute_rigid_transform_2(g[name + "_src"], g[name + "_tgt"], g[name + "_w"])
        assert inv is False
        assert torch.allclose(T, g[name + "_T"], atol=1e-6, rtol=0), name
        assert (torch.det(T[:, :, :3]) > 0).all()


def test_se3_matches_reference(golden):
    g = golden("se3")
    assert torch.equal(O.se3_identity(4), g["identity"])
    assert torch.allclose(O.se3_inverse(g["Ta"]), g["inverse"], atol=1e-7, rtol=0)
    assert torch.allclose(O.se3_concatenate(g["Ta"], g["Tb"]), g["concat"], atol=1e-7, rtol=0)
    assert torch.allclose(O.se3_transform(g["Ta"], g["pts"]), g["transform"], atol=1e-6, rtol=0)
    assert torch.allclose(O.se3_transform_V2(g["Ta"], g["pts"].permute(0, 2, 1).contiguous()), g["transform_v2"],
                          atol=1e-6, rtol=0)


def test_loop_matches_reference(golden):
    g = golden("loop_oxford_1200")
    b = synth.make_batch(2, 1200, 64, "oxford", config=5)
    assert _sha(b["feat_src"]) == str(g["sha_src"])
    xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    tr, pred, xyz = O.align_loop(b["feat_src"], b["feat_ref"], xs, xr, b["weights"], 3)
    assert torch.equal(torch.stack(pred), g["pred"])
    assert torch.allclose(torch.stack(tr), g["transforms"], atol=1e-6, rtol=0)
    assert torch.allclose(xyz, g["xyz_src_final"], atol=1e-5, rtol=0)
    # the planted pose is recovered up to the pull of the 10 % down-weighted outliers
    assert O.rotation_angle_deg(tr[0][:, :, :3], g["transform_gt"][:, :, :3]).max() < 0.5


def test_soft_kabsch_restatement_equals_hard_on_onehot():
    """compute_rigid_transform (soft) with one-hot weights == compute_rigid_transform_2 with unit weights."""
    b = synth.make_batch(2, 200, 32, "3dmatch", config=3, first_pair=5)
    src = b["points_src"][:, :, :3].contiguous()
    ref = b["points_ref"][:, :, :3].contiguous()
    W = torch.zeros(2, 200, 200)
    W[torch.arange(2)[:, None], torch.arange(200)[None], b["perm"]] = 1.0
    T_soft, _ = O.compute_rigid_transform(src, ref, W)
    tgt = torch.stack([ref[i, b["perm"][i]] for i in range(2)])
    T_hard, _ = O.compute_rigid_transform_2(src, tgt, torch.ones(2, 200, 1))
    assert torch.allclose(T_soft, T_hard, atol=1e-5)


def test_knn_oracle_c_vs_numpy_and_kdtree():
    from scipy.spatial import cKDTree
    g = torch.Generator().manual_seed(3)
    pts = synth.kitti_cloud(700, g)[:, :3][None].contiguous()
    qry = synth.kitti_cloud(300, g)[:, :3][None].contiguous()
    i_c, d_c = O.knn(pts, qry, 16)
    i_n, d_n = O.knn_numpy(pts, qry, 16)
    assert torch.equal(i_c, i_n) and torch.equal(d_c, d_n)
    assert (d_c[:, :, 1:] >= d_c[:, :, :-1]).all()
    _, i_k = cKDTree(pts[0].numpy().astype(np.float64)).query(qry[0].numpy().astype(np.float64), k=16)
    assert (torch.from_numpy(i_k) == i_c[0]).float().mean() > 0.999
    # self query: nearest neighbour is the point itself at distance 0
    i_s, d_s = O.knn(pts, pts, 4)
    assert torch.equal(i_s[0, :, 0], torch.arange(700)) and (d_s[:, :, 0] == 0).all()


def test_knn_oracle_ties_and_errors():
    # integer lattice: many exact ties -> lower index first
    ax = torch.arange(6, dtype=torch.float32)
    lat = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3).contiguous()
    i_c, d_c = O.knn(lat, lat, 7)
    i_n, d_n = O.knn_numpy(lat, lat, 7)
    assert torch.equal(i_c, i_n) and torch.equal(d_c, d_n)
    ties = d_c[:, :, 1:] == d_c[:, :, :-1]
    assert ties.any() and (i_c[:, :, 1:][ties] > i_c[:, :, :-1][ties]).all()
    # duplicated points (FixedResampler tiling)
    dup = torch.cat([lat[:, :50], lat[:, :50]], 1).contiguous()
    i_d, d_d = O.knn(dup, dup, 2)
    assert torch.equal(i_d[0, 50:, 0], torch.arange(50)) and (d_d[:, :, 1] == 0).all()
    import pytest
    with pytest.raises(RuntimeError):
        O.knn(lat[:, :5], lat, 16)


def test_nn_search_pyramid_shapes_and_c_driver():
    g = torch.Generator().manual_seed(5)
    pts = torch.stack([synth.kitti_cloud(1024, g) for _ in range(2)])
    a = O.nn_search(pts, 16, (4, 4, 4))
    c = O.nn_search_c(pts, 16, (4, 4, 4))
    assert a["xyz"].shape == (2, 1024 + 256 + 64, 3) and a["sub_idx"].shape == (2, 256 + 64 + 16, 16)
    for k in a:
        assert torch.equal(a[k], c[k]), k
    assert a["interp_idx"].max() < 256 and a["neigh_idx"][:, 1024:1280].max() < 256


def test_kdtree_baseline_agrees_with_exact_oracle():
    """The kd-tree stand-in that bench.py TIMES as the CPU baseline computes the same pyramid as the exact oracle
    (identical except at fp32 distance ties)."""
    from deepsir_b200 import synth
    b = synth.make_batch(1, 2048, 8, "kitti", config=1, first_pair=77)
    a = O.nn_search_c(b["points_src"], 16, (4, 4, 4, 4))
    k = O.nn_search_kdtree(b["points_src"], 16, (4, 4, 4, 4), workers=1)
    assert torch.equal(a["xyz"], k["xyz"])
    for key in ("neigh_idx", "sub_idx", "interp_idx"):
        assert a[key].shape == k[key].shape
        assert (a[key] == k[key]).float().mean() > 0.999, key


def test_graph_ops_oracle_matches_reference_fixture(golden):
    """KNN consumers (tools.py / RandLANet.py) and Sinkhorn incl. the eps early exit: oracle == reference output."""
    g = golden("graph_ops")
    assert torch.equal(O.gather_neighbour_V2(g["feat"], g["idx"]), g["gather_v2"])
    assert torch.equal(O.gather_neighbour_V2(g["feat"], g["idx"]).permute(0, 2, 3, 1), g["gather_v1"])
    assert torch.equal(O.relative_pos_encoding(g["xyz"], g["idx"]), g["rel_pos"])
    assert torch.equal(O.random_sample(g["feat"][:, :, :, None], g["pool"]), g["pooled"])
    assert torch.equal(O.nearest_interpolation(g["sub_feat"][:, :, :, None], g["interp"]), g["interpolated"])
    assert torch.allclose(O.sinkhorn(g["log_alpha"], 5, True), g["sinkhorn_slack_5"], atol=2e-6, rtol=0)
    assert torch.allclose(O.sinkhorn(g["log_alpha"], 3, False), g["sinkhorn_noslack_3"], atol=2e-6, rtol=0)
    assert torch.allclose(O.sinkhorn(g["log_alpha"], 50, True, eps=1e-2), g["sinkhorn_slack_eps"], atol=2e-6, rtol=0)


def test_keypoint_and_eval_oracle_matches_reference_fixture(golden):
    """score_fun / feat_score (model.py:668-757), find_correct_correspondence (loss.py:723-749), compute_metrics and
    rte_rre (metrics_util.py:27-85): oracle == the reference's own outputs."""
    g = golden("keypoint_eval")
    s = O.score_fun(g["feat"], g["xyz"], g["prob"], g["label"], g["neigh"], g["label_weights"])
    assert torch.allclose(s, g["score"], rtol=1e-6, atol=1e-7)
    v, i = O.topk_lower_index(g["score"], 200)
    assert torch.equal(v, g["sub_score"])                       # torch.topk values are order-independent
    tie_free = (v[:, 1:] != v[:, :-1]).all(dim=1)               # indices too wherever the values are distinct
    for b in range(v.shape[0]):
        if tie_free[b]:
            assert torch.equal(O.gather_neighbour_V3(g["xyz"], i)[b], g["sub_xyz"][b])
    c = O.find_correct_correspondence([g["pos0"].numpy(), g["pos1"].numpy()], g["pred_pairs"].numpy(), hash_seed=1024)
    assert np.array_equal(c, g["correct"].numpy().astype(bool))
    deg, mag = O.pose_residuals(g["transform_pred"], g["transform_gt"])
    assert torch.allclose(deg, g["err_r_deg"], atol=1e-5) and torch.allclose(mag, g["err_t"], atol=1e-6)
    for b in range(2):
        rte, rre = O.rte_rre(g["transform_pred"][b].numpy(), g["transform_gt"][b].numpy())
        assert abs(rte - g["rte_rre"][b, 1].item()) < 1e-7 and abs(rre - g["rte_rre"][b, 2].item()) < 1e-5
    src, ref = g["points_src"][:, :2048, :3], g["points_ref"][:, :2048, :3]
    raw = torch.cat([O.se3_transform(g["transform_gt"], src), ref], dim=1)
    ch = O.chamfer(src, ref, raw, g["transform_pred"], g["transform_gt"])
    assert torch.allclose(ch, g["chamfer_dist"], rtol=1e-6, atol=1e-8)


def test_log_optimal_transport_oracle_matches_reference_fixture(golden):
    """network/matchnet.py:827-856 restated in the oracle == the reference's own outputs (oracle/make_golden.py log_ot)."""
    z = golden("log_ot")
    for key, alpha, it in (("out_bin1_it20", 1.0, 20), ("out_bin2_it3", -0.7, 3), ("out_it0", 0.25, 0)):
        out = O.log_optimal_transport(z["scores"], alpha, it)
        assert out.shape == z[key].shape == (2, 42, 58)
        assert torch.allclose(out, z[key], rtol=0, atol=2e-6), key
    aff = O.compute_affinity(z["beta"], O.match_features_V2(z["feat_src"], z["feat_ref"]), z["alpha"])
    assert torch.allclose(O.log_optimal_transport(aff, 0.3, 5), z["out_affinity_it5"], atol=5e-6)
