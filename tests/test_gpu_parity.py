"""Parity of the sm_100a path (called through the C-ABI library via the host mirror) against the CPU oracle
and the golden fixtures produced by the reference.  Bars (BASELINE.json north_star):
  * KNN index sets and hard correspondences: bit-exact (ties to the lower index)
  * soft weights / targets: 1e-4 relative;  R: 1e-3 deg;  t: 1e-4 m
"""
import numpy as np
import pytest
import torch

import deepsir_b200 as D
from deepsir_b200 import synth
from oracle import deepsir_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROT_TOL_DEG, TRANS_TOL_M, SOFT_RTOL = 1e-3, 1e-4, 1e-4


def cu(t):
    return t.to(DEV)


def assert_pose_close(T_gpu, T_ref, scale=1.0):
    T_gpu = T_gpu.cpu()
    ang = O.rotation_angle_deg(T_gpu[:, :, :3], T_ref[:, :, :3]).max().item()
    dt = (T_gpu[:, :, 3] - T_ref[:, :, 3]).norm(dim=1).max().item()
    assert ang <= ROT_TOL_DEG, f"rotation differs by {ang} deg"
    assert dt <= TRANS_TOL_M * scale, f"translation differs by {dt} m"


# ------------------------------------------------------------------------------------------- KNN
@pytest.mark.parametrize("algo", [D.KNN_BRUTE, D.KNN_GRID, D.KNN_TREE, D.KNN_AUTO])
def test_knn_random_cloud_bit_exact(algo):
    g = torch.Generator().manual_seed(3)
    sup = torch.stack([synth.kitti_cloud(2500, g) for _ in range(2)])
    qry = torch.stack([synth.kitti_cloud(777, g) for _ in range(2)])
    for k in (1, 3, 16, 20):
        i_o, d_o = O.knn(sup[:, :, :3].contiguous(), qry[:, :, :3].contiguous(), k)
        i_g, d_g = D.knn(cu(sup), cu(qry), k, algo=algo)   # stride-4 inputs, xyz are the first 3 columns
        assert i_g.dtype == torch.int64 and i_g.shape == (2, 777, k)
        assert torch.equal(i_g.cpu(), i_o), k
        assert torch.equal(d_g.cpu(), d_o), k


@pytest.mark.parametrize("algo", [D.KNN_BRUTE, D.KNN_GRID, D.KNN_TREE, D.KNN_AUTO])
def test_knn_ties_duplicates_and_errors(algo):
    ax = torch.arange(9, dtype=torch.float32)
    lat = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3).contiguous()
    i_o, d_o = O.knn(lat, lat, 16)
    i_g, d_g = D.knn(cu(lat), cu(lat), 16, algo=algo)
    assert torch.equal(i_g.cpu(), i_o) and torch.equal(d_g.cpu(), d_o)
    b = synth.make_batch(1, 3000, 8, "kitti", config=1, first_pair=77, tiled_frac=0.27)   # FixedResampler duplicates
    p = b["points_src"][:, :, :3].contiguous()
    i_o, d_o = O.knn(p, p, 16)
    i_g, d_g = D.knn(cu(p), cu(p), 16, algo=algo)
    assert (d_o[:, :, 1] == 0).sum() > 500
    assert torch.equal(i_g.cpu(), i_o) and torch.equal(d_g.cpu(), d_o)
    with pytest.raises(D.DeepSIRError):
        D.knn(cu(lat[:, :10]), cu(lat), 16, algo=algo)


@pytest.mark.parametrize("algo", [D.KNN_BRUTE, D.KNN_GRID, D.KNN_TREE, D.KNN_AUTO])
def test_knn_pyramid_matches_nn_search(algo):
    b = synth.make_batch(2, 4096, 8, "kitti", config=1)
    for key in ("points_src", "points_ref"):
        o = O.nn_search_c(b[key], 16, (4, 4, 4, 4))
        g = D.nn_search_cloud(cu(b[key]), 16, (4, 4, 4, 4), algo=algo)
        assert g["xyz"].shape == (2, 5440, 3) and g["sub_idx"].shape == (2, 1360, 16)
        for name in ("xyz", "neigh_idx", "sub_idx", "interp_idx"):
            assert torch.equal(g[name].cpu(), o[name]), (key, name)
    d = D.nn_search({k: cu(v) for k, v in b.items() if k.startswith("points")})
    assert d["points_src_neigh_idx"].dtype == torch.int64 and d["points_ref_interp_idx"].shape == (2, 5440, 1)


def test_nn_search_pair_strided_views_and_side_stream():
    """data['points_src'][:, :, :3] (the slice DataBase.nn_search makes, data_base.py:159) is an inner-sliced view: it is
    passed by stride; a view that cannot be addressed with one point stride is copied BEFORE the fork of the side stream
    (otherwise the reference pyramid could read an unfinished copy)."""
    b = synth.make_batch(3, 3000, 8, "kitti", config=1, first_pair=21)
    wide_s = torch.cat([b["points_src"], torch.randn(3, 3000, 2)], 2)            # [B,N,6]
    wide_r = torch.cat([b["points_ref"], torch.randn(3, 3000, 2)], 2)
    g_s, g_r = D.nn_search_pair(cu(b["points_src"][:, :, :3].contiguous()), cu(b["points_ref"][:, :, :3].contiguous()))
    v_s, v_r = D.nn_search_pair(cu(wide_s)[:, :, :3], cu(wide_r)[:, :, :3])      # strided views, no copy
    t_s, t_r = D.nn_search_pair(cu(wide_s).permute(0, 2, 1).contiguous().permute(0, 2, 1)[:, :, :3],
                                cu(wide_r)[:, ::1, :3].flip(1).flip(1))            # needs a copy
    for name in ("xyz", "neigh_idx", "sub_idx", "interp_idx"):
        assert torch.equal(v_s[name], g_s[name]) and torch.equal(v_r[name], g_r[name]), name
        assert torch.equal(t_s[name], g_s[name]) and torch.equal(t_r[name], g_r[name]), name


def test_knn_grid_degenerate_clouds():
    """Grid path on shapes that stress the cell sizing: tiny clouds, all points identical, collinear points,
    points on a plane, a far outlier (huge bounding box), k = 32."""
    g = torch.Generator().manual_seed(9)
    clouds = {
        "tiny": torch.randn(2, 40, 3, generator=g),
        "identical": torch.ones(1, 600, 3) * 3.25,
        "line": torch.linspace(-5, 5, 700)[None, :, None] * torch.tensor([1.0, 0.5, -2.0]),
        "plane": torch.cat([torch.randn(1, 900, 2, generator=g) * 10, torch.zeros(1, 900, 1)], 2),
        "outlier": torch.cat([torch.randn(1, 1500, 3, generator=g), torch.tensor([[[1e4, -1e4, 5e3]]])], 1),
    }
    for name, c in clouds.items():
        c = c.contiguous()
        for k in (1, 16, 32):
            i_o, d_o = O.knn(c, c, k)
            for algo in (D.KNN_GRID, D.KNN_TREE):
                i_g, d_g = D.knn(cu(c), cu(c), k, algo=algo)
                assert torch.equal(i_g.cpu(), i_o) and torch.equal(d_g.cpu(), d_o), (name, k, algo)
    # queries far outside the support's bounding box
    sup = torch.randn(1, 800, 3, generator=g).contiguous()
    qry = (torch.randn(1, 300, 3, generator=g) * 50).contiguous()
    i_o, d_o = O.knn(sup, qry, 8)
    for algo in (D.KNN_GRID, D.KNN_TREE):
        i_g, d_g = D.knn(cu(sup), cu(qry), 8, algo=algo)
        assert torch.equal(i_g.cpu(), i_o) and torch.equal(d_g.cpu(), d_o)


def test_knn_full_size_properties():
    """C2 size (16384 points): properties that do not need the CPU brute force on the whole cloud + an oracle
    check on a slice of the queries."""
    b = synth.make_batch(2, 16384, 8, "kitti", config=2)
    p = cu(b["points_src"])
    i_g, d_g = D.knn(p, p, 16)
    assert torch.equal(i_g[:, :, 0].cpu(), torch.arange(16384).expand(2, -1))        # self is the nearest
    assert (d_g[:, :, 1:] >= d_g[:, :, :-1]).all()
    sub = b["points_src"][:, :512, :3].contiguous()
    i_o, d_o = O.knn(b["points_src"][:, :, :3].contiguous(), sub, 16)
    assert torch.equal(i_g[:, :512].cpu(), i_o) and torch.equal(d_g[:, :512].cpu(), d_o)
    i_b, d_b = D.knn(p, p, 16, algo=D.KNN_BRUTE)
    assert torch.equal(i_b, i_g) and torch.equal(d_b, d_g)


# ------------------------------------------------------------------------------------------- match
def test_match_dense_golden(golden):
    g = golden("match_dense")
    fs, fr = cu(g["feat_src"]), cu(g["feat_ref"])
    assert torch.allclose(D.match_features_V2(fs, fr, "l2").cpu(), g["l2"], atol=3e-6, rtol=0)
    assert torch.allclose(D.match_features_V2(fs, fr, "euclidean").cpu(), g["euclidean"], atol=3e-6, rtol=0)
    assert torch.allclose(D.match_features_V2(fs, fr, "angle").cpu(), g["angle"], atol=5e-6, rtol=0)
    assert torch.allclose(D.square_distance_V2(fs, fr).cpu(), g["l2"], atol=3e-6, rtol=0)
    nc = D.match_features(fs.permute(0, 2, 1).contiguous(), fr.permute(0, 2, 1).contiguous(), "l2")
    assert torch.allclose(nc.cpu(), g["nc_l2"], atol=3e-6, rtol=0)
    assert torch.allclose(D.feat_dist(fs, fr, "sqeuclidean").cpu(), g["fd_sq"], atol=3e-6, rtol=0)
    assert torch.allclose(D.feat_dist(fs, fr, "cityblock").cpu(), g["fd_city"], atol=2e-5, rtol=0)
    assert torch.allclose(D.feat_dist(fs, fr, "euclidean").cpu(), g["fd_euc"], atol=3e-6, rtol=0)
    # xyz use of square_distance (test.py:145), C = 3, [B,N,C] layout
    a, b = torch.randn(2, 100, 3) * 30, torch.randn(2, 90, 3) * 30
    assert torch.allclose(D.square_distance(cu(a), cu(b)).cpu(), O.square_distance(a, b), atol=2e-3, rtol=1e-6)


@pytest.mark.parametrize("algo", [D.MATCH_FP32, D.MATCH_TC, D.MATCH_AUTO])
def test_match_argmin_golden(golden, algo):
    g = golden("match_argmin_1500")
    b = synth.make_batch(2, 1500, 64, "kitti", config=1)
    idx = D.match_argmin(cu(b["feat_src"]), cu(b["feat_ref"]), algo=algo)
    assert idx.dtype == torch.int64 and torch.equal(idx.cpu(), g["idx_full"])
    g = golden("match_argmin_7000")
    b = synth.make_batch(1, 7000, 32, "3dmatch", config=3)
    assert torch.equal(D.match_argmin(cu(b["feat_src"]), cu(b["feat_ref"]), algo=algo).cpu(), g["idx"])
    g = golden("match_argmin_ties")        # duplicated reference columns: first index wins
    assert torch.equal(D.match_argmin(cu(g["feat_src"]), cu(g["feat_ref"]), algo=algo).cpu(), g["idx"])


@pytest.mark.parametrize("algo", [D.MATCH_FP32, D.MATCH_TC, D.MATCH_AUTO])
@pytest.mark.parametrize("shape", [(2, 64, 1000, 1300), (1, 32, 333, 257), (1, 64, 4096, 4096), (3, 3, 200, 50), (1, 40, 129, 1)])
def test_match_argmin_random_features_vs_oracle(algo, shape):
    """No planted matches: small top-2 gaps.  Rows whose fp64 gap is below fp32 round-off are 'tie-ambiguous'
    (the reference's own sgemm order decides them); every other row must be bit-exact."""
    B, C, J, K = shape
    fs = synth.random_features(B, C, J, 11)
    fr = synth.random_features(B, C, K, 12)
    ref = O.match_argmin(fs, fr)
    got = D.match_argmin(cu(fs), cu(fr), algo=algo).cpu()
    i64, gap = O.match_top2_fp64(fs, fr)
    ambiguous = gap < 2e-6 if K > 1 else torch.zeros_like(ref, dtype=torch.bool)
    assert ambiguous.float().mean() < 0.01
    assert torch.equal(got[~ambiguous], ref[~ambiguous])
    assert torch.equal(got[~ambiguous], i64[~ambiguous])


@pytest.mark.parametrize("algo", [D.MATCH_FP32, D.MATCH_TC, D.MATCH_AUTO])
def test_match_argmin_sliced_views_equal_full(algo):
    """network/model.py:565 passes feat_src[:, :, n*stride:(n+1)*stride]: views must work without copies and the
    chunked result must equal the single fused call."""
    b = synth.make_batch(2, 2600, 64, "kitti", config=1, first_pair=3)
    fs, fr = cu(b["feat_src"]), cu(b["feat_ref"])
    full = D.match_argmin(fs, fr, algo=algo)
    parts = [D.match_argmin(fs[:, :, lo:lo + 1000], fr, algo=algo) for lo in range(0, 2600, 1000)]
    assert torch.equal(torch.cat(parts, 1), full)
    idx, mind = D.match_argmin(fs, fr, return_min=True, algo=algo)
    dense = D.match_features_V2(fs[:, :, :300], fr)
    assert torch.equal(dense.min(dim=2)[1], idx[:, :300])
    assert torch.equal(dense.min(dim=2)[0], mind[:, :300])


@pytest.mark.parametrize("algo", [D.MATCH_FP32, D.MATCH_TC, D.MATCH_AUTO])
def test_match_argmin_full_size_planted(algo):
    """BASELINE config 2 size (16384 x 16384, D=64): planted permutation recovered on every inlier row; the row
    minimum agrees with the library's own dense rows on a sample."""
    b = synth.make_batch(1, 16384, 64, "kitti", config=2)
    fs, fr = cu(b["feat_src"]), cu(b["feat_ref"])
    idx, mind = D.match_argmin(fs, fr, return_min=True, algo=algo)
    inl = b["inlier"][0]
    assert torch.equal(idx[0].cpu()[inl], b["perm"][0][inl])
    rows = torch.arange(0, 16384, 61, device=DEV)
    dense = D.match_features_V2(fs[:, :, rows].contiguous(), fr)
    assert torch.equal(dense.min(dim=2)[1], idx[:, rows])
    assert torch.equal(dense.min(dim=2)[0], mind[:, rows])


def test_match_tc_equals_fp32_and_rarely_rescues():
    """The tcgen05 filter + fp32 refine must return the SAME indices and minima as the fp32 kernel, and the
    exhaustive rescue path must stay (almost) idle — otherwise the tensor-core stage is not doing the work."""
    b = synth.make_batch(2, 8192, 64, "kitti", config=2, first_pair=100)
    fs, fr = cu(b["feat_src"]), cu(b["feat_ref"])
    i_tc, d_tc, n_resc = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_TC, return_rescued=True)
    i_32, d_32 = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_FP32)
    assert torch.equal(i_tc, i_32) and torch.equal(d_tc, d_32)
    assert n_resc <= 2, f"{n_resc} of 16384 planted rows went to the rescue path"
    # random (un-planted) unit features, D = 64 and D = 32, ragged sizes: still identical, rescue rate small
    for (B, C, J, K) in [(1, 64, 5000, 7777), (2, 32, 3001, 4100), (1, 48, 700, 9000), (1, 8, 300, 300)]:
        fs, fr = cu(synth.random_features(B, C, J, 5)), cu(synth.random_features(B, C, K, 6))
        i_tc, d_tc, n_resc = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_TC, return_rescued=True)
        i_32, d_32 = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_FP32)
        assert torch.equal(i_tc, i_32) and torch.equal(d_tc, d_32), (B, C, J, K)
        assert n_resc <= 0.05 * B * J + 1 or C <= 8, (n_resc, B, C, J, K)
    # adversarial: every reference row identical -> every row saturates -> all rescued, first index wins
    fr = cu(synth.random_features(1, 64, 1, 9)).expand(1, 64, 600).contiguous()
    fs = cu(synth.random_features(1, 64, 300, 10))
    i_tc, n_resc = D.match_argmin(fs, fr, algo=D.MATCH_TC, return_rescued=True)
    assert n_resc == 300 and (i_tc == 0).all()
    # un-normalised features with a large dynamic range
    g = torch.Generator().manual_seed(1)
    fs = cu(torch.randn(1, 64, 2000, generator=g) * torch.logspace(-2, 2, 2000)[None, None, :])
    fr = cu(torch.randn(1, 64, 2500, generator=g) * 3.0)
    i_tc, d_tc = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_TC)
    i_32, d_32 = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_FP32)
    assert torch.equal(i_tc, i_32) and torch.equal(d_tc, d_32)


def test_gather_golden(golden):
    g = golden("gather_v3")
    assert torch.equal(D.gather_neighbour_V3(cu(g["inputs"]), cu(g["idx"])).cpu(), g["out"])


# ------------------------------------------------------------------------------------------- soft
@pytest.mark.parametrize("shape", [(2, 32, 500, 640), (1, 64, 130, 257), (2, 32, 1000, 1000)])
def test_match_soft_vs_oracle(shape):
    B, C, J, K = shape
    b = synth.make_batch(B, max(J, K), C, "3dmatch", config=3, first_pair=20)
    fs, fr = b["feat_src"][:, :, :J].contiguous(), b["feat_ref"][:, :, :K].contiguous()
    xyz = b["points_ref"][:, :K, :3].contiguous()
    beta = torch.tensor([10.0, 6.0, 3.0][:B])
    alpha = torch.tensor([0.5, 0.3, 0.1][:B])
    w, y, s, lse = O.soft_correspondence(fs, fr, xyz, beta, alpha)
    y_g, s_g, lse_g = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(beta), cu(alpha))
    assert torch.allclose(lse_g.cpu(), lse, rtol=SOFT_RTOL, atol=1e-5)
    assert torch.allclose(s_g.cpu(), s, rtol=SOFT_RTOL, atol=0)
    assert torch.allclose(y_g.cpu(), y, rtol=SOFT_RTOL, atol=1e-4)
    # weights reconstructed from lse match the oracle's weights to 1e-4 relative where they matter
    a = O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha)
    w_g = torch.exp(a - lse_g.cpu()[:, :, None])
    big = w > 1e-6
    assert ((w_g - w).abs()[big] / w[big]).max() < 5 * SOFT_RTOL
    # scalar alpha form of compute_affinity
    y2, _, _ = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(beta), 0.5)
    _, y2o, _, _ = O.soft_correspondence(fs, fr, xyz, beta, 0.5)
    assert torch.allclose(y2.cpu(), y2o, rtol=SOFT_RTOL, atol=1e-4)


def test_soft_pipeline_pose():
    b = synth.make_batch(2, 800, 32, "3dmatch", config=3, first_pair=40)
    src = b["points_src"][:, :, :3].contiguous()
    ref = b["points_ref"][:, :, :3].contiguous()
    beta = torch.tensor([10.0, 10.0])
    w, y, s, _ = O.soft_correspondence(b["feat_src"], b["feat_ref"], ref, beta, 0.5)
    T_o, _ = O.compute_rigid_transform(src, ref, w)
    y_g, s_g, _ = D.match_soft(cu(b["feat_src"]), cu(b["feat_ref"]), cu(ref), cu(beta), 0.5)
    T_g, inv = D.kabsch_soft(cu(src), y_g, s_g)
    assert not bool(inv)
    assert_pose_close(T_g, T_o)
    T_g2, _ = D.compute_rigid_transform(cu(src), cu(ref), cu(w))   # signature-compatible form ([B,M,N] weights)
    assert_pose_close(T_g2, T_o)


# ------------------------------------------------------------------------------------------- Kabsch / SE3
def test_kabsch_golden(golden):
    g = golden("kabsch2")
    for name in ["planted", "uniform_w", "planar", "reflection", "neg_w", "single_heavy"]:
        T, inv = D.compute_rigid_transform_2(cu(g[name + "_src"]), cu(g[name + "_tgt"]), cu(g[name + "_w"]))
        assert T.shape == (3, 3, 4) and T.dtype == torch.float32 and not bool(inv), name
        assert (torch.det(T[:, :, :3].cpu()) > 0).all(), name
        assert_pose_close(T, g[name + "_T"])


def test_kabsch_planted_and_layouts():
    b = synth.make_batch(4, 16384, 8, "kitti", config=2, first_pair=8)
    src = b["points_src"][:, :, :3].contiguous()
    tgt = torch.stack([b["points_ref"][i, b["perm"][i], :3] for i in range(4)])
    T_o, _ = O.compute_rigid_transform_2(src, tgt, b["weights"])
    T_g, st = D.compute_rigid_transform_2(cu(src), cu(tgt), cu(b["weights"]), return_status=True)
    assert (st == 0).all()
    assert_pose_close(T_g, T_o)
    # fused gather + [B,3,N] layout (the loop's form)
    xs = cu(src).permute(0, 2, 1).contiguous()
    xr = cu(b["points_ref"][:, :, :3]).permute(0, 2, 1).contiguous()
    T_f, st = D.kabsch_gather(xs, xr, cu(b["perm"]), cu(b["weights"]))
    assert torch.equal(T_f, T_g)
    # strided [B,M,3] view of a [B,M,4] tensor (no copy)
    T_v, _ = D.compute_rigid_transform_2(cu(b["points_src"])[:, :, :3], cu(tgt), cu(b["weights"]))
    assert torch.equal(T_v, T_g)
    # row-block sharded moments add up
    m = sum(D.kabsch_moments(cu(src[:, lo:lo + 4096]), cu(tgt[:, lo:lo + 4096]), cu(b["weights"][:, lo:lo + 4096]))
            for lo in range(0, 16384, 4096))
    T_m, st = D.kabsch_from_moments(m)
    assert_pose_close(T_m, T_o)
    assert torch.allclose(m.cpu(), O.kabsch_moments_fp64(src, tgt, b["weights"]), rtol=1e-12, atol=1e-9)


def test_kabsch_degenerate_cases():
    # rank-1 covariance (collinear points): the reference returns a LAPACK-dependent rotation about the line and the centroid
    # translation, without flagging; here the completion is the minimal rotation (identity for a pure shift) - never "invalid"
    line = torch.linspace(0, 1, 50)[None, :, None] * torch.tensor([1.0, 2.0, 3.0])
    T, st = D.compute_rigid_transform_2(cu(line), cu(line + 1.0), cu(torch.ones(1, 50, 1)), return_status=True)
    assert st.item() == 0
    assert_pose_close(T, torch.cat([torch.eye(3), torch.ones(3, 1)], 1)[None])
    Rq = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(2)))[0]
    Rq = Rq * torch.linalg.det(Rq).sign()
    T, st = D.compute_rigid_transform_2(cu(line), cu(line @ Rq.t() + 0.5), cu(torch.ones(1, 50, 1)), return_status=True)
    Tc = T.cpu()[0]
    assert st.item() == 0 and torch.allclose(Tc[:, :3] @ Tc[:, :3].t(), torch.eye(3), atol=1e-5) and torch.linalg.det(Tc[:, :3]) > 0.999
    assert torch.allclose(line[0] @ Tc[:, :3].t() + Tc[:, 3], line[0] @ Rq.t() + 0.5, atol=1e-4)   # the line is mapped exactly
    # rank 0 (all weights zero: centroids 0, covariance 0): identity like the reference, not flagged
    pts = torch.randn(2, 30, 3)
    T, inv = D.compute_rigid_transform_2(cu(pts), cu(pts), cu(torch.zeros(2, 30, 1)))
    assert not bool(inv) and torch.equal(T.cpu(), O.se3_identity(2))
    # a single non-zero weight: zero covariance, the translation between the two points survives
    w1 = torch.zeros(2, 30, 1); w1[:, 7] = 2.0
    T, inv = D.compute_rigid_transform_2(cu(pts), cu(pts + 3.0), cu(w1))
    assert not bool(inv)
    assert_pose_close(T, torch.cat([torch.eye(3), torch.full((3, 1), 3.0)], 1).expand(2, -1, -1))
    bad = pts.clone(); bad[0, 3, 1] = float("nan")
    T, st = D.compute_rigid_transform_2(cu(bad), cu(pts), cu(torch.ones(2, 30, 1)), return_status=True)
    assert st.tolist() == [1, 0]
    # identity / pure translation / 180 degree turn
    T, _ = D.compute_rigid_transform_2(cu(pts), cu(pts + torch.tensor([1.0, -2.0, 3.0])), cu(torch.ones(2, 30, 1)))
    assert_pose_close(T, torch.cat([torch.eye(3), torch.tensor([[1.0], [-2.0], [3.0]])], 1).expand(2, -1, -1))
    Rz = torch.diag(torch.tensor([-1.0, -1.0, 1.0]))
    T, _ = D.compute_rigid_transform_2(cu(pts), cu(pts @ Rz.t()), cu(torch.ones(2, 30, 1)))
    assert_pose_close(T, torch.cat([Rz, torch.zeros(3, 1)], 1).expand(2, -1, -1))


def test_kabsch_gather_index_out_of_range_poisons_the_pair():
    """torch.gather raises on an index outside the reference cloud (network/tools.py:211-221); the fused gather never
    dereferences it and reports the pair through its status instead (invalid_gradient)."""
    g = torch.Generator().manual_seed(8)
    xs, xr = torch.randn(2, 3, 400, generator=g), torch.randn(2, 3, 500, generator=g)
    idx = torch.randint(0, 500, (2, 400), generator=g)
    w = torch.rand(2, 400, generator=g)
    T0, st0 = D.kabsch_gather(cu(xs), cu(xr), cu(idx), cu(w))
    bad = idx.clone(); bad[1, 17] = 500; bad[1, 200] = -3
    T1, st1 = D.kabsch_gather(cu(xs), cu(xr), cu(bad), cu(w))
    assert st0.tolist() == [0, 0] and st1.tolist() == [0, 1]
    assert torch.equal(T1[0], T0[0]) and torch.equal(T1[1].cpu(), O.se3_identity(1)[0])


def test_se3_golden(golden):
    g = golden("se3")
    S = D.se3_torch
    Ta, Tb, pts = cu(g["Ta"]), cu(g["Tb"]), cu(g["pts"])
    assert torch.equal(S.identity(4), g["identity"])
    assert torch.allclose(S.inverse(Ta).cpu(), g["inverse"], atol=1e-6, rtol=0)
    assert torch.allclose(S.concatenate(Ta, Tb).cpu(), g["concat"], atol=1e-6, rtol=0)
    assert torch.allclose(S.transform(Ta, pts).cpu(), g["transform"], atol=2e-5, rtol=0)
    assert torch.allclose(S.transform_V2(Ta, pts.permute(0, 2, 1).contiguous()).cpu(), g["transform_v2"], atol=2e-5, rtol=0)
    T44 = torch.cat([Ta, torch.tensor([0, 0, 0, 1.0], device=DEV).expand(4, 1, 4)], 1)      # (B,4,4) inputs
    assert torch.allclose(S.concatenate(T44, Tb).cpu(), g["concat"], atol=1e-6, rtol=0)
    out, nrm = S.transform(Ta, pts, pts)
    assert torch.allclose(nrm.cpu(), g["pts"] @ g["Ta"][:, :, :3].transpose(1, 2), atol=2e-5, rtol=0)


# ------------------------------------------------------------------------------------------- loop
@pytest.mark.parametrize("algo", [D.MATCH_FP32, D.MATCH_TC, D.MATCH_AUTO])
def test_loop_golden_and_oracle(golden, algo):
    g = golden("loop_oxford_1200")
    b = synth.make_batch(2, 1200, 64, "oxford", config=5)
    xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    tr, pred, xyz, st = D.align_loop(cu(b["feat_src"]), cu(b["feat_ref"]), cu(xs), cu(xr), cu(b["weights"]), 3, algo=algo)
    assert (st == 0).all() and len(tr) == 3
    assert torch.equal(torch.stack(pred).cpu(), g["pred"])
    for i in range(3):
        assert_pose_close(tr[i], g["transforms"][i])
    assert torch.allclose(xyz.cpu(), g["xyz_src_final"], atol=2e-4, rtol=0)
    pp = D.pred_pairs(pred[-1])
    assert pp.dtype == torch.int32 and pp.device.type == "cpu" and pp.shape == (2, 1200, 2)
    # per-iteration form with callbacks == fused form
    tr2, pred2, xyz2, _ = D.align_loop(cu(b["feat_src"]), cu(b["feat_ref"]), cu(xs), cu(xr), cu(b["weights"]), 3,
                                       weight_fn=lambda s, r: cu(b["weights"]), algo=algo)
    assert torch.equal(torch.stack(pred2), torch.stack(pred))
    for i in range(3):
        assert torch.allclose(tr2[i], tr[i], atol=1e-6)


def test_loop_ten_iterations_converges():
    """BASELINE config 5 shape at reduced batch: 10 ICP-style iterations on Oxford-shaped 20k clouds."""
    b = synth.make_batch(1, 20000, 64, "oxford", config=5, first_pair=9)
    xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    tr, pred, xyz, st = D.align_loop(cu(b["feat_src"]), cu(b["feat_ref"]), cu(xs), cu(xr), cu(b["weights"]), 10)
    assert (st == 0).all()
    assert O.rotation_angle_deg(tr[-1].cpu()[:, :, :3], b["transform_gt"][:, :, :3]).max() < 0.5
    # after the first solve the residual transform is ~identity: cumulative transforms stay put
    assert O.rotation_angle_deg(tr[-1].cpu()[:, :, :3], tr[0].cpu()[:, :, :3]).max() < 1e-2
    # transforms[i] is the running composition (model.py:595)
    moved = D.se3_torch.transform_V2(tr[-1], cu(xs))
    assert torch.allclose(moved, xyz, atol=5e-4)


# ------------------------------------------------------------------------------------------- fp16 filter edge cases
def test_match_tc_extreme_scales_and_non_finite_rows():
    """The tensor-core stage rounds sigma-scaled features to fp16; the answer must not depend on the scale of the
    inputs, on rows of zeros, or on non-finite rows (those fall through to the exhaustive fp32 path)."""
    fs0, fr0 = synth.random_features(1, 64, 1500, 21), synth.random_features(1, 64, 2100, 22)
    for scale_s, scale_r in [(1e-6, 1e-6), (1e4, 1e4), (1e-3, 1e2), (37.0, 0.01)]:
        fs, fr = cu(fs0 * scale_s), cu(fr0 * scale_r)
        i_tc, d_tc = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_TC)
        i_32, d_32 = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_FP32)
        assert torch.equal(i_tc, i_32) and torch.equal(d_tc, d_32), (scale_s, scale_r)
    fs, fr = cu(fs0.clone()), cu(fr0.clone())
    fs[0, :, 7] = 0.0                      # a zero source row
    fr[0, :, 11] = 0.0                     # a zero reference row
    fs[0, 3, 100] = float("nan")           # a NaN source row: every distance NaN -> index 0 like the fp32 kernel
    fr[0, 5, 200] = float("inf")           # a non-finite reference row never wins
    i_tc = D.match_argmin(fs, fr, algo=D.MATCH_TC)
    i_32 = D.match_argmin(fs, fr, algo=D.MATCH_FP32)
    assert torch.equal(i_tc, i_32)
    assert (i_tc != 200).all()


def test_filter_timing_diagnostic():
    b = synth.make_batch(2, 4096, 64, "kitti", config=2, first_pair=50)
    tm = {}
    D.match_argmin(cu(b["feat_src"]), cu(b["feat_ref"]), algo=D.MATCH_TC, timing=tm)
    assert tm["span_ns"] > 0 and tm["cycles_per_unit"] > 0


# ------------------------------------------------------------------------------------------- host pipeline
def test_pipeline_equals_direct_calls():
    """RegistrationPipeline (pinned host in -> host out, upload overlapped with compute) returns exactly what the
    individual library calls return on device-resident copies of the same batches."""
    batches = [synth.make_batch(2, 3000, 64, "kitti", config=2, first_pair=10 * i) for i in range(3)]
    host = [dict(points_src=b["points_src"].pin_memory(), points_ref=b["points_ref"].pin_memory(),
                 feat_src=b["feat_src"].pin_memory(), feat_ref=b["feat_ref"].pin_memory(),
                 weights=b["weights"][:, :, 0].contiguous().pin_memory()) for b in batches]
    pipe = D.RegistrationPipeline(DEV, 16, (4, 4, 4, 4), iters=2, depth=2, keep_graph=True)
    outs = list(pipe.run(iter(host)))
    assert len(outs) == 3
    for b, o in zip(batches, outs):
        xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
        xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
        tr, pred, _, st = D.align_loop(cu(b["feat_src"]), cu(b["feat_ref"]), cu(xs), cu(xr), cu(b["weights"]), 2)
        assert torch.equal(o["T"], tr[-1].cpu())
        assert o["pred"].dtype == torch.int32 and torch.equal(o["pred"].long(), pred[-1].cpu())
        assert torch.equal(o["status"], st.cpu())
        g = D.nn_search_cloud(cu(b["points_ref"]), 16, (4, 4, 4, 4))
        for k in g:
            assert torch.equal(o["graph_ref"][k], g[k])


# ------------------------------------------------------------------------------------------- row-block sharding
def test_rowblock_sharding_on_one_device():
    """raw fp64 moments of source-row blocks add up to the moments of the whole pair, and the Kabsch solve
    from the summed moments equals the unsharded solve (ranks emulated as row blocks on one GPU; the collective itself is
    covered by the world-size-2 gloo test)."""
    from deepsir_b200 import dist as DD
    b = synth.make_batch(2, 5000, 64, "kitti", config=4, first_pair=3)
    xs = cu(b["points_src"][:, :, :3].permute(0, 2, 1).contiguous())
    xr = cu(b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous())
    fs, fr, w = cu(b["feat_src"]), cu(b["feat_ref"]), cu(b["weights"][:, :, 0].contiguous())
    idx = D.match_argmin(fs, fr)
    T_full, st = D.kabsch_gather(xs, xr, idx, w)
    mom = torch.zeros(2, 17, dtype=torch.float64, device=DEV)
    parts = []
    for r in range(3):
        lo, hi = DD.row_block(5000, 3, r)
        i_r = D.match_argmin(fs[:, :, lo:hi], fr)                 # a rank matches only its own rows
        parts.append(i_r)
        mom += DD.LibraryOps.moments(xs[:, :, lo:hi].contiguous(), xr, i_r, w[:, lo:hi].contiguous())
    assert torch.equal(torch.cat(parts, 1), idx)
    T_sh, st2 = D.kabsch_from_moments(mom)
    assert (st == 0).all() and (st2 == 0).all()
    assert O.rotation_angle_deg(T_sh.cpu()[:, :, :3], T_full.cpu()[:, :, :3]).max() < 1e-5
    assert (T_sh - T_full).abs().max() < 1e-5
    # the sharded loop with one rank is the fused loop
    tr_a, pred_a, xyz_a, _ = DD.align_rowblock(fs, fr, xs, xr, w, 3)
    tr_b, pred_b, xyz_b, _ = D.align_loop(fs, fr, xs, xr, w, 3)
    assert torch.equal(torch.stack(pred_a), torch.stack(pred_b))
    for a, c in zip(tr_a, tr_b):
        assert_pose_close(a, c.cpu())


# ------------------------------------------------------------------------------------------- KNN consumers / Sinkhorn
def test_graph_ops_golden(golden):
    """gather_neighbour[_V2/_V4], relative_pos_encoding, random_sample, nearest_interpolation against the reference's
    own outputs: pure data movement and one subtraction/sqrt -> bit-exact."""
    g = golden("graph_ops")
    feat, idx, xyz = cu(g["feat"]), cu(g["idx"]), cu(g["xyz"])
    assert torch.equal(D.gather_neighbour_V2(feat, idx).cpu(), g["gather_v2"])
    assert torch.equal(D.gather_neighbour(feat.permute(0, 2, 1).contiguous(), idx).cpu(), g["gather_v1"])
    assert torch.equal(D.gather_neighbour_V4(feat.permute(0, 2, 1).contiguous(), idx[:, :, 0].contiguous()).cpu(), g["gather_v4"])
    rp = D.relative_pos_encoding(xyz, idx).cpu()
    assert torch.equal(rp[:, 1:], g["rel_pos"][:, 1:])
    assert torch.allclose(rp[:, 0], g["rel_pos"][:, 0], rtol=2e-7, atol=0)     # sqrt of a 3-term sum
    assert torch.equal(D.random_sample(feat[:, :, :, None], cu(g["pool"])).cpu(), g["pooled"])
    assert torch.equal(D.nearest_interpolation(cu(g["sub_feat"])[:, :, :, None], cu(g["interp"])).cpu(), g["interpolated"])


def test_graph_ops_on_a_real_pyramid():
    """The KNN pyramid's own index tensors feed the consumers (level-local indices, int64)."""
    b = synth.make_batch(2, 4096, 8, "kitti", config=2, first_pair=5)
    gph = D.nn_search_cloud(cu(b["points_src"]), 16, (4, 4, 4, 4))
    n0 = 4096
    xyz0 = gph["xyz"][:, :n0].permute(0, 2, 1).contiguous()
    nb0 = gph["neigh_idx"][:, :n0]
    rp = D.relative_pos_encoding(xyz0, nb0)
    ref = O.relative_pos_encoding(xyz0.cpu(), nb0.cpu())
    assert torch.equal(rp.cpu()[:, 1:], ref[:, 1:]) and torch.allclose(rp.cpu()[:, 0], ref[:, 0], rtol=2e-7, atol=0)
    assert (rp[:, 0, :, 0] == 0).all()                      # the first neighbour of a point is the point itself
    feat = cu(torch.randn(2, 32, n0, generator=torch.Generator().manual_seed(3)))
    pooled = D.random_sample(feat[:, :, :, None], gph["sub_idx"][:, :n0 // 4])
    assert torch.equal(pooled.cpu(), O.random_sample(feat.cpu()[:, :, :, None], gph["sub_idx"][:, :n0 // 4].cpu()))
    up = D.nearest_interpolation(pooled, gph["interp_idx"][:, :n0])
    assert up.shape == (2, 32, n0, 1)


@pytest.mark.parametrize("shape", [(2, 37, 53), (1, 700, 650), (2, 5, 2000)])
def test_sinkhorn_vs_oracle(golden, shape):
    g = golden("graph_ops")
    la = cu(g["log_alpha"])
    assert torch.allclose(D.sinkhorn(la, 5, True).cpu(), g["sinkhorn_slack_5"], atol=1e-4, rtol=SOFT_RTOL)
    assert torch.allclose(D.sinkhorn(la, 3, False).cpu(), g["sinkhorn_noslack_3"], atol=1e-4, rtol=SOFT_RTOL)
    assert torch.allclose(D.sinkhorn(la, 50, True, eps=1e-2).cpu(), g["sinkhorn_slack_eps"], atol=1e-4, rtol=SOFT_RTOL)
    B, J, K = shape
    a = torch.randn(B, J, K, generator=torch.Generator().manual_seed(J)) * 4
    for slack in (True, False):
        out = D.sinkhorn(cu(a), 5, slack).cpu()
        assert torch.allclose(out, O.sinkhorn(a, 5, slack), atol=1e-4, rtol=SOFT_RTOL)
        # after the last column step every column of exp(out) (+ slack row) sums to 1: <= 1 without the slack entry
        col = torch.exp(out).sum(dim=1)
        assert (col <= 1 + 1e-4).all() and (slack or torch.allclose(col, torch.ones_like(col), atol=1e-4))


def test_sinkhorn_on_the_implicit_matrix():
    """Sinkhorn over the never-materialised affinity equals the reference's sinkhorn() on the materialised
    compute_affinity(match_features_V2()) matrix, and the soft targets / row masses feed the soft Kabsch."""
    b = synth.make_batch(2, 600, 32, "3dmatch", config=3, first_pair=2)
    fs, fr = b["feat_src"], b["feat_ref"][:, :, :500].contiguous()
    xr = b["points_ref"][:, :500, :3].contiguous()
    beta, alpha = torch.tensor([10.0, 6.0]), torch.tensor([0.5, 0.4])
    for slack in (True, False):
        logp = O.sinkhorn(O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha), 5, slack)     # [B,J,K]
        P = torch.exp(logp)
        y, mass, u, v = D.sinkhorn_implicit(cu(fs), cu(fr), cu(xr), cu(beta), cu(alpha), n_iters=5, slack=slack)
        mass_ref = P.sum(dim=2)
        y_ref = (P @ xr) / (mass_ref[:, :, None] + 1e-16)
        assert torch.allclose(mass.cpu(), mass_ref, rtol=5e-4, atol=1e-6)
        assert torch.allclose(y.cpu(), y_ref, rtol=5e-4, atol=1e-4)
        # duals reproduce the matrix itself
        a = O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha)
        assert torch.allclose(a - u.cpu()[:, :, None] - v.cpu()[:, None, :], logp, atol=2e-3, rtol=0)
    T, inv = D.kabsch_soft(cu(b["points_src"][:, :, :3].contiguous()), y, mass)
    assert not bool(inv) and T.shape == (2, 3, 4)


# ------------------------------------------------------------------------------------------- f-1 key points
def test_keypoint_score_and_topk_golden(golden):
    """score_fun / feat_score against the reference's own outputs (fp32 gather-reduce: 1e-5 relative; selection exact
    where the scores are distinct)."""
    g = golden("keypoint_eval")
    feat, xyz, prob, label, neigh = (cu(g[k]) for k in ("feat", "xyz", "prob", "label", "neigh"))
    s = D.score_fun(feat, xyz, prob, label, neigh, label_weights=g["label_weights"].tolist())
    assert torch.allclose(s.cpu(), g["score"], rtol=1e-5, atol=1e-7)
    v, i = D.topk(cu(g["score"]), 200)
    vo, io = O.topk_lower_index(g["score"], 200)
    assert torch.equal(v.cpu(), vo) and torch.equal(i.cpu(), io)
    assert torch.equal(v.cpu(), g["sub_score"])
    f2, x2, l2, s2 = D.feat_score(cu(g["feat"]), xyz, prob, label, neigh, num_sub=200, label_weights=g["label_weights"].tolist())
    assert f2.shape == g["sub_feat"].shape and x2.shape == g["sub_xyz"].shape and l2.shape == g["sub_label"].shape
    assert torch.allclose(s2.cpu(), g["sub_score"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n,k", [(1000, 1), (1000, 1000), (20000, 4096), (333, 17)])
def test_topk_ties_nan_and_sizes(n, k):
    g = torch.Generator().manual_seed(n + k)
    s = torch.randint(0, 7, (3, n), generator=g).float()          # heavy ties: equal values go to the lower index
    s[1, ::5] = 0.0
    s[2, n // 2] = float("nan")                                    # NaN ranks first like torch.topk
    v, i = D.topk(cu(s), k)
    vo, io = O.topk_lower_index(torch.nan_to_num(s, nan=float("inf")), k)
    assert torch.equal(i.cpu(), io)
    assert torch.equal(torch.nan_to_num(v.cpu(), nan=float("inf")), vo)
    tv, _ = torch.topk(s, k, dim=-1, largest=True)
    assert torch.equal(torch.nan_to_num(v.cpu(), nan=-1.0), torch.nan_to_num(tv, nan=-1.0))


def test_keypoint_score_c2_level0_properties():
    """Full C2 cloud size: scores are finite, zero wherever the gate closes, invariant to a permutation of the neighbours."""
    b = synth.make_batch(1, 16384, 8, "kitti", config=2, first_pair=9)
    gph = D.nn_search_cloud(cu(b["points_src"]), 16, (4, 4, 4, 4))
    xyz = cu(b["points_src"][:, :, :3].permute(0, 2, 1).contiguous())
    gen = torch.Generator().manual_seed(1)
    feat = cu(torch.rand(1, 64, 16384, generator=gen))
    prob = cu(torch.rand(1, 1, 16384, generator=gen))
    label = cu(torch.randint(0, 19, (1, 1, 16384), generator=gen))
    nb = gph["neigh_idx"][:, :16384]
    s = D.score_fun(feat, xyz, prob, label, nb)
    assert torch.isfinite(s).all() and (s >= 0).all()
    so = O.score_fun(feat.cpu(), xyz.cpu(), prob.cpu(), label.cpu(), nb.cpu(), D.keypoint.KITTI_LABEL_WEIGHTS)
    assert torch.allclose(s.cpu(), so, rtol=1e-5, atol=1e-7)
    s2 = D.score_fun(feat, xyz, prob, label, nb.flip(-1).contiguous())
    assert torch.allclose(s, s2, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------- f-4 evaluation
def test_eval_golden(golden):
    g = golden("keypoint_eval")
    c = D.metrics.find_correct_correspondence([cu(g["pos0"]), cu(g["pos1"])], cu(g["pred_pairs"]), hash_seed=1024)
    assert torch.equal(c.cpu(), g["correct"].bool())
    c2 = D.metrics.find_correct_correspondence([cu(g["pos0"]), cu(g["pos1"])], cu(g["pred_pairs"]), len_batch=[(1024, 1024), (1024, 900)])
    assert torch.equal(c2.cpu(), g["correct"].bool())
    pe = D.metrics.pose_errors(cu(g["transform_pred"]), cu(g["transform_gt"]), 2.0, 5.0)
    assert torch.allclose(pe["err_r_deg"].cpu(), g["err_r_deg"], atol=ROT_TOL_DEG)
    assert torch.allclose(pe["err_t"].cpu(), g["err_t"], atol=TRANS_TOL_M)
    assert torch.equal(pe["succ"].cpu(), g["succ"].bool())
    assert torch.allclose(pe["rte"].cpu().double(), g["rte_rre"][:, 1].double(), atol=TRANS_TOL_M)
    assert torch.allclose(pe["rre"].cpu().double(), g["rte_rre"][:, 2].double(), atol=ROT_TOL_DEG)
    data = dict(transform_gt=cu(g["transform_gt"]), points_src=cu(g["points_src"]), points_ref=cu(g["points_ref"]))
    m = D.metrics.compute_metrics(data, cu(g["transform_pred"]), 2.0, 5.0)
    assert torch.allclose(m["chamfer_dist"].cpu(), g["chamfer_dist"], rtol=1e-5, atol=1e-8)


def test_correspondence_check_empty_and_collisions():
    """np.isin semantics: an empty positive list gives all-false; keys collide exactly like _hash when an index exceeds
    the seed (loss.py:280-294)."""
    pred = torch.tensor([[[0, 5], [1, 2], [7, 0]]], dtype=torch.int32)
    pos = [torch.tensor([[5, 4], [1, 2]], dtype=torch.int32)]       # seed 5: key(0,5) = 25 = key(5,4)
    c = D.metrics.find_correct_correspondence([cu(pos[0])], cu(pred), hash_seed=5)
    assert c.cpu().tolist() == O.find_correct_correspondence([pos[0].numpy()], pred.numpy(), hash_seed=5).tolist() == [[True, True, False]]
    c = D.metrics.find_correct_correspondence([torch.zeros(0, 2, dtype=torch.int32, device=DEV)], cu(pred), hash_seed=5)
    assert not c.any()


def test_loop_to_metrics_without_host_sync():
    """The loop's last pose and correspondences feed the evaluation entirely on the device: ground-truth matches are found,
    the pose error of a planted pair is small."""
    b = synth.make_batch(2, 4096, 64, "kitti", config=2, first_pair=21)
    xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    tr, pred, _, _ = D.align_loop(cu(b["feat_src"]), cu(b["feat_ref"]), cu(xs), cu(xr), cu(b["weights"]), 2)
    pairs = torch.stack([torch.arange(4096, device=DEV).expand(2, -1), pred[-1]], dim=-1).int()
    pos = [torch.stack([torch.arange(4096), b["perm"][i]], 1).int() for i in range(2)]
    corr = D.metrics.find_correct_correspondence([cu(p) for p in pos], pairs, hash_seed=4096)
    assert corr.float().mean().item() > 0.85                       # 10 % planted outliers
    pe = D.metrics.pose_errors(tr[-1], cu(b["transform_gt"]), 2.0, 5.0)
    assert pe["succ"].all() and pe["err_r_deg"].max().item() < 0.5 and pe["err_t"].max().item() < 0.3


# ------------------------------------------------------------------------------------------- CUDA graphs, full sizes
def test_graph_capture_replays_the_step_bit_exactly():
    """Both KNN pyramids (with the library's internal fork/join) and a 2-iteration loop captured into ONE CUDA graph:
    replays on new inputs equal the eager calls bit for bit."""
    mk = lambda first: {k: cu(v) for k, v in synth.make_batch(2, 4096, 64, "kitti", config=2, first_pair=first).items()
                        if k in ("points_src", "points_ref", "feat_src", "feat_ref", "weights")}
    a, b = mk(60), mk(70)
    a["weights"], b["weights"] = a["weights"][:, :, 0].contiguous(), b["weights"][:, :, 0].contiguous()
    g = D.GraphedRegistration(a, 16, (4, 4, 4, 4), iters=2)
    for batch in (b, a, b):
        out = g.step(batch)
        torch.cuda.synchronize()
        gs, gr = D.nn_search_pair(batch["points_src"], batch["points_ref"], 16, (4, 4, 4, 4))
        xs = batch["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
        xr = batch["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
        tr, pred, xyz, st = D.align_loop(batch["feat_src"], batch["feat_ref"], xs, xr, batch["weights"], 2)
        assert torch.equal(out["T"], torch.stack(tr)) and torch.equal(out["pred"], torch.stack(pred))
        assert torch.equal(out["status"], st) and torch.equal(out["xyz_src"], xyz)
        for k in gs:
            assert torch.equal(out["graph_src"][k], gs[k]) and torch.equal(out["graph_ref"][k], gr[k])


def test_match_c4_full_size_properties():
    """BASELINE config 4 size (131072 x 131072, D=64) on one device: planted inliers are recovered, a row block matched
    alone (what a rank of the row-block sharding does) equals the same rows of the full match, and the fp32 kernel agrees
    on a sample of rows."""
    b = synth.make_batch(1, 131072, 64, "kitti", config=4, first_pair=0)
    fs, fr = cu(b["feat_src"]), cu(b["feat_ref"])
    idx = D.match_argmin(fs, fr, algo=D.MATCH_TC)
    ok = (idx.cpu() == b["perm"])
    assert ok[b["inlier"]].float().mean().item() > 0.999
    lo, hi = 70000, 70000 + 4096
    part = D.match_argmin(fs[:, :, lo:hi], fr, algo=D.MATCH_TC)
    assert torch.equal(part, idx[:, lo:hi])
    exact = D.match_argmin(fs[:, :, lo:hi], fr, algo=D.MATCH_FP32)
    assert torch.equal(exact, part)


@pytest.mark.parametrize("shape", [(2, 32, 3000, 2777), (1, 20, 1500, 4100), (3, 32, 700, 1100), (2, 64, 1500, 1300), (1, 48, 2100, 1000)])
def test_match_soft_tensor_core_path(shape):
    """The tcgen05 soft match (fp16 x2 split contraction + online softmax in the TMEM epilogue; C <= 64 and B*J*K >= 2e6):
    lse, soft targets and reconstructed weights against the fp32 oracle, with and without a column bias, J/K not
    multiples of the tile."""
    B, C, J, K = shape
    b = synth.make_batch(B, max(J, K), C, "3dmatch", config=3, first_pair=50)
    fs, fr = b["feat_src"][:, :, :J].contiguous(), b["feat_ref"][:, :, :K].contiguous()
    xyz = b["points_ref"][:, :K, :3].contiguous()
    beta = torch.tensor([10.0, 6.0, 3.0][:B])
    alpha = torch.tensor([0.5, 0.3, 0.1][:B])
    w, y, s, lse = O.soft_correspondence(fs, fr, xyz, beta, alpha)
    y_g, s_g, lse_g = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(beta), cu(alpha))
    assert torch.allclose(lse_g.cpu(), lse, rtol=SOFT_RTOL, atol=1e-5)
    assert torch.allclose(y_g.cpu(), y, rtol=SOFT_RTOL, atol=1e-4)
    a = O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha)
    w_g = torch.exp(a - lse_g.cpu()[:, :, None])
    big = w > 1e-6
    assert ((w_g - w).abs()[big] / w[big]).max() < 5 * SOFT_RTOL
    # column bias (the Sinkhorn sweeps): lse_j = log sum_k exp(a_jk + bias_k)
    bias = torch.randn(B, K, generator=torch.Generator().manual_seed(K)) * 2
    _, _, lse_b = D.match_soft(cu(fs), cu(fr), None, cu(beta), cu(alpha), col_bias=cu(bias))
    ref = torch.logsumexp(a + bias[:, None, :], dim=2)
    assert torch.allclose(lse_b.cpu(), ref, rtol=SOFT_RTOL, atol=1e-5)


@pytest.mark.parametrize("shape,topk", [((2, 32, 700, 900), 4), ((1, 64, 300, 2500), 16), ((2, 32, 2100, 1000), 8),
                                        ((2, 32, 900, 5000), 32), ((1, 48, 1500, 2777), 20)])
def test_match_soft_topk(shape, topk):
    """Top-k soft correspondences: the k largest weights of every row, descending, ties to the lower index; equal to the
    top-k of the reference's materialised softmax (matchnet.py:195-208,259)."""
    B, C, J, K = shape
    b = synth.make_batch(B, max(J, K), C, "3dmatch", config=3, first_pair=77)
    fs, fr = b["feat_src"][:, :, :J].contiguous(), b["feat_ref"][:, :, :K].contiguous()
    fr[:, :, 5] = fr[:, :, 3]                                      # an exactly duplicated reference point: tie -> lower index
    xyz = b["points_ref"][:, :K, :3].contiguous()
    beta, alpha = torch.tensor([10.0, 6.0][:B]), torch.tensor([0.5, 0.3][:B])
    w, y, s, lse = O.soft_correspondence(fs, fr, xyz, beta, alpha)
    y_g, _, lse_g, ti, tw = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(beta), cu(alpha), topk=topk)
    assert torch.allclose(lse_g.cpu(), lse, rtol=SOFT_RTOL, atol=1e-5) and torch.allclose(y_g.cpu(), y, rtol=SOFT_RTOL, atol=1e-4)
    a = O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha)
    ti, tw = ti.cpu(), tw.cpu()
    assert (tw[:, :, 1:] <= tw[:, :, :-1]).all()
    assert torch.allclose(tw, torch.gather(w, 2, ti), rtol=5 * SOFT_RTOL, atol=1e-7)
    ref_w, _ = torch.topk(w, topk, dim=2)
    assert torch.allclose(tw, ref_w, rtol=5 * SOFT_RTOL, atol=1e-7)
    # index sets: equal wherever the k-th and (k+1)-th affinities are separated by more than fp32 round-off
    srt, order = torch.sort(a, dim=2, descending=True, stable=True)
    clear = (srt[:, :, topk - 1] - srt[:, :, topk]) > 1e-4
    same = (torch.sort(ti, dim=2)[0] == torch.sort(order[:, :, :topk], dim=2)[0]).all(dim=2)
    assert same[clear].all() and clear.float().mean() > 0.9
    has3 = (ti == 3).any(dim=2) & (ti == 5).any(dim=2)            # the duplicated pair appears as (3 before 5)
    pos3 = (ti == 3).float().argmax(dim=2)
    pos5 = (ti == 5).float().argmax(dim=2)
    assert (pos3[has3] < pos5[has3]).all()


def _soft_topk_raw(fs, fr, beta, alpha, topk, bias=None):
    """dsir_match_soft(topk) through ctypes with a workspace the test keeps: (idx, w, lse, exhaustive rows or None)."""
    from deepsir_b200 import _lib as L
    lib = L.lib()
    B, C, J = fs.shape
    K = fr.shape[2]
    (f1, _a), (f2, _b) = L.feat_cn(fs), L.feat_cn(fr)
    lse = torch.empty(B, J, device=DEV)
    ti = torch.empty(B, J, topk, dtype=torch.int64, device=DEV)
    tw = torch.empty(B, J, topk, device=DEV)
    ws = L.workspace(lib.dsir_match_soft_topk_workspace_bytes(B, C, J, K, topk), fs.device)
    L.check(lib.dsir_match_soft(f1, f2, B, C, J, K, beta.data_ptr(), alpha.data_ptr(), L.ptr(bias), None, None, lse.data_ptr(),
                                topk, ti.data_ptr(), tw.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(fs.device)), "soft")
    ex = None
    if bias is None and lib.dsir_match_soft_topk_fused(B, C, J, K, topk):
        import ctypes
        out = ctypes.c_int32(-1)
        L.check(lib.dsir_match_soft_topk_exhaustive_rows(ws.data_ptr(), ws.numel(), B, C, J, K, topk, ctypes.byref(out),
                                                         L.stream_ptr(fs.device)), "exhaustive_rows")
        ex = out.value
    return ti.cpu(), tw.cpu(), lse.cpu(), ex


@pytest.mark.parametrize("shape,topk,fused", [
    ((2, 32, 900, 5000), 32, True),      # C3 width: 40 units >= 1.25 * 32 -> 128-column granules
    ((1, 64, 300, 16384), 5, True),      # C2 width, norm K-step skipped (unit features, K % 128 == 0)
    ((2, 32, 2100, 1000), 8, True),      # 8 units < 10 -> 32-column granules
    ((1, 16, 3000, 2777), 32, True),     # ragged K (folded norm, zero-filled tail), 32-column granules
    ((3, 20, 1300, 1153), 1, True),      # k = 1
    ((2, 32, 4000, 600), 32, False),     # 19 granules of 32 < 40: materialising route
])
def test_match_soft_topk_fused_equals_materialised(shape, topk, fused):
    """The two-sweep tensor-core top-k (no J x K object) returns the bits of the materialising route (exact fp32 distance
    chunks + one warp per row), which a zero column bias selects; few rows need the exhaustive pass."""
    from deepsir_b200 import _lib as L
    B, C, J, K = shape
    assert bool(L.lib().dsir_match_soft_topk_fused(B, C, J, K, topk)) == fused
    b = synth.make_batch(B, max(J, K), C, "3dmatch", config=3, first_pair=91)
    fs, fr = cu(b["feat_src"][:, :, :J].contiguous()), cu(b["feat_ref"][:, :, :K].contiguous())
    fr[:, :, 7] = fr[:, :, 2]                                    # exact duplicate: tie -> lower index first
    fr[:, :, K - 1] = fr[:, :, 2]
    beta, alpha = cu(torch.tensor([10.0, 6.0, 30.0][:B])), cu(torch.tensor([0.5, 0.3, 0.0][:B]))
    ti, tw, lse, ex = _soft_topk_raw(fs, fr, beta, alpha, topk)
    ti0, tw0, lse0, _ = _soft_topk_raw(fs, fr, beta, alpha, topk, bias=torch.zeros(B, K, device=DEV))
    assert torch.equal(lse, lse0)
    assert torch.equal(ti, ti0) and torch.equal(tw, tw0)
    if fused:
        assert ex is not None and ex <= 0.01 * B * J, ex


def test_match_soft_topk_fused_degenerate_rows():
    """Mass ties (every reference point equal), beta = 0 (no distance order), a NaN source row and a NaN reference column:
    the fused route sends such rows through its exhaustive pass and still returns the materialising route's bits."""
    B, C, J, K, topk = 3, 32, 600, 4096, 16
    g = torch.Generator().manual_seed(5)
    fs = torch.nn.functional.normalize(torch.randn(B, C, J, generator=g), dim=1)
    fr = torch.nn.functional.normalize(torch.randn(B, C, K, generator=g), dim=1)
    fr[0] = fr[0, :, :1]                                           # batch 0: all K reference points identical
    fs[2, :, 11] = float("nan")
    fr[2, :, 100] = float("nan")
    beta, alpha = cu(torch.tensor([8.0, 0.0, 12.0])), cu(torch.tensor([0.2, 0.5, 0.4]))
    fs, fr = cu(fs), cu(fr)
    ti, tw, lse, ex = _soft_topk_raw(fs, fr, beta, alpha, topk)
    ti0, tw0, lse0, _ = _soft_topk_raw(fs, fr, beta, alpha, topk, bias=torch.zeros(B, K, device=DEV))
    assert torch.equal(ti, ti0)
    assert torch.equal(torch.nan_to_num(tw, nan=-1.0), torch.nan_to_num(tw0, nan=-1.0))
    assert torch.equal(ti[0], torch.arange(topk).expand(J, topk))          # all tied: the first k columns
    assert torch.equal(ti[1], torch.arange(topk).expand(J, topk))          # beta = 0: every weight equal
    assert not (ti[2] == 100).any()
    assert ex >= 2 * J                                                       # batches 0 and 1 went through the exhaustive pass


@pytest.mark.parametrize("shape,beta", [((64, 16, 200, 180), 10.0), ((1, 32, 130, 20001), 25.0), ((2, 64, 1111, 1000), 100.0),
                                        ((4, 8, 777, 777), 1.0)])
def test_match_soft_tensor_core_edge_shapes(shape, beta):
    """Tensor-core soft path on awkward shapes: many tiny pairs, one short/very wide pair (K-splits), a sharp softmax
    (beta = 100: weights span > 100 binades), few channels.  Includes rows that are exact copies of reference rows
    (d = 0: the weight concentrates on one column)."""
    B, C, J, K = shape
    g = torch.Generator().manual_seed(J + K)
    fs = torch.nn.functional.normalize(torch.randn(B, C, J, generator=g), dim=1)
    fr = torch.nn.functional.normalize(torch.randn(B, C, K, generator=g), dim=1)
    fs[:, :, : min(J, K) // 2] = fr[:, :, : min(J, K) // 2]
    xyz = torch.rand(B, K, 3, generator=g) * 3
    bt = torch.full((B,), beta)
    w, y, s, lse = O.soft_correspondence(fs, fr, xyz, bt, 0.5)
    y_g, _, lse_g = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(bt), 0.5)
    # the bar itself, for every beta: |delta lse| <= 1e-4 is a 1e-4 relative error on every weight of the row.  Beyond
    # beta * max|f|^2 = 32 the library hands the batch element to its exact fp32 kernel by itself (soft_pick_kernel).
    assert torch.allclose(lse_g.cpu(), lse, rtol=SOFT_RTOL, atol=1e-4)
    assert torch.allclose(y_g.cpu(), y, rtol=SOFT_RTOL, atol=1e-4)
    assert torch.isfinite(y_g).all() and torch.isfinite(lse_g).all()


def test_match_soft_picks_the_exact_kernel_per_batch_element():
    """One call, three batch elements: beta = 10 (tensor-core split), beta = 200 and un-normalised features with
    |f|^2 = 25 at beta = 10 (both beyond the bound of the split -> exact fp32 kernel, chosen on the device).  All three
    meet 1e-4."""
    g = torch.Generator().manual_seed(77)
    B, C, J, K = 3, 32, 1500, 1400
    fs = torch.nn.functional.normalize(torch.randn(B, C, J, generator=g), dim=1)
    fr = torch.nn.functional.normalize(torch.randn(B, C, K, generator=g), dim=1)
    fs[2] *= 5.0; fr[2] *= 5.0
    fs[:, :, :300] = fr[:, :, :300]
    xyz = torch.rand(B, K, 3, generator=g) * 3
    bt = torch.tensor([10.0, 200.0, 10.0])
    w, y, s, lse = O.soft_correspondence(fs, fr, xyz, bt, 0.5)
    y_g, _, lse_g = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(bt), 0.5)
    assert torch.allclose(lse_g.cpu(), lse, rtol=SOFT_RTOL, atol=1e-4)
    assert torch.allclose(y_g.cpu(), y, rtol=SOFT_RTOL, atol=1e-4)


def test_sinkhorn_implicit_tensor_core_sweeps():
    """Sinkhorn on the never-materialised affinity with the tcgen05 sweeps (operands prepared once per direction and re-used
    across the half-steps): equals the reference's sinkhorn() on the materialised matrix."""
    b = synth.make_batch(2, 1500, 32, "3dmatch", config=3, first_pair=11)
    fs, fr = b["feat_src"], b["feat_ref"][:, :, :1300].contiguous()
    xr = b["points_ref"][:, :1300, :3].contiguous()
    beta, alpha = torch.tensor([10.0, 6.0]), torch.tensor([0.5, 0.4])
    for slack in (True, False):
        logp = O.sinkhorn(O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha), 5, slack)     # [B,J,K]
        P = torch.exp(logp)
        y, mass, u, v = D.sinkhorn_implicit(cu(fs), cu(fr), cu(xr), cu(beta), cu(alpha), n_iters=5, slack=slack)
        mass_o = P.sum(dim=2)
        y_o = (P @ xr) / mass_o[:, :, None]
        assert torch.allclose(mass.cpu(), mass_o, rtol=5e-4, atol=1e-6)
        assert torch.allclose(y.cpu(), y_o, rtol=5e-4, atol=2e-4)


def test_log_optimal_transport_golden(golden):
    """log_optimal_transport + log_sinkhorn_iterations (network/matchnet.py:827-856): the reference's own outputs."""
    z = golden("log_ot")
    for key, alpha, it in (("out_bin1_it20", z["bin1"], 20), ("out_bin2_it3", -0.7, 3), ("out_it0", 0.25, 0)):
        out = D.log_optimal_transport(cu(z["scores"]), cu(alpha) if isinstance(alpha, torch.Tensor) else alpha, it)
        assert out.shape == (2, 42, 58)
        assert torch.allclose(out.cpu(), z[key], rtol=1e-5, atol=1e-4), key
    g = torch.Generator().manual_seed(5)
    sc = torch.randn(3, 700, 513, generator=g) * 4                           # ragged sizes, sharper scores
    assert torch.allclose(D.log_optimal_transport(cu(sc), 0.5, 7).cpu(), O.log_optimal_transport(sc, 0.5, 7), rtol=1e-5, atol=1e-4)


def test_log_optimal_transport_implicit_tensor_core_sweeps():
    """The same OT on the never-materialised affinity (fused tcgen05 sweeps, dustbin row / column handled as vectors):
    potentials, row masses and soft targets equal those of the reference's log_optimal_transport on the materialised matrix."""
    b = synth.make_batch(2, 1500, 32, "3dmatch", config=3, first_pair=13)
    fs, fr = b["feat_src"], b["feat_ref"][:, :, :1300].contiguous()
    xr = b["points_ref"][:, :1300, :3].contiguous()
    beta, alpha = torch.tensor([10.0, 6.0]), torch.tensor([0.5, 0.4])
    aff = O.compute_affinity(beta, O.match_features_V2(fs, fr), alpha)
    M, N = 1500, 1300
    for bin_score, iters in ((0.3, 5), (-1.0, 2)):
        Z = O.log_optimal_transport(aff, bin_score, iters)                   # [B, M+1, N+1], times (M + N)
        P = torch.exp(Z[:, :M, :N])
        mass_o = P.sum(dim=2)
        y_o = (P @ xr) / mass_o[:, :, None]
        y, mass, u, v = D.log_optimal_transport_implicit(cu(fs), cu(fr), cu(xr), cu(beta), cu(alpha), bin_score, iters)
        assert u.shape == (2, M + 1) and v.shape == (2, N + 1)
        assert torch.allclose(mass.cpu(), mass_o, rtol=5e-4, atol=1e-6)
        assert torch.allclose(y.cpu(), y_o, rtol=5e-4, atol=2e-4)
        # the potentials reproduce the dustbin column of the reference's result: Z[j, N] = bin + u_j + v_N - norm
        norm = -np.log(M + N)
        assert torch.allclose((bin_score + u[:, :M] + v[:, N:] - norm).cpu(), Z[:, :M, N], rtol=1e-4, atol=5e-4)


def test_match_argmin_hint_never_changes_the_result():
    """dsir_match_argmin_hint: a correct, a partly wrong, a random and an out-of-range prior all give the indices of the
    unhinted call (the hint only tightens the filter)."""
    b = synth.make_batch(2, 6000, 64, "kitti", config=2, first_pair=31)
    fs, fr = cu(b["feat_src"]), cu(b["feat_ref"])
    base, dmin = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_TC)
    g = torch.Generator().manual_seed(4)
    wrong = base.clone()
    wrong[:, ::3] = cu(torch.randint(0, 6000, (2, 2000), generator=g))
    for prior in (base, wrong, cu(torch.randint(0, 6000, (2, 6000), generator=g)), torch.full_like(base, -1), torch.full_like(base, 10**6)):
        idx, dm = D.match_argmin(fs, fr, return_min=True, algo=D.MATCH_TC, prior=prior)
        assert torch.equal(idx, base) and torch.equal(dm, dmin)
    rnd_s, rnd_r = cu(synth.random_features(1, 64, 3000, 5)), cu(synth.random_features(1, 64, 3000, 6))   # small top-2 gaps
    base = D.match_argmin(rnd_s, rnd_r, algo=D.MATCH_TC)
    assert torch.equal(D.match_argmin(rnd_s, rnd_r, algo=D.MATCH_TC, prior=base), base)
    assert torch.equal(D.match_argmin(rnd_s, rnd_r, algo=D.MATCH_TC, prior=torch.zeros_like(base)), base)


def test_c_abi_error_codes_on_device():
    """The C ABI never throws: bad arguments, unsupported shapes and short workspaces come back as negative codes with a
    message (dsir_strerror), and the mirror turns them into DeepSIRError."""
    import ctypes
    from deepsir_b200 import _lib as L
    lib = D.lib()
    fs, fr = cu(synth.random_features(1, 64, 700, 1)), cu(synth.random_features(1, 64, 600, 2))
    (a, _k1), (b, _k2) = L.feat_cn(fs), L.feat_cn(fr)
    idx = torch.empty(1, 700, dtype=torch.int64, device=DEV)
    st = L.stream_ptr(torch.device(DEV))
    ws = L.workspace(lib.dsir_match_argmin_workspace_bytes(1, 64, 700, 600, D.MATCH_TC), torch.device(DEV))
    assert lib.dsir_match_argmin(a, b, 1, 64, 700, 600, None, None, ws.data_ptr(), ws.numel(), D.MATCH_TC, st) == -1      # null output
    assert lib.dsir_match_argmin(a, b, 1, 64, 700, 600, idx.data_ptr(), None, ws.data_ptr(), 64, D.MATCH_TC, st) == -3   # workspace
    assert lib.dsir_match_argmin(a, b, 1, 100, 700, 600, idx.data_ptr(), None, ws.data_ptr(), ws.numel(), D.MATCH_TC, st) == -2  # C > 64 on the tensor path
    assert b"workspace" in lib.dsir_strerror(-3)
    pts = cu(torch.rand(1, 10, 3))
    with pytest.raises(D.DeepSIRError):
        D.knn(pts, pts, 16)                                                       # fewer support points than k
    with pytest.raises(D.DeepSIRError):
        D.match_soft(fs, fr, cu(torch.rand(1, 600, 3)), cu(torch.tensor([10.0])), 0.5, topk=33)   # top-k > 32
    with pytest.raises(D.DeepSIRError):
        D.topk(cu(torch.rand(2, 50)), 51)                                          # k > N
    assert lib.dsir_match_argmin(a, b, 1, 64, 700, 600, idx.data_ptr(), None, ws.data_ptr(), ws.numel(), D.MATCH_TC, st) == 0   # still usable
    torch.cuda.synchronize()


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 3, 1, 500), (2, 64, 300, 1), (3, 5, 129, 127), (1, 64, 513, 129), (70, 16, 40, 33)])
def test_degenerate_sizes_all_paths(shape):
    """One row, one column, one channel, sizes one off the tile edges, many tiny pairs; un-normalised features; both argmin
    algorithms and the soft path."""
    B, C, J, K = shape
    g = torch.Generator().manual_seed(J * 1000 + K)
    fs, fr = torch.randn(B, C, J, generator=g), torch.randn(B, C, K, generator=g)
    d = O.match_features_V2(fs, fr)
    ref = O.match_argmin(fs, fr)
    for algo in (D.MATCH_FP32, D.MATCH_TC):
        idx = D.match_argmin(cu(fs), cu(fr), algo=algo).cpu()
        assert torch.allclose(torch.gather(d, 2, idx[:, :, None]), torch.gather(d, 2, ref[:, :, None]), atol=1e-5)
    xyz = torch.rand(B, K, 3, generator=g)
    beta = torch.full((B,), 2.0)
    y, s, lse = D.match_soft(cu(fs), cu(fr), cu(xyz), cu(beta), 0.5)
    w, yo, so, lo = O.soft_correspondence(fs, fr, xyz, beta, 0.5)
    assert torch.allclose(lse.cpu(), lo, rtol=SOFT_RTOL, atol=1e-5) and torch.allclose(y.cpu(), yo, rtol=SOFT_RTOL, atol=1e-4)
