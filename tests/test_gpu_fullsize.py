"""Parity against the ORACLE at the BASELINE.json sizes (one pair / cloud per case, so that the CPU side stays in
seconds), plus the strongest independent pin available for the KNN (whose reference kernel, torch_points_kernels, is
absent): an exact kd-tree (scipy cKDTree) with an fp64 k-th / (k+1)-th gap classifier.

  C2  16384 x 16384, D = 64   hard correspondences vs O.match_argmin, fp64 tie classifier, planted AND random features
  C2  KNN pyramid of a 16384-point cloud: all 21760 rows vs O.nn_search_c
  C3  5000 x 5000, D = 32     soft correspondence vs O.soft_correspondence, 1e-4
  C5  20000 points, 10 iterations of re-match / re-solve vs O.align_loop
"""
import numpy as np
import pytest
import torch

import deepsir_b200 as D
from deepsir_b200 import synth
from oracle import deepsir_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(t):
    return t.to(DEV)


@pytest.mark.parametrize("features", ["planted", "random"])
@pytest.mark.parametrize("algo", [D.MATCH_TC, D.MATCH_FP32])
def test_c2_argmin_vs_oracle(features, algo):
    if features == "planted":
        b = synth.make_batch(1, 16384, 64, "kitti", config=2, first_pair=5)
        fs, fr = b["feat_src"], b["feat_ref"]
    else:                                   # un-planted unit features: small top-2 gaps (learned-descriptor regime)
        fs, fr = synth.random_features(1, 64, 16384, 21), synth.random_features(1, 64, 16384, 22)
    ref = O.match_argmin(fs, fr)            # the reference's 6000-row chunking + torch.min (model.py:558-569)
    got = D.match_argmin(cu(fs), cu(fr), algo=algo).cpu()
    i64, gap = O.match_top2_fp64(fs, fr)
    ambiguous = gap < 2e-6
    assert ambiguous.float().mean() < 0.01
    assert torch.equal(got[~ambiguous], ref[~ambiguous])
    assert torch.equal(got[~ambiguous], i64[~ambiguous])
    # an ambiguous row may go either way, but only between its two fp64-nearest columns
    if ambiguous.any():
        fs64, fr64 = fs.double(), fr.double()
        rows = ambiguous[0].nonzero()[:, 0]
        src = fs64[0][:, rows]
        d_got = ((src - fr64[0][:, got[0, rows]]) ** 2).sum(0)
        d_min = ((src - fr64[0][:, i64[0, rows]]) ** 2).sum(0)
        assert (d_got - d_min <= 4e-6).all()


def test_c2_knn_pyramid_all_rows_vs_oracle():
    b = synth.make_batch(1, 16384, 8, "kitti", config=2, first_pair=9)
    for key in ("points_src", "points_ref"):
        o = O.nn_search_c(b[key], 16, (4, 4, 4, 4))
        for algo in (D.KNN_AUTO, D.KNN_TREE):
            g = D.nn_search_cloud(cu(b[key]), 16, (4, 4, 4, 4), algo=algo)
            assert g["neigh_idx"].shape == (1, 21760, 16) and g["interp_idx"].shape == (1, 21760, 1)
            for name in ("xyz", "neigh_idx", "sub_idx", "interp_idx"):
                assert torch.equal(g[name].cpu(), o[name]), (key, name, algo)


@pytest.mark.parametrize("shape,cloud", [((16384, 16384, 16), "kitti"), ((5000, 5000, 8), "3dmatch"), ((4096, 16384, 1), "kitti")])
@pytest.mark.parametrize("algo", [D.KNN_GRID, D.KNN_TREE, D.KNN_BRUTE])
def test_knn_vs_independent_kdtree(shape, cloud, algo):
    """Independent pin: scipy's cKDTree is exact in fp64 on the fp32 coordinates.  A query whose fp64 k-th and (k+1)-th
    distances differ by more than fp32 round-off has ONE correct index set; the GPU result must be that set, in the same
    order wherever consecutive fp64 distances are separated as well.  Queries below the gap are 'tie-ambiguous'
    (the fp32 distance, which this project defines, decides them) and are only counted."""
    from scipy.spatial import cKDTree
    ns, nq, k = shape
    g = torch.Generator().manual_seed(31)
    if cloud == "kitti":
        sup, qry = synth.kitti_cloud(ns, g)[:, :3].contiguous(), synth.kitti_cloud(nq, g)[:, :3].contiguous()
    else:
        c = synth.make_batch(1, ns, 8, "3dmatch", config=3, first_pair=2)
        sup, qry = c["points_src"][0, :, :3].contiguous(), c["points_ref"][0, :nq, :3].contiguous()
    i_g, d_g = D.knn(cu(sup[None]), cu(qry[None]), k, algo=algo)
    i_g, d_g = i_g[0].cpu().numpy(), d_g[0].cpu().numpy()
    dd, ii = cKDTree(sup.numpy().astype(np.float64)).query(qry.numpy().astype(np.float64), k=k + 1, workers=-1)
    d2 = dd ** 2
    scale = np.maximum(d2[:, k - 1], 1e-12)
    clear_set = (d2[:, k] - d2[:, k - 1]) > 4e-6 * scale + 1e-9        # the k-th neighbour is unambiguous
    assert clear_set.mean() > 0.99
    got_sets = np.sort(i_g, axis=1)
    ref_sets = np.sort(ii[:, :k], axis=1)
    assert (got_sets[clear_set] == ref_sets[clear_set]).all(), int((got_sets[clear_set] != ref_sets[clear_set]).any(1).sum())
    if k > 1:
        gaps = np.diff(d2[:, :k + 1], axis=1)
        clear_order = (gaps > 4e-6 * scale[:, None] + 1e-9).all(1)
        assert (i_g[clear_order] == ii[clear_order, :k]).all()
    assert np.allclose(d_g, d2[:, :k], rtol=2e-6, atol=1e-9)


def test_c3_soft_vs_oracle():
    c = synth.make_batch(1, 5000, 32, "3dmatch", config=3, first_pair=11)
    ref = c["points_ref"][:, :, :3].contiguous()
    fs, fr = c["feat_src"], c["feat_ref"]
    beta = torch.tensor([10.0])
    w, y, s, lse = O.soft_correspondence(fs, fr, ref, beta, 0.5)
    y_g, s_g, lse_g = D.match_soft(cu(fs), cu(fr), cu(ref), cu(beta), 0.5)
    assert torch.allclose(lse_g.cpu(), lse, rtol=1e-4, atol=1e-5)
    assert torch.allclose(s_g.cpu(), s, rtol=1e-4, atol=1e-6)
    assert torch.allclose(y_g.cpu(), y, rtol=1e-4, atol=1e-4)
    T_g, _ = D.kabsch_soft(cu(c["points_src"][:, :, :3].contiguous()), y_g, s_g)
    T_o, _ = O.compute_rigid_transform(c["points_src"][:, :, :3].contiguous(), ref, w)
    assert O.rotation_angle_deg(T_g.cpu()[:, :, :3], T_o[:, :, :3]).max() < 1e-3
    assert (T_g.cpu()[:, :, 3] - T_o[:, :, 3]).abs().max() < 1e-4


def test_c5_ten_iterations_vs_oracle():
    b = synth.make_batch(1, 20000, 64, "oxford", config=5, first_pair=3)
    xs = b["points_src"][:, :, :3].permute(0, 2, 1).contiguous()
    xr = b["points_ref"][:, :, :3].permute(0, 2, 1).contiguous()
    tr, pred, xyz, st = D.align_loop(cu(b["feat_src"]), cu(b["feat_ref"]), cu(xs), cu(xr), cu(b["weights"]), 10)
    tro, predo, xyzo = O.align_loop(b["feat_src"], b["feat_ref"], xs, xr, b["weights"], 10)
    assert (st == 0).all()
    assert torch.equal(torch.stack(pred).cpu(), torch.stack(predo))
    for a, c in zip(tr, tro):
        assert O.rotation_angle_deg(a.cpu()[:, :, :3], c[:, :, :3]).max() < 1e-3
        assert (a.cpu()[:, :, 3] - c[:, :, 3]).abs().max() < 1e-4
    assert torch.allclose(xyz.cpu(), xyzo, atol=2e-4)


def test_svdhead_rule_coincides_with_kabsch_at_unit_weights():
    """SURVEY a-16: SVDHead (network/matchnet.py:474-492, dead code in the reference: NameError on `math`) solves the
    unweighted problem with R = V diag(1,1,det(V U^T)) U^T.  That rule and compute_rigid_transform_2's "flip V[:, 2]"
    (model.py:49-54) give the same rotation; dsir_kabsch at w = 1 must reproduce it, reflection case included."""
    g = torch.Generator().manual_seed(4)
    src = torch.randn(3, 500, 3, generator=g, dtype=torch.float64)
    Rg = torch.linalg.qr(torch.randn(3, 3, 3, generator=g, dtype=torch.float64))[0]
    Rg = Rg * torch.linalg.det(Rg)[:, None, None].sign()
    tgt = src @ Rg.transpose(1, 2) + torch.randn(3, 1, 3, generator=g, dtype=torch.float64)
    tgt[2] = tgt[2] * torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64)      # a reflected target: det(V U^T) < 0
    tgt = tgt + 0.01 * torch.randn(3, 500, 3, generator=g, dtype=torch.float64)
    sc, tc = src - src.mean(1, keepdim=True), tgt - tgt.mean(1, keepdim=True)
    H = sc.transpose(1, 2) @ tc                                               # matchnet.py:476
    U, S, Vh = torch.linalg.svd(H)
    V = Vh.transpose(1, 2)
    det = torch.linalg.det(V @ U.transpose(1, 2))
    Dm = torch.diag_embed(torch.stack([torch.ones_like(det), torch.ones_like(det), det], 1))
    R_head = V @ Dm @ U.transpose(1, 2)                                       # matchnet.py:484-487
    t_head = tgt.mean(1) - (R_head @ src.mean(1)[:, :, None])[:, :, 0]
    T, flag = D.compute_rigid_transform_2(cu(src.float()), cu(tgt.float()), cu(torch.ones(3, 500, 1)))
    assert O.rotation_angle_deg(T.cpu()[:, :, :3], R_head.float()).max() < 1e-3
    assert (T.cpu()[:, :, 3] - t_head.float()).abs().max() < 1e-4
