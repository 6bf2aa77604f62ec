"""The drop-in, executed: the reference's own `Network` (unmodified copy under baseline/_ref, shipped by build()) runs
`forward_align_4` (network/model.py:520-607, called from test.py:400) stock on the GPU and again with deepsir_b200
patched in; same seeded random-init weights (the shipped checkpoint is a missing blob), same inputs.

Bars: `endpoints` keys / types / dtypes / devices identical; `pred_pairs` equal on every row whose fp64 top-2 gap is
above fp32 round-off at iteration 0 (later iterations see poses that agree only to tolerance, so rows near a tie may
flip: they are counted and bounded).  Poses: a random-init network matches at random, so the weighted Kabsch problem it
poses is ill-conditioned and the reference's own fp32 centroid / covariance sums (model.py:38-45) are off by ~1e-2 deg
there.  The bar is therefore stated against the exact answer: on the SAME inputs (recorded inside the patched run) the
library must be within 1e-3 deg / 1e-4 m of the fp64 solution or at least as close to it as the reference's function is;
the end-to-end transforms of the two arms must agree to the reference's own error level.
"""
import os
import sys

import pytest
import torch

import deepsir_b200 as D
from deepsir_b200 import patch as P
from deepsir_b200 import synth
from oracle import deepsir_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
DEV = "cuda:0"

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "network")),
                               reason="baseline/_ref absent (made by __graft_entry__.build() where /root/reference exists)")


def _reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import arguments
    from network import model as M
    args = arguments.eval_arguments().parse_args([])
    return M, args


def _net(M, args, seed=0):
    torch.manual_seed(seed)
    net = M.Network(args).eval().to(DEV)
    net.label_weights = net.label_weights.to(DEV)   # a plain attribute in the reference (model.py:150), indexed at :742
    return net


def _kabsch64(src, tgt, w):
    """network/model.py:22-58 evaluated entirely in fp64 on the host: the exact answer both implementations approximate."""
    src, tgt, w = src.double().cpu(), tgt.double().cpu(), w.double().cpu()
    wn = w / (w.abs().sum(dim=1, keepdim=True) + 1e-16)
    cs, ct = (src * wn).sum(1), (tgt * wn).sum(1)
    cov = (src - cs[:, None, :]).transpose(1, 2) @ ((tgt - ct[:, None, :]) * wn)
    u, s, vh = torch.linalg.svd(cov)
    v = vh.transpose(1, 2).clone()
    flip = torch.linalg.det(v @ u.transpose(1, 2)) <= 0
    v[flip, :, 2] *= -1
    R = v @ u.transpose(1, 2)
    t = ct - (R @ cs[:, :, None])[:, :, 0]
    return torch.cat([R, t[:, :, None]], 2).float()


def _data(n, B=1, first_pair=0):
    b = synth.make_batch(B, n, 64, "kitti", config=1, first_pair=first_pair)
    data = {"points_src": b["points_src"].to(DEV), "points_ref": b["points_ref"].to(DEV)}
    return D.nn_search(data)      # the reference's loader-side KNN (torch_points_kernels) is absent: both arms share this graph


def _run_arms(M, net, data, iters, level):
    """(stock transforms, endpoints), (patched transforms, endpoints), recorded Kabsch calls of the patched arm."""
    rec = []
    real = P._K.compute_rigid_transform_2

    def recording(src, tgt, weights):
        T, flag = real(src, tgt, weights)
        rec.append((src.clone(), tgt.clone(), weights.clone(), T.clone()))
        return T, flag
    P.unpatch()
    tr0, ep0 = net(dict(data), (iters, False))
    P._K.compute_rigid_transform_2 = recording
    P.patch(level=level)
    try:
        tr1, ep1 = net(dict(data), (iters, False))
    finally:
        P.unpatch()
        P._K.compute_rigid_transform_2 = real
    return (tr0, ep0), (tr1, ep1), rec


@needs_ref
@pytest.mark.parametrize("n,iters", [(4096, 5), (2048, 2)])
def test_forward_align_4_stock_vs_patched(n, iters):
    M, args = _reference()
    net = _net(M, args)
    strict_iterations = 0
    with torch.no_grad():
        for first_pair, level in ((0, "loop"), (0, "leaf"), (1, "loop"), (2, "loop"), (3, "loop")):
            data = _data(n, first_pair=first_pair)
            (tr0, ep0), (tr1, ep1), rec = _run_arms(M, net, data, iters, level)
            # ---- structure of the outputs (model.py:520-607)
            assert len(tr1) == len(tr0) == iters and len(rec) == iters
            assert set(ep1) == set(ep0)
            for k in ep0:
                a, b = ep0[k], ep1[k]
                assert type(a) is type(b), k
                if isinstance(a, list):
                    assert len(a) == len(b) == iters, k
                    a, b = a[0], b[0]
                if isinstance(a, torch.Tensor):
                    assert a.shape == b.shape and a.dtype == b.dtype and a.device == b.device, k
            assert ep1["pred_pairs"][0].dtype == torch.int32 and ep1["pred_pairs"][0].device.type == "cpu"
            assert ep1["invalid_gradient"] == ep0["invalid_gradient"]
            assert torch.equal(ep0["pred_pairs"][0][..., 0], ep1["pred_pairs"][0][..., 0])
            # ---- iteration 0: identical inputs -> identical correspondences wherever the fp64 gap is above fp32 round-off
            f0, x0, l0, s0, f1, x1, l1, s1 = net.forward_pair(dict(data))
            fs, fr = net.aggregation(x0, x1, f0, f1, l0, l1, s0, s1)
            _, gap = O.match_top2_fp64(fs.cpu(), fr.cpu())
            clear = gap[0] > 2e-6
            assert clear.float().mean() > 0.99
            p0, p1 = ep0["pred_pairs"][0][0, :, 1], ep1["pred_pairs"][0][0, :, 1]
            assert torch.equal(p0[clear], p1[clear]), (level, int((p0 != p1)[clear].sum()))
            # ---- every Kabsch call of the patched arm, on its own recorded inputs: library vs the reference's function vs fp64
            for src, tgt, w, T_lib in rec:
                T_ref, _ = M.compute_rigid_transform_2(src, tgt, w)                       # the reference's own, stock on the GPU
                T_64 = _kabsch64(src, tgt, w)
                a_lib = O.rotation_angle_deg(T_lib.cpu()[:, :, :3], T_64[:, :, :3]).max().item()
                a_ref = O.rotation_angle_deg(T_ref.cpu()[:, :, :3], T_64[:, :, :3]).max().item()
                t_lib = (T_lib.cpu()[:, :, 3] - T_64[:, :, 3]).norm(dim=1).max().item()
                t_ref = (T_ref.cpu()[:, :, 3] - T_64[:, :, 3]).norm(dim=1).max().item()
                assert a_lib <= max(1e-3, a_ref) and t_lib <= max(1e-4, t_ref), (level, a_lib, a_ref, t_lib, t_ref)
            # ---- end to end.  A random-init network matches at random (weights ~0.47 everywhere), so ONE tie-ambiguous row
            # that the two arms resolve differently (cuBLAS sgemm order vs the library's fma chain; 0.2 % of the rows are
            # ambiguous) moves the pose by up to 1e-2 deg.  The strict bar applies to the iterations up to the first
            # such flip; after it the arms are only required to stay close (and the per-call check above still holds).
            same = True
            for it in range(iters):
                flips = int((ep0["pred_pairs"][it][..., 1] != ep1["pred_pairs"][it][..., 1]).sum())
                same = same and flips == 0
                assert flips < 5e-3 * n, (level, it, flips)
                ang = O.rotation_angle_deg(tr1[it][:, :, :3].cpu(), tr0[it][:, :, :3].cpu()).max().item()
                dt = (tr1[it][:, :, 3] - tr0[it][:, :, 3]).norm(dim=1).max().item()
                if same:
                    assert ang < 1e-3 and dt < 1e-4, (level, first_pair, it, ang, dt)
                    # (the inlier network amplifies the 1e-4 m pose agreement of the previous iteration)
                    assert torch.allclose(ep1["perm_matrices"][it], ep0["perm_matrices"][it], atol=1e-4 if it == 0 else 5e-3), (level, it)
                    strict_iterations += 1
                else:
                    assert ang < 0.5 and dt < 0.1, (level, first_pair, it, ang, dt)
            if same:
                assert torch.allclose(ep1["pt_ref_new"], ep0["pt_ref_new"], atol=1e-4)
    assert strict_iterations >= 1, "no input without a tie-ambiguous flip at iteration 0: the strict bar was never exercised"


@needs_ref
def test_patch_is_reversible_and_cpu_inputs_raise():
    M, args = _reference()
    orig = (M.match_features_V2, M.compute_rigid_transform_2, M.gather_neighbour_V3, M.se3_torch, M.Network.forward_align_4)
    P.patch()
    assert M.match_features_V2 is D.match_features_V2 and M.se3_torch is D.se3_torch
    assert M.Network.forward_align_4 is P.forward_align_4
    with pytest.raises(D.DeepSIRError):
        M.match_features_V2(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8))      # no CPU fallback behind the patched name
    P.unpatch()
    assert (M.match_features_V2, M.compute_rigid_transform_2, M.gather_neighbour_V3, M.se3_torch,
            M.Network.forward_align_4) == orig


def test_patch_knn_binds_the_loader_namespace():
    import types
    fake = types.ModuleType("data_base")
    fake.Util = types.SimpleNamespace(knn=lambda *a: None, ball_query=None)
    keep = fake.Util.knn
    P.patch_knn(fake)
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(2, 700, 3, generator=g)
    idx, d2 = fake.Util.knn(pts.to(DEV), pts.to(DEV), 16)                     # data_base.py:165 call shape
    io, do = O.knn(pts, pts, 16)
    assert torch.equal(idx.cpu(), io) and torch.equal(d2.cpu(), do)
    P.unpatch()
    assert fake.Util.knn is keep
    bare = types.ModuleType("data_base")                                      # torch_points_kernels absent: namespace installed
    P.patch_knn(bare)
    assert bare.Util.knn(pts.to(DEV), pts.to(DEV), 1)[0].shape == (2, 700, 1)
