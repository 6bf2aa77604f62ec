"""Host-side logic of deepsir_b200.patch (no GPU): rebinding inside the reference's modules and restoring them."""
import os
import sys

import pytest
import torch

import deepsir_b200 as D
from deepsir_b200 import patch as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "network")), reason="baseline/_ref absent")
def test_patch_rebinds_and_restores():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from network import model as M
    orig = {k: getattr(M, k) for k in ("match_features_V2", "gather_neighbour_V3", "compute_rigid_transform_2",
                                       "compute_rigid_transform", "se3_torch")}
    fwd = M.Network.forward_align_4
    assert P.patch() is M
    assert M.match_features_V2 is D.match_features_V2 and M.gather_neighbour_V3 is D.gather_neighbour_V3
    assert M.compute_rigid_transform_2 is P._compute_rigid_transform_2 and M.se3_torch is D.se3_torch
    assert M.Network.forward_align_4 is P.forward_align_4
    P.patch(level="leaf")                       # idempotent; the leaf level keeps the reference's own loop
    assert M.Network.forward_align_4 is fwd and M.match_features_V2 is D.match_features_V2
    with pytest.raises(D.DeepSIRError):         # the rebound names have no CPU fallback
        M.match_features_V2(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8))
    P.unpatch()
    assert all(getattr(M, k) is v for k, v in orig.items()) and M.Network.forward_align_4 is fwd
    with pytest.raises(ValueError):
        P.patch(level="everything")
    P.unpatch()


def test_patch_knn_namespace():
    import types
    fake = types.ModuleType("data_base")
    fake.Util = types.SimpleNamespace(knn=len)
    P.patch_knn(fake)
    assert fake.Util.knn is D.knn
    P.unpatch()
    assert fake.Util.knn is len
    bare = types.ModuleType("data_base")
    P.patch_knn(bare)
    assert bare.Util.knn is D.knn
