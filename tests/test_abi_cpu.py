"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol the
header declares, and the host mirror refuses to run without CUDA (no fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

import deepsir_b200 as D
from deepsir_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return D.build()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "deepsir_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dsir_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/deepsir_b200.h but not exported"
    # and the Python binding table covers the header exactly
    assert sorted(_lib.EXPORTS) == names


def test_error_strings_and_version(built):
    L = D.lib()
    assert L.dsir_version() >= 100
    assert L.dsir_strerror(0) == b"ok"
    for code in range(-6, 0):
        assert L.dsir_strerror(code) not in (b"ok", b"unknown error")
    assert L.dsir_strerror(-99) == b"unknown error"


def test_workspace_queries_are_pure(built):
    L = D.lib()
    assert L.dsir_knn_workspace_bytes(2, 1024, 1024, 16, 0) >= 2 * 1024 * 16
    assert L.dsir_match_argmin_workspace_bytes(1, 64, 1000, 1000, 1) >= 8000
    assert L.dsir_kabsch_workspace_bytes(4, 16384) >= 4 * 17 * 8
    assert L.dsir_align_loop_workspace_bytes(2, 64, 512, 512, 0) > 0
    # top-k soft match: never less than the plain soft pass, and the route query is a pure function of the shape
    base = L.dsir_match_soft_workspace_bytes(4, 32, 5000, 5000)
    assert L.dsir_match_soft_topk_workspace_bytes(4, 32, 5000, 5000, 0) == base
    assert L.dsir_match_soft_topk_workspace_bytes(4, 32, 5000, 5000, 16) > base
    assert L.dsir_match_soft_topk_fused(4, 32, 5000, 5000, 16) in (0, 1)       # 1 only where a driver (TMA descriptors) exists
    assert L.dsir_match_soft_topk_fused(4, 32, 5000, 200, 32) == 0             # 7 granules of 32 columns < 1.25 x 32
    assert L.dsir_match_soft_topk_fused(4, 65, 5000, 5000, 16) == 0            # C > 64: fp32 CUDA-core route


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_no_device_is_reported(built):
    assert D.lib().dsir_device_check() == -6


def test_cpu_tensors_are_refused(built):
    f = torch.randn(1, 8, 16)
    with pytest.raises(D.DeepSIRError):
        D.match_argmin(f, f)
    with pytest.raises(D.DeepSIRError):
        D.knn(torch.randn(1, 32, 3), torch.randn(1, 8, 3), 4)
    with pytest.raises(D.DeepSIRError):
        D.compute_rigid_transform_2(torch.randn(1, 8, 3), torch.randn(1, 8, 3), torch.ones(1, 8, 1))
    with pytest.raises(D.DeepSIRError):
        D.se3_torch.concatenate(torch.eye(3, 4)[None], torch.eye(3, 4)[None])


def test_missing_library_fails_loudly(monkeypatch, built):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "deepsir_b200", "does_not_exist.so"))
    with pytest.raises(D.DeepSIRError):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deepsir_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "deepsir_oracle" not in txt, f
