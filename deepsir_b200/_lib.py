"""ctypes binding of libdeepsir_b200.so (include/deepsir_b200.h).

There is exactly one code path: the sm_100a CUDA library.  If the shared object is missing, or a
call is made with CPU tensors, or the device is not a B200-class (sm_100) GPU, the wrappers raise —
they never fall back to PyTorch ops or to the CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSIR_B200_LIB") or os.path.join(_PKG, "libdeepsir_b200.so")   # override: development builds only
CSRC = os.path.join(_PKG, "csrc")
SOURCES = ["api.cu", "knn.cu", "knn_grid.cu", "knn_tree.cu", "match_fp32.cu", "match_tc.cu", "match_tc_soft.cu", "kabsch.cu", "graph.cu", "keypoint.cu", "metrics.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

OK = 0
KNN_AUTO, KNN_BRUTE, KNN_GRID, KNN_TREE = 0, 1, 2, 3
MATCH_AUTO, MATCH_FP32, MATCH_TC = 0, 1, 2
METRIC_L2, METRIC_EUCLIDEAN, METRIC_ACOS_DOT, METRIC_SQDIFF, METRIC_CITYBLOCK, METRIC_SQDIFF_SQRT = range(6)


class DeepSIRError(RuntimeError):
    pass


class Feat(ctypes.Structure):  # dsir_feat
    _fields_ = [("ptr", ctypes.c_void_p), ("batch_stride", ctypes.c_int64), ("chan_stride", ctypes.c_int64),
                ("point_stride", ctypes.c_int64)]


class Points(ctypes.Structure):  # dsir_points
    _fields_ = [("ptr", ctypes.c_void_p), ("batch_stride", ctypes.c_int64), ("point_stride", ctypes.c_int64),
                ("coord_stride", ctypes.c_int64)]


def build(verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into the in-tree shared object (nvcc cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(os.path.dirname(_PKG), "include", "deepsir_b200.h")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + srcs + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None

_c = ctypes
_SIGS = {
    "dsir_version": (_c.c_int, []),
    "dsir_strerror": (_c.c_char_p, [_c.c_int]),
    "dsir_last_cuda_error": (_c.c_char_p, []),
    "dsir_device_check": (_c.c_int, []),
    "dsir_launch_count": (_c.c_uint64, []),
    "dsir_profile_begin": (_c.c_int, [_c.c_void_p]),
    "dsir_profile_report": (_c.c_int, [_c.c_void_p, _c.c_size_t]),
    "dsir_knn_workspace_bytes": (_c.c_size_t, [_c.c_int] * 5),
    "dsir_knn_xyz": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "dsir_knn_pyramid_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int]),
    "dsir_knn_pyramid": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int,
                                    _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t,
                                    _c.c_int, _c.c_void_p]),
    "dsir_match_dense_workspace_bytes": (_c.c_size_t, [_c.c_int] * 3),
    "dsir_match_dense": (_c.c_int, [Feat, Feat, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                    _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_match_argmin_workspace_bytes": (_c.c_size_t, [_c.c_int] * 5),
    "dsir_match_argmin": (_c.c_int, [Feat, Feat, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                     _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "dsir_match_argmin_hint": (_c.c_int, [Feat, Feat, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                          _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_void_p]),
    "dsir_match_argmin_rescued_rows": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                                  _c.c_void_p, _c.c_void_p]),
    "dsir_match_argmin_filter_timing": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                                   _c.c_void_p, _c.c_void_p]),
    "dsir_match_argmin_filter_trace": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                                  _c.c_void_p, _c.c_void_p]),
    "dsir_match_soft_workspace_bytes": (_c.c_size_t, [_c.c_int] * 4),
    "dsir_match_soft_sweep": (_c.c_int, [Feat, Feat, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                         _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_match_soft_topk_workspace_bytes": (_c.c_size_t, [_c.c_int] * 5),
    "dsir_match_soft_topk_fused": (_c.c_int, [_c.c_int] * 5),
    "dsir_match_soft_topk_exhaustive_rows": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                                        _c.c_void_p, _c.c_void_p]),
    "dsir_match_soft": (_c.c_int, [Feat, Feat, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                   _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p,
                                   _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_gather_points": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p,
                                      _c.c_void_p]),
    "dsir_gather_neighbours": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int,
                                          _c.c_void_p, _c.c_void_p]),
    "dsir_rel_pos_encoding": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "dsir_pool_max": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                 _c.c_void_p]),
    "dsir_sinkhorn_workspace_bytes": (_c.c_size_t, [_c.c_int] * 3),
    "dsir_sinkhorn": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                 _c.c_size_t, _c.c_void_p]),
    "dsir_log_optimal_transport_workspace_bytes": (_c.c_size_t, [_c.c_int] * 3),
    "dsir_log_optimal_transport": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p,
                                              _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_kabsch_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int]),
    "dsir_kabsch": (_c.c_int, [Points, Points, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                               _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_kabsch_moments": (_c.c_int, [Points, Points, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_kabsch_from_moments": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "dsir_soft_targets": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                     _c.c_void_p, _c.c_void_p]),
    "dsir_kabsch_soft": (_c.c_int, [Points, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                    _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_se3_apply": (_c.c_int, [_c.c_void_p, _c.c_int64, Points, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int64,
                                  _c.c_int64, _c.c_int64, _c.c_int, _c.c_void_p]),
    "dsir_se3_compose": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p,
                                    _c.c_void_p]),
    "dsir_se3_inverse": (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "dsir_align_loop_workspace_bytes": (_c.c_size_t, [_c.c_int] * 5),
    "dsir_align_loop": (_c.c_int, [Feat, Feat, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                   _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                   _c.c_size_t, _c.c_int, _c.c_void_p]),
    "dsir_keypoint_score_workspace_bytes": (_c.c_size_t, [_c.c_int] * 3),
    "dsir_keypoint_score": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p,
                                       _c.c_int, _c.c_int, _c.c_float, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                       _c.c_size_t, _c.c_void_p]),
    "dsir_topk_rows": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "dsir_pose_errors": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p,
                                    _c.c_void_p]),
    "dsir_correspondence_check_workspace_bytes": (_c.c_size_t, [_c.c_int64]),
    "dsir_correspondence_check": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                             _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "dsir_nn_sqdist_workspace_bytes": (_c.c_size_t, [_c.c_int]),
    "dsir_nn_sqdist_mean": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                       _c.c_void_p, _c.c_size_t, _c.c_void_p]),
}
EXPORTS = tuple(_SIGS)


def lib() -> ctypes.CDLL:
    """Load the C-ABI library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DeepSIRError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(deepsir_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        L = lib()
        msg = L.dsir_strerror(rc).decode()
        if rc == -4:
            msg += ": " + L.dsir_last_cuda_error().decode()
        raise DeepSIRError(f"{what}: {msg} (code {rc})")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DeepSIRError("deepsir_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise DeepSIRError("all tensors must live on the same CUDA device")
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        # the library launches on the process's CURRENT device (one process per GPU): make it explicit instead of failing
        # later with an invalid-resource-handle
        raise DeepSIRError(f"tensors live on cuda:{dev.index} but the current device is cuda:{torch.cuda.current_device()}: "
                           "call torch.cuda.set_device() (one process per GPU)")
    return dev


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def feat_cn(t: torch.Tensor) -> Feat:
    """[B,C,N] view (any batch/channel stride, unit point stride) -> dsir_feat."""
    if t.dtype != torch.float32:
        raise DeepSIRError("features must be float32")
    if t.stride(2) != 1 and t.shape[2] > 1:
        t = t.contiguous()
    return Feat(t.data_ptr(), t.stride(0), t.stride(1), 1), t


def feat_nc(t: torch.Tensor) -> Feat:
    """[B,N,C] view -> dsir_feat (channel index is the last axis)."""
    if t.dtype != torch.float32:
        raise DeepSIRError("features must be float32")
    if t.stride(2) != 1 and t.shape[2] > 1:
        t = t.contiguous()
    return Feat(t.data_ptr(), t.stride(0), 1, t.stride(1)), t


def points_bm3(t: torch.Tensor) -> Points:
    """[B,M,3] (any strides) -> dsir_points."""
    return Points(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def points_b3m(t: torch.Tensor) -> Points:
    """[B,3,M] (any strides) -> dsir_points."""
    return Points(t.data_ptr(), t.stride(0), t.stride(2), t.stride(1))
