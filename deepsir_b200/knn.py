"""xyz KNN with the contract of torch_points_kernels.knn as used by DataBase.nn_search
(dataloader/data_base.py:153-183), on the device."""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L


def knn(pos_support, pos, k, algo=L.KNN_AUTO):
    """knn(support [B,Ns,3+], query [B,Nq,3+], k) -> (idx int64 [B,Nq,k], dist2 fp32 [B,Nq,k]), ascending in
    (dist2, index).  Raises when Ns < k like the reference kernel."""
    dev = L.require_cuda(pos_support, pos)
    if pos_support.dtype != torch.float32 or pos.dtype != torch.float32:
        raise L.DeepSIRError("knn expects float32 points")
    s = pos_support if pos_support.is_contiguous() else pos_support.contiguous()
    q = pos if pos.is_contiguous() else pos.contiguous()
    B, Ns, ss = s.shape
    Nq, qs = q.shape[1], q.shape[2]
    idx = torch.empty(B, Nq, k, dtype=torch.int64, device=dev)
    d2 = torch.empty(B, Nq, k, dtype=torch.float32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_knn_workspace_bytes(B, Ns, Nq, k, algo), dev)
    L.check(lib.dsir_knn_xyz(s.data_ptr(), ss, q.data_ptr(), qs, B, Ns, Nq, k, idx.data_ptr(), d2.data_ptr(),
                             ws.data_ptr(), ws.numel(), algo, L.stream_ptr(dev)), "dsir_knn_xyz")
    return idx, d2


def nn_search_cloud(points, num_knn=16, sub_sampling_ratio=(4, 4, 4, 4), algo=L.KNN_AUTO, _stream=None, _keep=None):
    """One cloud tensor [B,N,C>=3] -> dict(xyz, neigh_idx, sub_idx, interp_idx), the four tensors
    DataBase.nn_search (data_base.py:179-182) attaches per cloud.  One library call for the whole pyramid."""
    dev = L.require_cuda(points)
    if points.dtype != torch.float32:
        raise L.DeepSIRError("nn_search expects float32 points")
    p = _rows_view(points)
    B, N = p.shape[0], p.shape[1]
    S = p.stride(1) if N > 1 else p.shape[2]
    ratios = [int(r) for r in sub_sampling_ratio]
    Lv = len(ratios)
    sumN, sumSub, n = 0, 0, N
    for r in ratios:
        sumN += n
        sumSub += n // r
        n //= r
    xyz = torch.empty(B, sumN, 3, dtype=torch.float32, device=dev)
    neigh = torch.empty(B, sumN, num_knn, dtype=torch.int64, device=dev)
    sub = torch.empty(B, sumSub, num_knn, dtype=torch.int64, device=dev)
    interp = torch.empty(B, sumN, 1, dtype=torch.int64, device=dev)
    rat = (ctypes.c_int * Lv)(*ratios)
    lib = L.lib()
    ws = L.workspace(lib.dsir_knn_pyramid_workspace_bytes(B, N, num_knn, ctypes.addressof(rat), Lv, algo), dev)
    L.check(lib.dsir_knn_pyramid(p.data_ptr(), S, B, N, ctypes.addressof(rat), Lv, num_knn, xyz.data_ptr(),
                                 neigh.data_ptr(), sub.data_ptr(), interp.data_ptr(), ws.data_ptr(), ws.numel(), algo,
                                 L.stream_ptr(dev) if _stream is None else _stream.cuda_stream), "dsir_knn_pyramid")
    if _keep is not None:
        _keep += [ws, p]   # the caller frees them after it has joined `_stream`
    return dict(xyz=xyz, neigh_idx=neigh, sub_idx=sub, interp_idx=interp)


def _rows_view(points):
    """[B,N,C>=3] tensor whose rows the library can address with one point stride: contiguous, or an inner slice such as
    data['points_src'][:, :, :3] (data_base.py:159) - passed by stride, not copied; anything else is made contiguous."""
    B, N, C = points.shape
    if points.stride(2) == 1 and (N <= 1 or (points.stride(1) >= C and (B <= 1 or points.stride(0) == N * points.stride(1)))):
        return points
    return points.contiguous()


_side_streams = {}


def nn_search_pair(points_src, points_ref, num_knn=16, sub_sampling_ratio=(4, 4, 4, 4), algo=L.KNN_AUTO):
    """Both pyramids of a pair batch at once: the source pyramid on the caller's stream, the reference pyramid on a side
    stream that forks from and joins back into it.  The two are independent (data_base.py:157 loops over the two keys), and
    the small launches of the coarse levels of one fill the tail of the other.  Returns (graph_src, graph_ref)."""
    dev = L.require_cuda(points_src, points_ref)
    cur = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(dev)
    # Outputs and workspace of BOTH pyramids are allocated on the caller's stream (no cross-stream ownership for the
    # caching allocator to track); only the kernels of the reference pyramid run on the side stream, between a fork
    # (side waits for everything enqueued so far) and a join (the caller's stream waits for the side stream).
    keep = []
    # a copy made here (non-viewable input) is enqueued on the caller's stream BEFORE the fork, so the side stream sees it
    points_src, points_ref = _rows_view(points_src), _rows_view(points_ref)
    side.wait_stream(cur)                                                       # fork
    g_ref = nn_search_cloud(points_ref, num_knn, sub_sampling_ratio, algo, _stream=side, _keep=keep)
    g_src = nn_search_cloud(points_src, num_knn, sub_sampling_ratio, algo)
    cur.wait_stream(side)                                                       # join
    del keep   # workspace of the side pyramid: released only now, so its next user on `cur` is ordered after the join
    return g_src, g_ref


def nn_search(data_list_stack, num_knn=16, sub_sampling_ratio=(4, 4, 4, 4), algo=L.KNN_AUTO):
    """DataBase.nn_search (data_base.py:153-183) on a dict already moved to the device: adds
    '<k>_xyz', '<k>_neigh_idx', '<k>_sub_idx', '<k>_interp_idx' for k in points_src, points_ref."""
    pair = nn_search_pair(data_list_stack["points_src"], data_list_stack["points_ref"], num_knn, sub_sampling_ratio, algo)
    for k, r in zip(["points_src", "points_ref"], pair):
        data_list_stack[k + "_xyz"] = r["xyz"]
        data_list_stack[k + "_neigh_idx"] = r["neigh_idx"]
        data_list_stack[k + "_sub_idx"] = r["sub_idx"]
        data_list_stack[k + "_interp_idx"] = r["interp_idx"]
    return data_list_stack
