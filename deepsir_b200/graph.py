"""KNN consumers of the RandLA-Net local aggregation and Sinkhorn, mirroring network/tools.py, network/RandLANet.py and
network/matchnet.py:211-271 of the reference (same names, shapes and argument meaning), executed by libdeepsir_b200.so."""
from __future__ import annotations

import torch

from . import _lib as L


def _idx64(idx):
    return idx if (idx.dtype == torch.int64 and idx.is_contiguous()) else idx.to(torch.int64).contiguous()


def gather_neighbour_V2(inputs, neigh_idx):
    """network/tools.py:197-209.  inputs [B,C,N], neigh_idx [B,M,k] -> [B,C,M,k]."""
    dev = L.require_cuda(inputs, neigh_idx)
    B, C, N = inputs.shape
    _, M, k = neigh_idx.shape
    inputs, neigh_idx = inputs.contiguous(), _idx64(neigh_idx)
    out = torch.empty(B, C, M, k, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_gather_neighbours(inputs.data_ptr(), B, C, N, neigh_idx.data_ptr(), M, k, out.data_ptr(),
                                           L.stream_ptr(dev)), "dsir_gather_neighbours")
    return out


def gather_neighbour(inputs, neigh_idx):
    """network/tools.py:183-195.  inputs [B,N,C], neigh_idx [B,N,k] -> [B,N,k,C] (layout views around the same kernel)."""
    return gather_neighbour_V2(inputs.permute(0, 2, 1), neigh_idx).permute(0, 2, 3, 1)


def gather_neighbour_V4(inputs, neigh_idx):
    """network/tools.py:223-233.  inputs [B,N,C], neigh_idx [B,M] -> [B,M,C]."""
    from .match import gather_neighbour_V3
    return gather_neighbour_V3(inputs.permute(0, 2, 1), neigh_idx).permute(0, 2, 1)


def relative_pos_encoding(xyz, neigh_idx):
    """Building_block.relative_pos_encoding (network/RandLANet.py:197-212).  xyz [B,3,N], neigh_idx [B,N,k] -> [B,10,N,k]."""
    dev = L.require_cuda(xyz, neigh_idx)
    B, _, N = xyz.shape
    k = neigh_idx.shape[-1]
    xyz, neigh_idx = xyz.contiguous(), _idx64(neigh_idx)
    out = torch.empty(B, 10, N, k, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_rel_pos_encoding(xyz.data_ptr(), B, N, neigh_idx.data_ptr(), k, out.data_ptr(), L.stream_ptr(dev)),
            "dsir_rel_pos_encoding")
    return out


def random_sample(feature, pool_idx):
    """RandLA.random_sample (network/RandLANet.py:374-391).  feature [B,C,N,1], pool_idx [B,M,k] -> [B,C,M,1]
    (max over the k neighbours without materialising [B,C,M,k])."""
    dev = L.require_cuda(feature, pool_idx)
    f = feature.squeeze(3).contiguous()
    B, C, N = f.shape
    _, M, k = pool_idx.shape
    pool_idx = _idx64(pool_idx)
    out = torch.empty(B, C, M, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_pool_max(f.data_ptr(), B, C, N, pool_idx.data_ptr(), M, k, out.data_ptr(), L.stream_ptr(dev)),
            "dsir_pool_max")
    return out.unsqueeze(3)


def nearest_interpolation(feature, interp_idx):
    """RandLA.nearest_interpolation (network/RandLANet.py:393-408).  feature [B,C,N,1], interp_idx [B,M,1] -> [B,C,M,1]."""
    from .match import gather_neighbour_V3
    return gather_neighbour_V3(feature.squeeze(3), interp_idx.reshape(interp_idx.shape[0], -1)).unsqueeze(3)


def sinkhorn(log_alpha, n_iters=5, slack=True, eps=-1):
    """network/matchnet.py:211-271.  log_alpha [B,J,K] -> log of the (near) doubly stochastic matrix [B,J,K].
    eps > 0 (early termination, handcrafted RPM only) runs one iteration per call and checks the reference's criterion."""
    dev = L.require_cuda(log_alpha)
    B, J, K = log_alpha.shape
    a = log_alpha.contiguous()
    lib = L.lib()
    ws = L.workspace(lib.dsir_sinkhorn_workspace_bytes(B, J, K), dev)

    def run(n):
        out = torch.empty(B, J, K, dtype=torch.float32, device=dev)
        L.check(lib.dsir_sinkhorn(a.data_ptr(), B, J, K, n, int(bool(slack)), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                  L.stream_ptr(dev)), "dsir_sinkhorn")
        return out

    if eps <= 0:
        return run(n_iters)
    prev, out = None, a
    for i in range(n_iters):      # matchnet.py:246-251 / :262-267
        out = run(i + 1)
        cur = torch.exp(out)
        if prev is not None and torch.max(torch.sum(torch.abs(cur - prev), dim=[1, 2])) < eps:
            break
        prev = cur
    return out


def log_optimal_transport(scores, alpha, iters: int):
    """network/matchnet.py:836-856 (with log_sinkhorn_iterations :827-833).  scores [B,M,N], alpha: the learned dustbin
    score (0-d / 1-element tensor or float) -> [B,M+1,N+1] log couplings (times M+N).  The augmented matrix with the
    dustbin row and column is never built."""
    dev = L.require_cuda(scores)
    B, M, N = scores.shape
    s = scores.contiguous()
    if s.dtype != torch.float32:
        raise L.DeepSIRError("log_optimal_transport expects float32 scores")
    a = alpha.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous() if isinstance(alpha, torch.Tensor) \
        else torch.full((1,), float(alpha), dtype=torch.float32, device=dev)
    out = torch.empty(B, M + 1, N + 1, dtype=torch.float32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_log_optimal_transport_workspace_bytes(B, M, N), dev)
    L.check(lib.dsir_log_optimal_transport(s.data_ptr(), B, M, N, a.data_ptr(), int(iters), out.data_ptr(), ws.data_ptr(),
                                           ws.numel(), L.stream_ptr(dev)), "dsir_log_optimal_transport")
    return out
