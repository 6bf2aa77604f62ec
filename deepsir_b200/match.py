"""Feature distance + correspondence, mirroring network/matchnet.py and network/model.py:558-571 of the
reference (same function names, argument meaning and shapes), executed by libdeepsir_b200.so."""
from __future__ import annotations

import torch

from . import _lib as L

_EPS = 1e-16  # network/matchnet.py:_EPS


def _dense(fs, fr, B, C, J, K, metric, keep, dev):
    out = torch.empty(B, J, K, dtype=torch.float32, device=dev)
    lib = L.lib()
    ws = L.workspace(lib.dsir_match_dense_workspace_bytes(B, J, K), dev)
    L.check(lib.dsir_match_dense(fs, fr, B, C, J, K, metric, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                 L.stream_ptr(dev)), "dsir_match_dense")
    del keep
    return out


def square_distance_V2(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """network/matchnet.py:96-113.  src [B,C,N], dst [B,C,M] -> [B,N,M]."""
    dev = L.require_cuda(src, dst)
    (fs, a), (fr, b) = L.feat_cn(src), L.feat_cn(dst)
    return _dense(fs, fr, src.shape[0], src.shape[1], src.shape[2], dst.shape[2], L.METRIC_L2, (a, b), dev)


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """network/matchnet.py:49-66.  src [B,N,C], dst [B,M,C] -> [B,N,M]."""
    dev = L.require_cuda(src, dst)
    (fs, a), (fr, b) = L.feat_nc(src), L.feat_nc(dst)
    return _dense(fs, fr, src.shape[0], src.shape[2], src.shape[1], dst.shape[1], L.METRIC_L2, (a, b), dev)


def match_features_V2(feat_src, feat_ref, metric="l2"):
    """network/matchnet.py:116-144.  feat_src [B,C,J], feat_ref [B,C,K] -> [B,J,K]."""
    assert feat_src.shape[1] == feat_ref.shape[1]
    dev = L.require_cuda(feat_src, feat_ref)
    B, C, J = feat_src.shape
    K = feat_ref.shape[2]
    if metric == "l2":
        code = L.METRIC_L2
    elif metric == "euclidean":
        code = L.METRIC_EUCLIDEAN
    elif metric == "angle":
        feat_src = feat_src / (torch.norm(feat_src, dim=1, keepdim=True) + _EPS)
        feat_ref = feat_ref / (torch.norm(feat_ref, dim=1, keepdim=True) + _EPS)
        code = L.METRIC_ACOS_DOT
    else:
        raise NotImplementedError
    (fs, a), (fr, b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
    return _dense(fs, fr, B, C, J, K, code, (a, b), dev)


def match_features(feat_src, feat_ref, metric="l2"):
    """network/matchnet.py:69-93.  feat_src [B,J,C], feat_ref [B,K,C] -> [B,J,K]."""
    assert feat_src.shape[-1] == feat_ref.shape[-1]
    dev = L.require_cuda(feat_src, feat_ref)
    B, J, C = feat_src.shape
    K = feat_ref.shape[1]
    if metric == "l2":
        code = L.METRIC_L2
    elif metric == "angle":
        feat_src = feat_src / (torch.norm(feat_src, dim=-1, keepdim=True) + _EPS)
        feat_ref = feat_ref / (torch.norm(feat_ref, dim=-1, keepdim=True) + _EPS)
        code = L.METRIC_ACOS_DOT
    else:
        raise NotImplementedError
    (fs, a), (fr, b) = L.feat_nc(feat_src), L.feat_nc(feat_ref)
    return _dense(fs, fr, B, C, J, K, code, (a, b), dev)


def feat_dist(feat_src, feat_ref, metric="sqeuclidean"):
    """network/matchnet.py:147-192 without the [B,C,J,K] intermediate.  [B,C,J],[B,C,K] -> [B,J,K]."""
    assert feat_src.shape[1] == feat_ref.shape[1]
    if metric == "angle":
        return match_features_V2(feat_src, feat_ref, "angle")
    code = {"sqeuclidean": L.METRIC_SQDIFF, "euclidean": L.METRIC_SQDIFF_SQRT, "cityblock": L.METRIC_CITYBLOCK}.get(metric)
    if code is None:
        raise NotImplementedError("The following metric is not implemented: {}".format(metric))
    dev = L.require_cuda(feat_src, feat_ref)
    (fs, a), (fr, b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
    return _dense(fs, fr, feat_src.shape[0], feat_src.shape[1], feat_src.shape[2], feat_ref.shape[2], code, (a, b), dev)


def match_argmin(feat_src, feat_ref, return_min=False, algo=L.MATCH_AUTO, return_rescued=False, timing=None, prior=None):
    """The fused replacement of network/model.py:558-569 (chunked match_features_V2 + .min(dim=2)[1]):
    feat_src [B,C,J], feat_ref [B,C,K] -> indexs int64 [B,J]; the [J,K] matrix is never written.
    prior: optional int64 [B,J] correspondences of the previous loop iteration - a hint that speeds the filter up and
    never changes the result."""
    assert feat_src.shape[1] == feat_ref.shape[1]
    dev = L.require_cuda(feat_src, feat_ref)
    B, C, J = feat_src.shape
    K = feat_ref.shape[2]
    (fs, a), (fr, b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
    idx = torch.empty(B, J, dtype=torch.int64, device=dev)
    mind = torch.empty(B, J, dtype=torch.float32, device=dev) if return_min else None
    lib = L.lib()
    ws = L.workspace(lib.dsir_match_argmin_workspace_bytes(B, C, J, K, algo), dev)
    if prior is not None:
        prior = prior.to(torch.int64).contiguous()
        assert prior.shape == (B, J)
    L.check(lib.dsir_match_argmin_hint(fs, fr, B, C, J, K, idx.data_ptr(), L.ptr(mind), L.ptr(prior), ws.data_ptr(), ws.numel(),
                                       algo, L.stream_ptr(dev)), "dsir_match_argmin")
    if timing is not None:  # diagnostic: device-side span/cycles of the tcgen05 filter kernel; forces a stream sync
        import ctypes
        t = (ctypes.c_double * 3)()
        L.check(lib.dsir_match_argmin_filter_timing(ws.data_ptr(), ws.numel(), B, C, J, K, ctypes.addressof(t),
                                                    L.stream_ptr(dev)), "dsir_match_argmin_filter_timing")
        timing.update(span_ns=t[0], cycles_per_cta=t[1], cycles_per_unit=t[2])
    if return_rescued:  # diagnostic of the tcgen05 path; forces a stream sync
        import ctypes
        n = ctypes.c_int32(-1)
        L.check(lib.dsir_match_argmin_rescued_rows(ws.data_ptr(), ws.numel(), B, C, J, K, ctypes.addressof(n),
                                                   L.stream_ptr(dev)), "dsir_match_argmin_rescued_rows")
        return (idx, mind, n.value) if return_min else (idx, n.value)
    return (idx, mind) if return_min else idx


def compute_affinity(beta, feat_distance, alpha=0.5):
    """network/matchnet.py:195-208 (elementwise; kept for signature parity — the fused path is match_soft)."""
    if isinstance(alpha, float):
        return -beta[:, None, None] * (feat_distance - alpha)
    return -beta[:, None, None] * (feat_distance - alpha[:, None, None])


def match_soft(feat_src, feat_ref, xyz_ref, beta, alpha=0.5, col_bias=None, topk=0):
    """Fused compute_affinity (matchnet.py:195-208) + row softmax (sinkhorn row pass, :259) + soft target
    (network/model.py:81-84).  feat [B,C,J],[B,C,K]; xyz_ref [B,K,3]; beta [B]; alpha float or [B].
    Returns (y_soft [B,J,3], rowmass [B,J], lse [B,J]); with topk > 0 additionally (topk_idx int64 [B,J,topk],
    topk_w [B,J,topk]): the topk largest soft weights of every row, descending, ties to the lower index."""
    dev = L.require_cuda(feat_src, feat_ref, xyz_ref, beta)
    B, C, J = feat_src.shape
    K = feat_ref.shape[2]
    (fs, a), (fr, b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
    beta = beta.to(torch.float32).contiguous()
    alpha_t = torch.full((B,), float(alpha), dtype=torch.float32, device=dev) if isinstance(alpha, float) \
        else alpha.to(torch.float32).contiguous()
    xyz_ref = xyz_ref.contiguous() if xyz_ref is not None else None
    cb = col_bias.contiguous() if col_bias is not None else None
    y = torch.empty(B, J, 3, dtype=torch.float32, device=dev) if xyz_ref is not None else None
    lse = torch.empty(B, J, dtype=torch.float32, device=dev)
    tk_i = torch.empty(B, J, topk, dtype=torch.int64, device=dev) if topk > 0 else None
    tk_w = torch.empty(B, J, topk, dtype=torch.float32, device=dev) if topk > 0 else None
    lib = L.lib()
    ws = L.workspace(lib.dsir_match_soft_topk_workspace_bytes(B, C, J, K, topk), dev)
    L.check(lib.dsir_match_soft(fs, fr, B, C, J, K, beta.data_ptr(), alpha_t.data_ptr(), L.ptr(cb), L.ptr(xyz_ref),
                                L.ptr(y), lse.data_ptr(), topk, L.ptr(tk_i), L.ptr(tk_w), ws.data_ptr(), ws.numel(),
                                L.stream_ptr(dev)), "dsir_match_soft")
    ones = torch.ones(B, J, dtype=torch.float32, device=dev)
    return (y, ones, lse, tk_i, tk_w) if topk > 0 else (y, ones, lse)


def sinkhorn_implicit(feat_src, feat_ref, xyz_ref, beta, alpha=0.5, n_iters=5, slack=True):
    """Sinkhorn (network/matchnet.py:211-271) on the NEVER-MATERIALISED affinity a_jk = -beta (d_jk - alpha)
    (matchnet.py:195-208) in its dual form: log P_jk = a_jk - u_j - v_k with
        u_j = LSE_k(a_jk - v_k (, 0 with slack)),   v_k = LSE_j(a_jk - u_j (, 0 with slack)).
    Every half-step is one fused distance + log-sum-exp sweep (dsir_match_soft with a column bias; the column step is
    the same kernel with the roles of the two clouds swapped, since d_jk is symmetric).  Returns
    (y_soft [B,J,3] = sum_k P_jk r_k / sum_k P_jk, rowmass [B,J] = sum_k P_jk, u [B,J], v [B,K]) -- the inputs of
    kabsch_soft / compute_rigid_transform (network/model.py:68-116) -- without ever forming [B,J,K]."""
    dev = L.require_cuda(feat_src, feat_ref, xyz_ref, beta)
    B, C, J = feat_src.shape
    K = feat_ref.shape[2]
    u = torch.zeros(B, J, dtype=torch.float32, device=dev)
    v = torch.zeros(B, K, dtype=torch.float32, device=dev)
    zero = torch.zeros((), dtype=torch.float32, device=dev)
    lib = L.lib()
    (fs, _a), (fr, _b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
    beta_t = beta.to(torch.float32).contiguous()
    alpha_t = torch.full((B,), float(alpha), dtype=torch.float32, device=dev) if isinstance(alpha, float) \
        else alpha.to(torch.float32).contiguous()
    xyz_c = xyz_ref.contiguous()
    # one workspace per direction: after its first sweep the norms / tensor-core operands in it are re-used (only the
    # column bias changes between the half-steps)
    ws_row = L.workspace(lib.dsir_match_soft_workspace_bytes(B, C, J, K), dev)
    ws_col = L.workspace(lib.dsir_match_soft_workspace_bytes(B, C, K, J), dev)
    lse_r = torch.empty(B, J, dtype=torch.float32, device=dev)
    lse_c = torch.empty(B, K, dtype=torch.float32, device=dev)

    def sweep(row, bias, reuse, y=None):
        a, b2, n1, n2, ws, out = (fs, fr, J, K, ws_row, lse_r) if row else (fr, fs, K, J, ws_col, lse_c)
        bias = bias.contiguous()
        L.check(lib.dsir_match_soft_sweep(a, b2, B, C, n1, n2, beta_t.data_ptr(), alpha_t.data_ptr(), bias.data_ptr(),
                                          xyz_c.data_ptr() if y is not None else None, L.ptr(y), out.data_ptr(), int(reuse),
                                          ws.data_ptr(), ws.numel(), L.stream_ptr(dev)), "dsir_match_soft_sweep")
        return out

    for it in range(n_iters):
        u = sweep(True, -v, it > 0)                                                          # [B,J]
        u = torch.logaddexp(u, zero) if slack else u.clone()
        v = sweep(False, -u, it > 0)                                                         # [B,K]
        v = torch.logaddexp(v, zero) if slack else v.clone()
    y = torch.empty(B, J, 3, dtype=torch.float32, device=dev)
    lse_r = sweep(True, -v, n_iters > 0, y=y)
    return y, torch.exp(lse_r - u), u, v


def log_optimal_transport_implicit(feat_src, feat_ref, xyz_ref, beta, alpha, bin_score, iters=5):
    """log_optimal_transport (network/matchnet.py:836-856) on the NEVER-MATERIALISED scores s_jk = -beta (d_jk - alpha)
    (compute_affinity, matchnet.py:195-208).  The dustbin row / column of the augmented couplings are the constant
    `bin_score` plus the potentials, so every half-iteration is one fused tcgen05 distance + log-sum-exp sweep
    (dsir_match_soft_sweep with a column bias) and a few vector operations on [B,M+1] / [B,N+1]:
        u_j = log_mu_j - logaddexp(LSE_k(s_jk + v_k), bin + v_N)   (j < M),   u_M = log_mu_M - (bin + LSE(v))
        v_k = log_nu_k - logaddexp(LSE_j(s_jk + u_j), bin + u_M)   (k < N),   v_N = log_nu_N - (bin + LSE(u))
    Returns (y_soft [B,M,3] = sum_k P_jk r_k / sum_k P_jk over the real columns, rowmass [B,M] = sum_{k<N} P_jk,
    u [B,M+1], v [B,N+1]) with log P_jk = s_jk + u_j + v_k - norm, norm = -log(M+N) - i.e. log_optimal_transport(...)[:, :M, :N]
    without ever forming it."""
    import math
    dev = L.require_cuda(feat_src, feat_ref, xyz_ref, beta)
    B, C, M = feat_src.shape
    N = feat_ref.shape[2]
    lib = L.lib()
    (fs, _a), (fr, _b) = L.feat_cn(feat_src), L.feat_cn(feat_ref)
    beta_t = beta.to(torch.float32).contiguous()
    alpha_t = torch.full((B,), float(alpha), dtype=torch.float32, device=dev) if isinstance(alpha, float) \
        else alpha.to(torch.float32).contiguous()
    bin_t = bin_score.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1] if isinstance(bin_score, torch.Tensor) \
        else torch.full((1,), float(bin_score), dtype=torch.float32, device=dev)
    xyz_c = xyz_ref.contiguous()
    ws_row = L.workspace(lib.dsir_match_soft_workspace_bytes(B, C, M, N), dev)
    ws_col = L.workspace(lib.dsir_match_soft_workspace_bytes(B, C, N, M), dev)
    lse_r = torch.empty(B, M, dtype=torch.float32, device=dev)
    lse_c = torch.empty(B, N, dtype=torch.float32, device=dev)

    def sweep(row, bias, reuse, y=None):
        a, b2, n1, n2, ws, out = (fs, fr, M, N, ws_row, lse_r) if row else (fr, fs, N, M, ws_col, lse_c)
        bias = bias.contiguous()
        L.check(lib.dsir_match_soft_sweep(a, b2, B, C, n1, n2, beta_t.data_ptr(), alpha_t.data_ptr(), bias.data_ptr(),
                                          xyz_c.data_ptr() if y is not None else None, L.ptr(y), out.data_ptr(), int(reuse),
                                          ws.data_ptr(), ws.numel(), L.stream_ptr(dev)), "dsir_match_soft_sweep")
        return out

    norm = -math.log(M + N)
    log_mu = torch.full((B, M + 1), norm, dtype=torch.float32, device=dev)
    log_mu[:, M] = math.log(N) + norm
    log_nu = torch.full((B, N + 1), norm, dtype=torch.float32, device=dev)
    log_nu[:, N] = math.log(M) + norm
    u = torch.zeros(B, M + 1, dtype=torch.float32, device=dev)
    v = torch.zeros(B, N + 1, dtype=torch.float32, device=dev)
    for it in range(iters):
        r = sweep(True, v[:, :N], it > 0)                                          # LSE_k(s_jk + v_k), k < N
        u = torch.cat([log_mu[:, :M] - torch.logaddexp(r, bin_t + v[:, N:]),
                       log_mu[:, M:] - (bin_t + torch.logsumexp(v, dim=1, keepdim=True))], dim=1)
        c = sweep(False, u[:, :M], it > 0)                                         # LSE_j(s_jk + u_j), j < M
        v = torch.cat([log_nu[:, :N] - torch.logaddexp(c, bin_t + u[:, M:]),
                       log_nu[:, N:] - (bin_t + torch.logsumexp(u, dim=1, keepdim=True))], dim=1)
    y = torch.empty(B, M, 3, dtype=torch.float32, device=dev)
    r = sweep(True, v[:, :N], iters > 0, y=y)
    return y, torch.exp(r + u[:, :M] - norm), u, v


def gather_neighbour_V3(inputs, neigh_idx):
    """network/tools.py:211-221.  inputs [B,C,N], neigh_idx [B,M] int64 -> [B,C,M]."""
    dev = L.require_cuda(inputs, neigh_idx)
    B, C, N = inputs.shape
    M = neigh_idx.shape[1]
    inputs = inputs.contiguous()
    neigh_idx = neigh_idx.to(torch.int64).contiguous()
    out = torch.empty(B, C, M, dtype=torch.float32, device=dev)
    L.check(L.lib().dsir_gather_points(inputs.data_ptr(), B, C, N, neigh_idx.data_ptr(), M, out.data_ptr(),
                                       L.stream_ptr(dev)), "dsir_gather_points")
    return out
