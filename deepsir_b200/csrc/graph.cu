// KNN consumers of RandLA-Net local aggregation and Sinkhorn normalisation: HBM-bound gather / reduce kernels.
//
//   gather_neighbours   network/tools.py:197-209 (gather_neighbour_V2)         [B,C,N], idx [B,M,k] -> [B,C,M,k]
//   rel_pos_encoding    network/RandLANet.py:197-212 (relative_pos_encoding)   [B,3,N], idx [B,N,k] -> [B,10,N,k]
//   pool_max            network/RandLANet.py:374-391 (random_sample)           [B,C,N], idx [B,M,k] -> [B,C,M]
//   sinkhorn            network/matchnet.py:211-271                            [B,J,K] -> [B,J,K]
//
// The reference materialises a k-fold copy of every tensor with repeat + gather before it reduces; here every
// output element is produced from one index read and one (cached) value read, and the pooled maximum never writes the
// [B,C,M,k] intermediate.  Sinkhorn is run in its dual form: the iterate is log_alpha - u_j - v_k, so one half-step is
// ONE read-only sweep that refreshes u (row log-sum-exp) or v (column log-sum-exp); the matrix is written once.
#include "graph.cuh"

namespace dsir {

namespace {

// one thread per (m, j) neighbour slot, looping over channels: index read once, output writes coalesced along (m,j)
__global__ void gather_neighbours_kernel(const float *__restrict__ in, int C, int N, const int64_t *__restrict__ idx, long long MK,
                                         float *__restrict__ out) {
    const int b = blockIdx.y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= MK) return;
    const long long n = idx[(size_t)b * MK + t];
    const float *src = in + (size_t)b * C * N;
    float *dst = out + (size_t)b * C * MK + t;
    const bool ok = n >= 0 && n < N;
    for (int c = 0; c < C; ++c) dst[(size_t)c * MK] = ok ? src[(size_t)c * N + n] : 0.f;
}

__global__ void rel_pos_encoding_kernel(const float *__restrict__ xyz, int N, const int64_t *__restrict__ idx, int k,
                                        float *__restrict__ out) {
    const int b = blockIdx.y;
    const long long NK = (long long)N * k;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NK) return;
    const int n = (int)(t / k);
    const long long m = idx[(size_t)b * NK + t];
    const float *p = xyz + (size_t)b * 3 * N;
    const bool ok = m >= 0 && m < N;
    const float cx = p[n], cy = p[N + n], cz = p[2 * (size_t)N + n];
    const float nx = ok ? p[m] : 0.f, ny = ok ? p[N + m] : 0.f, nz = ok ? p[2 * (size_t)N + m] : 0.f;
    const float rx = __fsub_rn(nx, cx), ry = __fsub_rn(ny, cy), rz = __fsub_rn(nz, cz);
    // torch.sum(torch.pow(rel, 2), dim=1): x^2 + y^2 + z^2 left to right, then sqrt
    const float dis = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    float *o = out + (size_t)b * 10 * NK + t;
    o[0] = dis;
    o[1 * NK] = rx; o[2 * NK] = ry; o[3 * NK] = rz;
    o[4 * NK] = cx; o[5 * NK] = cy; o[6 * NK] = cz;
    o[7 * NK] = nx; o[8 * NK] = ny; o[9 * NK] = nz;
}

// one thread per pooled point m, looping over channels; the k indices live in registers
template <int KMAX>
__global__ void pool_max_kernel(const float *__restrict__ in, int C, int N, const int64_t *__restrict__ idx, int M, int k,
                                float *__restrict__ out) {
    const int b = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int nb[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        long long v = j < k ? idx[((size_t)b * M + m) * k + j] : -1;
        nb[j] = (v >= 0 && v < N) ? (int)v : -1;
    }
    const float *src = in + (size_t)b * C * N;
    for (int c = 0; c < C; ++c) {
        float best = -INFINITY;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (nb[j] >= 0) {
                const float v = src[(size_t)c * N + nb[j]];
                best = (v > best || v != v) ? v : best;   // NaN propagates like torch.max
            }
        out[((size_t)b * C + c) * M + m] = best;
    }
}

// ---------------------------------------------------------------- Sinkhorn (dual form)
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
    const float M = fmaxf(m, m2);
    if (M == -INFINITY) { m = M; s = 0.f; return; }
    s = s * __expf(m - M) + s2 * __expf(m2 - M);
    m = M;
}

// u[j] = LSE_k(A[j,k] - v[k]  (, 0 when slack)) : one warp per row
__global__ void sinkhorn_row_kernel(const float *__restrict__ A, int J, int K, const float *__restrict__ v, float *__restrict__ u,
                                    int slack) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= J) return;
    const float *row = A + ((size_t)b * J + j) * K;
    const float *vb = v + (size_t)b * K;
    float m = -INFINITY, s = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float x = row[k] - vb[k];
        if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
        else if (x > -INFINITY) s += __expf(x - m);
        else if (x != x) { m = x; s = x; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    if (slack) lse_merge(m, s, 0.f, 1.f);
    if (lane == 0) u[(size_t)b * J + j] = m + __logf(s);
}

// v[k] = LSE_j(A[j,k] - u[j]  (, 0 when slack)) : a block owns 32 columns, 8 row lanes
__global__ __launch_bounds__(256) void sinkhorn_col_kernel(const float *__restrict__ A, int J, int K, const float *__restrict__ u,
                                                           float *__restrict__ v, int slack) {
    __shared__ float sm[8][33], ss[8][33];
    const int b = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + tx;
    const float *Ab = A + (size_t)b * J * K;
    const float *ub = u + (size_t)b * J;
    float m = -INFINITY, s = 0.f;
    if (k < K)
        for (int j = ty; j < J; j += 8) {
            const float x = Ab[(size_t)j * K + k] - ub[j];
            if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
            else if (x > -INFINITY) s += __expf(x - m);
            else if (x != x) { m = x; s = x; }
        }
    sm[ty][tx] = m; ss[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && k < K) {
        for (int r = 1; r < 8; ++r) lse_merge(m, s, sm[r][tx], ss[r][tx]);
        if (slack) lse_merge(m, s, 0.f, 1.f);
        v[(size_t)b * K + k] = m + __logf(s);
    }
}

__global__ void sinkhorn_apply_kernel(const float *__restrict__ A, int J, int K, const float *__restrict__ u, const float *__restrict__ v,
                                      float *__restrict__ out) {
    const int b = blockIdx.z, j = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const size_t o = ((size_t)b * J + j) * K + k;
    out[o] = A[o] - u[(size_t)b * J + j] - v[(size_t)b * K + k];
}

// ---------------------------------------------------------------- log_optimal_transport (network/matchnet.py:827-856)
// SuperGlue's dustbin OT on the scores S [B,M,N]: couplings Z = [[S, alpha], [alpha, alpha]] of shape [M+1, N+1], marginals
// log_mu = (norm x M, log N + norm), log_nu = (norm x N, log M + norm), norm = -log(M + N); `iters` times
//     u = log_mu - LSE_k(Z + v),   v = log_nu - LSE_j(Z + u);    result Z + u + v - norm.
// The augmented matrix is never built: the dustbin row / column are the constant alpha plus the potentials.
__device__ __forceinline__ void lse_add(float &m, float &s, float x) {
    if (x > m) { s = s * __expf(m - x) + 1.f; m = x; }
    else if (x > -INFINITY) s += __expf(x - m);
    else if (x != x) { m = x; s = x; }
}
// u[j], j in [0, M]: one warp per row of the augmented matrix
__global__ void logot_row_kernel(const float *__restrict__ S, int M, int N, const float *__restrict__ alpha_p, const float *__restrict__ v,
                                 float *__restrict__ u, float norm) {
    const float alpha = alpha_p[0];
    const int b = blockIdx.y;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j > M) return;
    const float *vb = v + (size_t)b * (N + 1);
    float m = -INFINITY, s = 0.f;
    if (j < M) {
        const float *row = S + ((size_t)b * M + j) * N;
        for (int k = lane; k < N; k += 32) lse_add(m, s, row[k] + vb[k]);
        if (lane == 0) lse_add(m, s, alpha + vb[N]);
    } else {
        for (int k = lane; k <= N; k += 32) lse_add(m, s, alpha + vb[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    if (lane == 0) u[(size_t)b * (M + 1) + j] = (j < M ? norm : __logf((float)N) + norm) - (m + __logf(s));
}
// v[k], k in [0, N]: a block owns 32 columns of the augmented matrix, 8 row lanes
__global__ __launch_bounds__(256) void logot_col_kernel(const float *__restrict__ S, int M, int N, const float *__restrict__ alpha_p,
                                                        const float *__restrict__ u, float *__restrict__ v, float norm) {
    __shared__ float sm[8][33], ss[8][33];
    const float alpha = alpha_p[0];
    const int b = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + tx;
    const float *Sb = S + (size_t)b * M * N;
    const float *ub = u + (size_t)b * (M + 1);
    float m = -INFINITY, s = 0.f;
    if (k < N) {
        for (int j = ty; j < M; j += 8) lse_add(m, s, Sb[(size_t)j * N + k] + ub[j]);
        if (ty == 0) lse_add(m, s, alpha + ub[M]);
    } else if (k == N) {
        for (int j = ty; j <= M; j += 8) lse_add(m, s, alpha + ub[j]);
    }
    sm[ty][tx] = m; ss[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && k <= N) {
        for (int r = 1; r < 8; ++r) lse_merge(m, s, sm[r][tx], ss[r][tx]);
        v[(size_t)b * (N + 1) + k] = (k < N ? norm : __logf((float)M) + norm) - (m + __logf(s));
    }
}
__global__ void logot_apply_kernel(const float *__restrict__ S, int M, int N, const float *__restrict__ alpha_p, const float *__restrict__ u,
                                   const float *__restrict__ v, float norm, float *__restrict__ out) {
    const float alpha = alpha_p[0];
    const int b = blockIdx.z, j = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > N) return;
    const float z = (j < M && k < N) ? S[((size_t)b * M + j) * N + k] : alpha;
    out[((size_t)b * (M + 1) + j) * (N + 1) + k] = z + u[(size_t)b * (M + 1) + j] + v[(size_t)b * (N + 1) + k] - norm;
}

}  // namespace

int launch_log_ot(const float *scores, int B, int M, int N, const float *alpha, int iters, float *out, float *u, float *v, cudaStream_t st) {
    const float norm = -logf((float)M + (float)N);
    DSIR_CUDA_TRY(cudaMemsetAsync(u, 0, (size_t)B * (M + 1) * sizeof(float), st));
    DSIR_CUDA_TRY(cudaMemsetAsync(v, 0, (size_t)B * (N + 1) * sizeof(float), st));
    for (int it = 0; it < iters; ++it) {
        logot_row_kernel<<<dim3(cdiv(M + 1, 8), B), 256, 0, st>>>(scores, M, N, alpha, v, u, norm);
        DSIR_LAUNCH_CHECK();
        logot_col_kernel<<<dim3(cdiv(N + 1, 32), B), 256, 0, st>>>(scores, M, N, alpha, u, v, norm);
        DSIR_LAUNCH_CHECK();
    }
    logot_apply_kernel<<<dim3(cdiv(N + 1, 256), M + 1, B), 256, 0, st>>>(scores, M, N, alpha, u, v, norm, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_gather_neighbours(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out, cudaStream_t st) {
    const long long MK = (long long)M * k;
    if (MK <= 0) return DSIR_OK;
    dim3 grid((unsigned)((MK + 255) / 256), B);
    gather_neighbours_kernel<<<grid, 256, 0, st>>>(in, C, N, idx, MK, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_rel_pos_encoding(const float *xyz, int B, int N, const int64_t *idx, int k, float *out, cudaStream_t st) {
    const long long NK = (long long)N * k;
    if (NK <= 0) return DSIR_OK;
    dim3 grid((unsigned)((NK + 255) / 256), B);
    rel_pos_encoding_kernel<<<grid, 256, 0, st>>>(xyz, N, idx, k, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_pool_max(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out, cudaStream_t st) {
    if (M <= 0) return DSIR_OK;
    if (k > 32) return DSIR_ERR_UNSUPPORTED;
    dim3 grid(cdiv(M, 128), B);
    if (k <= 16) pool_max_kernel<16><<<grid, 128, 0, st>>>(in, C, N, idx, M, k, out);
    else pool_max_kernel<32><<<grid, 128, 0, st>>>(in, C, N, idx, M, k, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_sinkhorn(const float *log_alpha, int B, int J, int K, int n_iters, int slack, float *out, float *u, float *v,
                    cudaStream_t st) {
    DSIR_CUDA_TRY(cudaMemsetAsync(u, 0, (size_t)B * J * sizeof(float), st));
    DSIR_CUDA_TRY(cudaMemsetAsync(v, 0, (size_t)B * K * sizeof(float), st));
    for (int it = 0; it < n_iters; ++it) {
        sinkhorn_row_kernel<<<dim3(cdiv(J, 8), B), 256, 0, st>>>(log_alpha, J, K, v, u, slack);
        DSIR_LAUNCH_CHECK();
        sinkhorn_col_kernel<<<dim3(cdiv(K, 32), B), 256, 0, st>>>(log_alpha, J, K, u, v, slack);
        DSIR_LAUNCH_CHECK();
    }
    sinkhorn_apply_kernel<<<dim3(cdiv(K, 256), J, B), 256, 0, st>>>(log_alpha, J, K, u, v, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
