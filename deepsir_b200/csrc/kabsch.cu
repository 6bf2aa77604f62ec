// Weighted Kabsch / Procrustes on the device, replacing compute_rigid_transform_2
// (network/model.py:22-66 of the reference) and its host round trip:
//   kernel 1  warp-shuffle reduction of the 17 additive raw moments in fp64, fused with the
//             correspondence gather (model.py:571) and any [B,M,3] / [B,3,M] layout
//   kernel 2  per-pair: centred covariance from the moments, fp64 one-sided Jacobi SVD of the 3x3
//             (stands in for LAPACK gesdd at model.py:47), determinant fix (:49-54), R, t (:57-58)
// plus the SE(3) helpers of common/math/se3_torch.py.
#include "kabsch.cuh"

namespace dsir {

constexpr int KB_THREADS = 256;

__global__ __launch_bounds__(KB_THREADS) void kabsch_moments_kernel(KabschParams P) {
    const int b = blockIdx.y;
    const float *sp = P.src.ptr + (size_t)b * P.src.batch_stride;
    const float *tp = P.tgt.ptr + (size_t)b * P.tgt.batch_stride;
    const float *wp = P.w ? P.w + (size_t)b * P.w_bs : nullptr;
    const int64_t *gp = P.gather ? P.gather + (size_t)b * P.M : nullptr;

    double acc[KB_NMOM];
#pragma unroll
    for (int i = 0; i < KB_NMOM; ++i) acc[i] = 0.0;

    for (int m = blockIdx.x * KB_THREADS + threadIdx.x; m < P.M; m += gridDim.x * KB_THREADS) {
        float wf = wp ? wp[m] : 1.f;
        size_t tm = gp ? (size_t)gp[m] : (size_t)m;
        // torch.gather raises on an index outside the target cloud (tools.py:211-221); here it is never dereferenced and
        // poisons the pair instead: NaN moments -> status 1 / invalid_gradient
        const bool bad = gp != nullptr && P.n_tgt > 0 && (unsigned long long)gp[m] >= (unsigned long long)P.n_tgt;
        if (bad) tm = 0;
        const float *s = sp + (size_t)m * P.src.point_stride;
        const float *t = tp + tm * P.tgt.point_stride;
        double w = (double)wf;
        double x0 = s[0], x1 = s[P.src.coord_stride], x2 = s[2 * P.src.coord_stride];
        double y0 = t[0], y1 = t[P.tgt.coord_stride], y2 = t[2 * P.tgt.coord_stride];
        if (bad) y0 = nan("");
        acc[0] += fabs(w);
        acc[1] += w;
        double wx0 = w * x0, wx1 = w * x1, wx2 = w * x2;
        acc[2] += wx0; acc[3] += wx1; acc[4] += wx2;
        acc[5] += w * y0; acc[6] += w * y1; acc[7] += w * y2;
        acc[8] = fma(wx0, y0, acc[8]);   acc[9] = fma(wx0, y1, acc[9]);   acc[10] = fma(wx0, y2, acc[10]);
        acc[11] = fma(wx1, y0, acc[11]); acc[12] = fma(wx1, y1, acc[12]); acc[13] = fma(wx1, y2, acc[13]);
        acc[14] = fma(wx2, y0, acc[14]); acc[15] = fma(wx2, y1, acc[15]); acc[16] = fma(wx2, y2, acc[16]);
    }

    __shared__ double red[KB_THREADS / 32][KB_NMOM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < KB_NMOM; ++i) {
        double v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < KB_NMOM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < KB_THREADS / 32; ++w) v += red[w][threadIdx.x];
        P.partials[((size_t)b * gridDim.x + blockIdx.x) * KB_NMOM + threadIdx.x] = v;
    }
}

// sum the per-block partials of one pair in block order (deterministic): out[b][17]
__global__ void kabsch_reduce_kernel(const double *__restrict__ partials, int nblk, double *__restrict__ out) {
    const int b = blockIdx.x;
    if (threadIdx.x < KB_NMOM) {
        double v = 0.0;
        for (int i = 0; i < nblk; ++i) v += partials[((size_t)b * nblk + i) * KB_NMOM + threadIdx.x];
        out[(size_t)b * KB_NMOM + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 solve
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// One-sided (Hestenes) Jacobi: rotate column pairs of G = H V until they are orthogonal.
// Afterwards |G[:,i]| are the singular values, G[:,i]/|G[:,i]| the left and V[:,i] the right vectors.
__device__ void svd3_jacobi(const double H[3][3], double G[3][3], double V[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { G[i][j] = H[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            double a = 0, bb = 0, g = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) { a += G[i][p] * G[i][p]; bb += G[i][q] * G[i][q]; g += G[i][p] * G[i][q]; }
            if (g == 0.0 || fabs(g) <= 2e-16 * sqrt(a * bb)) continue;
            rotated = true;
            double zeta = (bb - a) / (2.0 * g);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double gp = G[i][p], gq = G[i][q];
                G[i][p] = c * gp - s * gq;
                G[i][q] = s * gp + c * gq;
                double vp = V[i][p], vq = V[i][q];
                V[i][p] = c * vp - s * vq;
                V[i][q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
}

// moments -> T (fp32 [3,4]) ; returns status (0 ok, 1 non-finite input -> identity).
// Rank-deficient covariances (collinear or duplicated correspondences, a single non-zero weight, all-zero weights): the
// reference still returns V U^T of whatever basis LAPACK completes the null space with and t = c_tgt - R c_src, and flags
// nothing (model.py:47-58 raises only when the SVD itself fails).  The completion here is deterministic: rank 1 -> the
// MINIMAL rotation that takes u1 to v1 (the rotation is only determined up to a turn about that axis); rank 0 -> R = I.
// The centroid translation is kept in both cases.
__device__ int kabsch_solve(const double *mom, float *T, int signed_norm) {
    const double EPS = 1e-16;  // _EPS of network/model.py:19
    // hard variant normalises by sum|w| (:36); the soft variant by the plain sum of the row masses (:82)
    double S = (signed_norm ? mom[1] : mom[0]) + EPS;
    double cs[3] = {mom[2] / S, mom[3] / S, mom[4] / S};  // centroid_src = sum(src * w/(sum|w|+eps))  (:36,38)
    double ct[3] = {mom[5] / S, mom[6] / S, mom[7] / S};  // centroid_tgt                               (:39)
    double sw = mom[1] / S;                               // sum of the normalised weights
    // cov = sum wn (x - cs)(y - ct)^T = Sxy/S - cs (Swy/S)^T - (Swx/S) ct^T + sw cs ct^T                (:40-42)
    double H[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) H[i][j] = mom[8 + 3 * i + j] / S - (2.0 - sw) * cs[i] * ct[j];

    bool finite = true;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) finite = finite && isfinite(H[i][j]);
    finite = finite && isfinite(cs[0]) && isfinite(cs[1]) && isfinite(cs[2]) && isfinite(ct[0]) && isfinite(ct[1]) &&
             isfinite(ct[2]);

    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    int status = 1;
    if (finite) {
        double G[3][3], V[3][3];
        svd3_jacobi(H, G, V);
        double sg[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) sg[i] = sqrt(G[0][i] * G[0][i] + G[1][i] * G[1][i] + G[2][i] * G[2][i]);
        // indices of the two largest singular values
        int i1 = 0;
        if (sg[1] > sg[i1]) i1 = 1;
        if (sg[2] > sg[i1]) i1 = 2;
        int i2 = (i1 == 0) ? 1 : 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i != i1 && sg[i] > sg[i2]) i2 = i;
        if (sg[i1] > 0.0 && sg[i2] > 1e-12 * sg[i1]) {
            double u1[3], u2[3], v1[3], v2[3], u3[3], v3[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { u1[i] = G[i][i1] / sg[i1]; u2[i] = G[i][i2]; v1[i] = V[i][i1]; v2[i] = V[i][i2]; }
            double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
            for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
            double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
#pragma unroll
            for (int i = 0; i < 3; ++i) u2[i] /= n2;
            cross3(u1, u2, u3);
            cross3(v1, v2, v3);
            // R = V U^T with the sign of the third pair chosen so det R = +1: equals the reference's
            // "V U^T, else flip V[:, :, 2]" rule (:49-54) because u3 = +-u1xu2 and v3 = +-v1xv2.
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) R[i][j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
            status = 0;
        } else if (sg[i1] > 0.0) {
            // rank 1: minimal rotation u1 -> v1 (Rodrigues); antiparallel: half turn about an axis orthogonal to u1
            double u1[3], v1[3], w[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { u1[i] = G[i][i1] / sg[i1]; v1[i] = V[i][i1]; }
            const double c = u1[0] * v1[0] + u1[1] * v1[1] + u1[2] * v1[2];
            cross3(u1, v1, w);
            if (c > -1.0 + 1e-12) {
                const double k = 1.0 / (1.0 + c);
                const double K[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        double kk = 0;
#pragma unroll
                        for (int t = 0; t < 3; ++t) kk += K[i][t] * K[t][j];
                        R[i][j] = (i == j ? 1.0 : 0.0) + K[i][j] + k * kk;
                    }
            } else {
                double a[3] = {fabs(u1[0]) < 0.6 ? 1.0 : 0.0, fabs(u1[0]) < 0.6 ? 0.0 : 1.0, 0.0}, n[3];
                cross3(u1, a, n);
                const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) R[i][j] = 2.0 * n[i] * n[j] / (nn * nn) - (i == j ? 1.0 : 0.0);
            }
            status = 0;
        } else {
            status = 0;   // rank 0 (zero covariance): R = I, translation between the centroids
        }
    }
    if (status != 0) { cs[0] = cs[1] = cs[2] = 0.0; ct[0] = ct[1] = ct[2] = 0.0; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // t = -R c_src + c_tgt                                                                           (:57)
        double t = ct[i] - (R[i][0] * cs[0] + R[i][1] * cs[1] + R[i][2] * cs[2]);
        T[4 * i + 0] = (float)R[i][0];
        T[4 * i + 1] = (float)R[i][1];
        T[4 * i + 2] = (float)R[i][2];
        T[4 * i + 3] = (float)t;
    }
    return status;
}

// one thread per pair; partial blocks are summed in block order first
__global__ void kabsch_solve_kernel(const double *__restrict__ partials, int nblk, int B, float *__restrict__ T,
                                    int32_t *__restrict__ status, double *__restrict__ moments_out,
                                    const float *__restrict__ compose_with, float *__restrict__ composed, int signed_norm) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double mom[KB_NMOM];
#pragma unroll
    for (int i = 0; i < KB_NMOM; ++i) mom[i] = 0.0;
    for (int k = 0; k < nblk; ++k)
#pragma unroll
        for (int i = 0; i < KB_NMOM; ++i) mom[i] += partials[((size_t)b * nblk + k) * KB_NMOM + i];
    if (moments_out)
#pragma unroll
        for (int i = 0; i < KB_NMOM; ++i) moments_out[(size_t)b * KB_NMOM + i] = mom[i];
    float Tl[12];
    int st = kabsch_solve(mom, Tl, signed_norm);
#pragma unroll
    for (int i = 0; i < 12; ++i) T[(size_t)b * 12 + i] = Tl[i];
    if (status) status[b] = st;
    if (composed) {  // composed = T o compose_with  (se3_torch.concatenate, se3_torch.py:28-48); null -> T
        float o[12];
        if (compose_with) {
            const float *P = compose_with + (size_t)b * 12;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    o[4 * i + j] = __fmaf_rn(Tl[4 * i + 2], P[8 + j], __fmaf_rn(Tl[4 * i + 1], P[4 + j], __fmul_rn(Tl[4 * i], P[j])));
                o[4 * i + 3] = __fadd_rn(__fmaf_rn(Tl[4 * i + 2], P[11], __fmaf_rn(Tl[4 * i + 1], P[7], __fmul_rn(Tl[4 * i], P[3]))), Tl[4 * i + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) o[i] = Tl[i];
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) composed[(size_t)b * 12 + i] = o[i];
    }
}

// compute_rigid_transform's first two statements (network/model.py:81-84) for a GIVEN weight matrix W [B,M,N]:
//   rowmass_j = sum_k W_jk,   y_j = (sum_k W_jk tgt_k) / (rowmass_j + eps)
// One warp per row, 8 rows per block sharing target tiles staged in shared memory; W is read exactly once (HBM-bound).
constexpr int ST_ROWS = 8, ST_TILE = 1024;
__global__ __launch_bounds__(ST_ROWS * 32) void soft_targets_kernel(const float *__restrict__ W, long long w_bs, long long w_rs,
                                                                    const float *__restrict__ tgt, int M, int N,
                                                                    float *__restrict__ y, float *__restrict__ mass) {
    __shared__ float st[ST_TILE * 3];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * ST_ROWS + warp;
    const float *row = W + (size_t)b * w_bs + (size_t)(j < M ? j : 0) * w_rs;
    const float *tb = tgt + (size_t)b * N * 3;
    float s = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    for (int k0 = 0; k0 < N; k0 += ST_TILE) {
        const int nk = min(ST_TILE, N - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < nk * 3; t += ST_ROWS * 32) st[t] = tb[(size_t)k0 * 3 + t];
        __syncthreads();
        if (j < M)
            for (int k = lane; k < nk; k += 32) {
                const float w = row[k0 + k];
                s += w;
                ax = __fmaf_rn(w, st[3 * k], ax);
                ay = __fmaf_rn(w, st[3 * k + 1], ay);
                az = __fmaf_rn(w, st[3 * k + 2], az);
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        az += __shfl_xor_sync(0xffffffffu, az, o);
    }
    if (lane == 0 && j < M) {
        const float inv = 1.f / (s + 1e-16f);
        mass[(size_t)b * M + j] = s;
        y[((size_t)b * M + j) * 3 + 0] = ax * inv;
        y[((size_t)b * M + j) * 3 + 1] = ay * inv;
        y[((size_t)b * M + j) * 3 + 2] = az * inv;
    }
}

int launch_soft_targets(const float *W, long long w_bs, long long w_rs, const float *tgt, int B, int M, int N, float *y, float *mass,
                        cudaStream_t st) {
    dim3 grid(cdiv(M, ST_ROWS), B);
    soft_targets_kernel<<<grid, ST_ROWS * 32, 0, st>>>(W, w_bs, w_rs, tgt, M, N, y, mass);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int kabsch_num_blocks(int M) {
    int n = cdiv(M, KB_THREADS * 8);
    return n < 1 ? 1 : (n > 64 ? 64 : n);
}

int launch_kabsch_moments(const KabschParams &P, int nblk, cudaStream_t st) {
    dim3 grid(nblk, P.B);
    kabsch_moments_kernel<<<grid, KB_THREADS, 0, st>>>(P);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_kabsch_reduce(const double *partials, int nblk, int B, double *out, cudaStream_t st) {
    kabsch_reduce_kernel<<<B, 32, 0, st>>>(partials, nblk, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_kabsch_solve(const double *partials, int nblk, int B, float *T, int32_t *status, double *moments_out,
                        const float *compose_with, float *composed, int signed_norm, cudaStream_t st) {
    kabsch_solve_kernel<<<cdiv(B, 32), 32, 0, st>>>(partials, nblk, B, T, status, moments_out, compose_with, composed,
                                                    signed_norm);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

// ------------------------------------------------------------------------------------------------
// SE(3)
// ------------------------------------------------------------------------------------------------
__global__ void se3_apply_kernel(const float *__restrict__ T, long long T_bs, dsir_points pts, int N, float *__restrict__ out,
                                 long long o_bs, long long o_ps, long long o_cs, int rotate_only) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (n >= N) return;
    const float *t = T + (size_t)b * T_bs;
    const float *p = pts.ptr + (size_t)b * pts.batch_stride + (size_t)n * pts.point_stride;
    float x = p[0], y = p[pts.coord_stride], z = p[2 * pts.coord_stride];
    float *o = out + (size_t)b * o_bs + (size_t)n * o_ps;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // a @ R^T + t  (se3_torch.py:66) / R @ a + t (:94)
        float v = __fmaf_rn(t[4 * i + 2], z, __fmaf_rn(t[4 * i + 1], y, __fmul_rn(t[4 * i], x)));
        if (!rotate_only) v = __fadd_rn(v, t[4 * i + 3]);
        o[(size_t)i * o_cs] = v;
    }
}

int launch_se3_apply(const float *T, long long T_bs, dsir_points pts, int B, int N, float *out, long long o_bs,
                     long long o_ps, long long o_cs, int rotate_only, cudaStream_t st) {
    if (B <= 0 || N <= 0) return DSIR_OK;
    dim3 grid(cdiv(N, 256), B);
    se3_apply_kernel<<<grid, 256, 0, st>>>(T, T_bs, pts, N, out, o_bs, o_ps, o_cs, rotate_only);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

__global__ void se3_compose_kernel(const float *__restrict__ a, long long a_bs, const float *__restrict__ bq, long long b_bs,
                                   int B, float *__restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *A = a + (size_t)b * a_bs;
    const float *Q = bq + (size_t)b * b_bs;
    float o[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o[4 * i + j] = __fmaf_rn(A[4 * i + 2], Q[8 + j], __fmaf_rn(A[4 * i + 1], Q[4 + j], __fmul_rn(A[4 * i], Q[j])));
        o[4 * i + 3] = __fadd_rn(__fmaf_rn(A[4 * i + 2], Q[11], __fmaf_rn(A[4 * i + 1], Q[7], __fmul_rn(A[4 * i], Q[3]))), A[4 * i + 3]);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) out[(size_t)b * 12 + i] = o[i];
}

__global__ void se3_inverse_kernel(const float *__restrict__ T, long long T_bs, int B, float *__restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *A = T + (size_t)b * T_bs;
    float o[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o[4 * i + j] = A[4 * j + i];
        // R^T @ (-t)
        o[4 * i + 3] = __fmaf_rn(A[8 + i], -A[11], __fmaf_rn(A[4 + i], -A[7], __fmul_rn(A[i], -A[3])));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) out[(size_t)b * 12 + i] = o[i];
}

int launch_se3_compose(const float *a, long long a_bs, const float *b, long long b_bs, int B, float *out, cudaStream_t st) {
    if (B <= 0) return DSIR_OK;
    se3_compose_kernel<<<cdiv(B, 64), 64, 0, st>>>(a, a_bs, b, b_bs, B, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_se3_inverse(const float *T, long long T_bs, int B, float *out, cudaStream_t st) {
    if (B <= 0) return DSIR_OK;
    se3_inverse_kernel<<<cdiv(B, 64), 64, 0, st>>>(T, T_bs, B, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

// gather_neighbour_V3 (network/tools.py:211-221)
__global__ void gather_points_kernel(const float *__restrict__ in, int C, int N, const int64_t *__restrict__ idx, int M,
                                     float *__restrict__ out) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (m >= M) return;
    int64_t s = idx[(size_t)b * M + m];
    for (int c = 0; c < C; ++c) out[((size_t)b * C + c) * M + m] = in[((size_t)b * C + c) * N + s];
}

int launch_gather_points(const float *in, int B, int C, int N, const int64_t *idx, int M, float *out, cudaStream_t st) {
    if (B <= 0 || M <= 0) return DSIR_OK;
    dim3 grid(cdiv(M, 256), B);
    gather_points_kernel<<<grid, 256, 0, st>>>(in, C, N, idx, M, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
