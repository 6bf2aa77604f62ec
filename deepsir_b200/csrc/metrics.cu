// On-device evaluation of a registration result: what the reference does on the host after every forward
// pass, kept on the GPU so that the evaluation loop has no device->host round trip per batch.
//
//   pose errors            common/metrics_util.py:55-63 (isotropic residual rotation / translation of inverse(gt) o pred)
//                          and test.py rte_rre (:27-33 of metrics_util.py: |t_pred - t_gt|, angle of R_pred^T R_gt)
//   correspondence check   network/loss.py:723-749 (find_correct_correspondence: np.isin over _hash keys, loss.py:280-294)
//   one-sided chamfer      common/metrics_util.py:38-40,72-74 (min_k |a_j - b_k|^2 by direct differences, mean over j)
#include "metrics.cuh"

namespace dsir {

namespace {

constexpr float RAD2DEG = 57.29577951308232f;

// metrics_util.py:55-63 in fp32 like the reference: concatenated = inverse(gt) o pred (se3_torch.py:10-48)
__global__ void pose_errors_kernel(const float *__restrict__ Tp, const float *__restrict__ Tg, int B, float rte_thresh,
                                   float rre_thresh, float eps, float *__restrict__ out /* [B,4] */, int *__restrict__ success) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *p = Tp + (size_t)b * 12, *g = Tg + (size_t)b * 12;
    // inverse(gt) = [Rg^T | -Rg^T tg];  concatenate(a, b) = [Ra Rb | Ra tb + ta]
    float tr = 0.f, t[3];
    for (int i = 0; i < 3; ++i) {
        float dii = 0.f, ti = 0.f, ci = 0.f;
        for (int k = 0; k < 3; ++k) {
            dii = __fmaf_rn(g[k * 4 + i], p[k * 4 + i], dii);   // (Rg^T Rp)_ii
            ti = __fmaf_rn(g[k * 4 + i], p[k * 4 + 3], ti);     // (Rg^T tp)_i
            ci = __fmaf_rn(g[k * 4 + i], g[k * 4 + 3], ci);     // (Rg^T tg)_i
        }
        tr += dii;
        t[i] = ti - ci;
    }
    const float c = fminf(fmaxf(0.5f * (tr - 1.f), -1.f + eps), 1.f - eps);
    const float err_r = acosf(c) * RAD2DEG;
    const float err_t = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    // rte_rre (numpy fp64 in the reference): |tp - tg|, acos((trace(Rp^T Rg) - 1) / 2)
    double dt = 0.0, tr2 = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double d = (double)p[i * 4 + 3] - (double)g[i * 4 + 3];
        dt += d * d;
        for (int k = 0; k < 3; ++k) tr2 += (double)p[k * 4 + i] * (double)g[k * 4 + i];
    }
    const double c2 = fmin(fmax((tr2 - 1.0) / 2.0, -1.0 + 1e-16), 1.0 - 1e-16);
    const double rre = acos(c2) * 180.0 / 3.141592653589793;
    const double rte = sqrt(dt);
    out[b * 4 + 0] = err_r;
    out[b * 4 + 1] = err_t;
    out[b * 4 + 2] = (float)rre;
    out[b * 4 + 3] = (float)rte;
    if (success) success[b] = (err_t < rte_thresh && err_r < rre_thresh) ? 1 : 0;
}

// ---- set membership of pair keys (loss.py:280-294, 723-749) ----
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// key = a0 + a1 * M exactly like _hash (collisions of the reference included); the table stores key * B + b
__global__ void pairs_insert_kernel(const int32_t *__restrict__ pos, const int64_t *__restrict__ offsets, int B,
                                    const int64_t *__restrict__ seeds, long long *__restrict__ table, unsigned long long tmask) {
    const long long total = offsets[B];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = B;   // batch element of pair i: offsets[b] <= i < offsets[b+1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (offsets[mid] <= i) lo = mid; else hi = mid;
        }
        const long long key = (long long)pos[2 * i] + (long long)pos[2 * i + 1] * seeds[lo];
        const long long h = key * B + lo;
        unsigned long long slot = mix64((unsigned long long)h) & tmask;
        while (true) {
            const long long prev = (long long)atomicCAS((unsigned long long *)&table[slot], ~0ull, (unsigned long long)h);
            if (prev == -1 || prev == h) break;
            slot = (slot + 1) & tmask;
        }
    }
}

__global__ void pairs_lookup_kernel(const int32_t *__restrict__ pred, int B, int N, const int64_t *__restrict__ seeds,
                                    const long long *__restrict__ table, unsigned long long tmask, uint8_t *__restrict__ correct) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int b = (int)(t / N);
    const long long key = (long long)pred[2 * t] + (long long)pred[2 * t + 1] * seeds[b];
    const long long h = key * B + b;
    unsigned long long slot = mix64((unsigned long long)h) & tmask;
    bool found = false;
    while (true) {
        const long long v = table[slot];
        if (v == h) { found = true; break; }
        if (v == -1) break;
        slot = (slot + 1) & tmask;
    }
    correct[t] = found ? 1 : 0;
}

// ---- min_k |a_j - b_k|^2 by direct differences (metrics_util.py:38-40), one thread per j, b staged through smem ----
constexpr int CH_TILE = 1024;
__global__ __launch_bounds__(256) void nn_sqdist_kernel(const float *__restrict__ a, const float *__restrict__ bpts, int N, int M,
                                                        float *__restrict__ min_d, double *__restrict__ sum /* [B] */) {
    __shared__ float sx[CH_TILE], sy[CH_TILE], sz[CH_TILE];
    __shared__ double red[8];
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const float *A = a + (size_t)b * N * 3, *Bp = bpts + (size_t)b * M * 3;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (j < N) { ax = A[3 * j]; ay = A[3 * j + 1]; az = A[3 * j + 2]; }
    float best = INFINITY;
    for (int m0 = 0; m0 < M; m0 += CH_TILE) {
        const int cnt = min(CH_TILE, M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            sx[i] = Bp[3 * (size_t)(m0 + i)]; sy[i] = Bp[3 * (size_t)(m0 + i) + 1]; sz[i] = Bp[3 * (size_t)(m0 + i) + 2];
        }
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            const float dx = __fsub_rn(ax, sx[i]), dy = __fsub_rn(ay, sy[i]), dz = __fsub_rn(az, sz[i]);
            // torch.sum(d ** 2, dim=-1): x^2 + y^2 + z^2 left to right
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            best = (d < best || d != d) ? d : best;
        }
    }
    if (j < N && min_d) min_d[(size_t)b * N + j] = best;
    double s = j < N ? (double)best : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        atomicAdd(&sum[b], tot);
    }
}

__global__ void mean_finalize_kernel(const double *__restrict__ sum, int B, int N, float *__restrict__ mean) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) mean[b] = (float)(sum[b] / (double)N);
}

}  // namespace

int launch_pose_errors(const float *Tp, const float *Tg, int B, float rte_thresh, float rre_thresh, float *out, int *success,
                       cudaStream_t st) {
    pose_errors_kernel<<<cdiv(B, 128), 128, 0, st>>>(Tp, Tg, B, rte_thresh, rre_thresh, 1e-16f, out, success);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

static size_t table_slots(long long total_pos) {
    size_t s = 1024;
    while (s < (size_t)total_pos * 2) s <<= 1;
    return s;
}

size_t correspondence_check_workspace_bytes(long long total_pos) { return ws_block(table_slots(total_pos) * sizeof(long long)) + 256; }

int launch_correspondence_check(const int32_t *pos, const int64_t *offsets, long long total_pos, const int32_t *pred, int B, int N,
                                const int64_t *seeds, uint8_t *correct, void *ws, size_t ws_bytes, cudaStream_t st) {
    Workspace W(ws, ws_bytes);
    const size_t slots = table_slots(total_pos);
    long long *table = W.take<long long>(slots);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    DSIR_CUDA_TRY(cudaMemsetAsync(table, 0xff, slots * sizeof(long long), st));
    if (total_pos > 0) {
        int blocks = (int)((total_pos + 255) / 256);
        blocks = blocks > 2048 ? 2048 : blocks;
        pairs_insert_kernel<<<blocks, 256, 0, st>>>(pos, offsets, B, seeds, table, (unsigned long long)slots - 1);
        DSIR_LAUNCH_CHECK();
    }
    const long long t = (long long)B * N;
    pairs_lookup_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(pred, B, N, seeds, table, (unsigned long long)slots - 1, correct);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

size_t nn_sqdist_workspace_bytes(int B) { return ws_block((size_t)B * sizeof(double)) + 256; }

int launch_nn_sqdist_mean(const float *a, const float *b, int B, int N, int M, float *min_d, float *mean, void *ws, size_t ws_bytes,
                          cudaStream_t st) {
    Workspace W(ws, ws_bytes);
    double *sum = W.take<double>((size_t)B);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    DSIR_CUDA_TRY(cudaMemsetAsync(sum, 0, (size_t)B * sizeof(double), st));
    nn_sqdist_kernel<<<dim3(cdiv(N, 256), B), 256, 0, st>>>(a, b, N, M, min_d, sum);
    DSIR_LAUNCH_CHECK();
    if (mean) {
        mean_finalize_kernel<<<cdiv(B, 128), 128, 0, st>>>(sum, B, N, mean);
        DSIR_LAUNCH_CHECK();
    }
    return DSIR_OK;
}

}  // namespace dsir
