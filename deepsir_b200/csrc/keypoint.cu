// Key-point scoring and top-k selection on the KNN graph: HBM-bound gather / reduce kernels.
//
//   keypoint score   Network.score_fun  (network/model.py:700-757)   feat [B,C,N], xyz [B,3,N], prob, label, neigh_idx
//                                                                   -> score [B,N]
//   top-k rows       torch.topk(score, k, largest=True) as used by Network.feat_score (network/model.py:692)
//
// The reference gathers a [B,C,N,16] copy of the features (repeat + gather) to take the neighbour mean.  Here the
// features are transposed once to point-major [B,N,C] (already divided by the per-sample maximum, model.py:719-720), so
// that the 16 neighbour rows of a point are 16 contiguous C-float reads; one warp scores one point and nothing of size
// N x k is ever written.
//
// top-k: keys are (value descending, index ascending) packed into 64 bits, so every key is distinct and ties resolve to
// the LOWER index (torch.topk leaves the tie order unspecified; this is the stricter rule SURVEY f-1 asks for).  One CTA
// per row: 8-bit radix select of the k-th key over the packed keys, compaction of the k winners into shared memory,
// bitonic sort, write-out.
#include "keypoint.cuh"

namespace dsir {

namespace {

__device__ __forceinline__ unsigned int f2ord(float f) {   // monotone float -> uint (NaN above +inf, like torch.max/topk)
    if (f != f) return 0xffffffffu;
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
    if (o == 0xffffffffu) return __uint_as_float(0x7fc00000u);
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ---- per-sample maxima: feat (model.py:719), prob (:747), label weight (:744) ----
__global__ __launch_bounds__(256) void kp_max_kernel(const float *__restrict__ feat, long long CN, const float *__restrict__ prob,
                                                     const int64_t *__restrict__ label, const float *__restrict__ lw, int num_class,
                                                     int N, unsigned int *__restrict__ maxima /* [B][3] ordered bits */) {
    const int b = blockIdx.y;
    const float *f = feat + (size_t)b * CN;
    unsigned int mf = 0, mp = 0, ml = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < CN; i += (long long)gridDim.x * blockDim.x) {
        mf = max(mf, f2ord(f[i]));
        if (i < N) {
            if (prob) mp = max(mp, f2ord(prob[(size_t)b * N + i]));
            if (label) {
                const long long c = label[(size_t)b * N + i];
                ml = max(ml, f2ord((c >= 0 && c < num_class) ? lw[c] : 0.f));
            }
        }
    }
    mf = __reduce_max_sync(0xffffffffu, mf);
    mp = __reduce_max_sync(0xffffffffu, mp);
    ml = __reduce_max_sync(0xffffffffu, ml);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&maxima[b * 3 + 0], mf);
        atomicMax(&maxima[b * 3 + 1], mp);
        atomicMax(&maxima[b * 3 + 2], ml);
    }
}

// ---- [B,C,N] -> [B,N,C], divided by (max + eps) ----
__global__ __launch_bounds__(256) void kp_transpose_kernel(const float *__restrict__ feat, int C, int N,
                                                           const unsigned int *__restrict__ maxima, float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float den = __fadd_rn(ord2f(maxima[b * 3 + 0]), 1e-16f);
    const float *src = feat + (size_t)b * C * N;
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, n = n0 + tx;
        tile[i][tx] = (c < C && n < N) ? __fdiv_rn(src[(size_t)c * N + n], den) : 0.f;
    }
    __syncthreads();
    float *dst = out + (size_t)b * N * C;
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, c = c0 + tx;
        if (n < N && c < C) dst[(size_t)n * C + c] = tile[tx][i];
    }
}

__device__ __forceinline__ float softplus_ref(float x) {   // F.softplus, beta = 1, threshold = 20
    return x > 20.f ? x : log1pf(expf(x));
}

// ---- one warp per point ----
struct ScoreParams {
    const float *featT;   // [B,N,C] normalised
    const float *xyz;     // [B,3,N]
    const float *prob;    // [B,N] or null
    const int64_t *label; // [B,N] or null
    const float *lw;
    int num_class;
    const int64_t *idx;   // [B,N,idx_stride]
    int idx_stride, k;    // neighbours used (<= 32)
    float ball_r;
    const unsigned int *maxima;
    int B, C, N;
    float *score;         // [B,N]
};

__global__ __launch_bounds__(256) void kp_score_kernel(ScoreParams P) {
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)P.B * P.N) return;
    const int b = (int)(w / P.N), n = (int)(w % P.N);
    const float *FT = P.featT + (size_t)b * P.N * P.C;
    const float *X = P.xyz + (size_t)b * 3 * P.N;
    // neighbour j lives in lane j
    long long nj = -1;
    if (lane < P.k) nj = P.idx[((size_t)b * P.N + n) * P.idx_stride + lane];
    const bool okj = nj >= 0 && nj < P.N;
    // 2. aggregation score (model.py:728-733): mean_j |xyz_j - xyz_n| < ball_r
    const float cx = X[n], cy = X[P.N + n], cz = X[2 * (size_t)P.N + n];
    float dist = 0.f;
    if (lane < P.k) {
        const float rx = __fsub_rn(okj ? X[nj] : 0.f, cx), ry = __fsub_rn(okj ? X[P.N + nj] : 0.f, cy),
                    rz = __fsub_rn(okj ? X[2 * (size_t)P.N + nj] : 0.f, cz);
        dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, o);
    const float agg = (__fdiv_rn(dist, (float)P.k) < P.ball_r) ? 1.f : 0.f;
    // 4. semantic score (model.py:741-749)
    float ls = 1.f;
    if (P.label) {
        const long long c = P.label[(size_t)b * P.N + n];
        const float wgt = (c >= 0 && c < P.num_class) ? P.lw[c] : 0.f;
        ls = __fdiv_rn(wgt, __fadd_rn(ord2f(P.maxima[b * 3 + 2]), 1e-16f));
    }
    if (P.prob) {
        const float pn = __fdiv_rn(P.prob[(size_t)b * P.N + n], __fadd_rn(ord2f(P.maxima[b * 3 + 1]), 1e-16f));
        ls = pn > 0.2f ? ls : 0.f;
    }
    // 3. channel-wise maximum of the point (model.py:736-738)
    float dwm = -INFINITY;
    for (int c = lane; c < P.C; c += 32) dwm = fmaxf(dwm, FT[(size_t)n * P.C + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dwm = fmaxf(dwm, __shfl_xor_sync(0xffffffffu, dwm, o));
    const float dden = __fadd_rn(dwm, 1e-16f);
    // 1. saliency (model.py:722-725) and 5. product, channel maximum (model.py:752-755)
    float best = -INFINITY;
    for (int c0 = 0; c0 < P.C; c0 += 32) {
        const int c = c0 + lane;
        float sum = 0.f;
        for (int j = 0; j < P.k; ++j) {
            const long long m = __shfl_sync(0xffffffffu, nj, j);
            const float v = (c < P.C && m >= 0 && m < P.N) ? FT[(size_t)m * P.C + c] : 0.f;
            sum = __fadd_rn(sum, v);
        }
        if (c < P.C) {
            const float f = FT[(size_t)n * P.C + c];
            const float lms = softplus_ref(__fsub_rn(f, __fdiv_rn(sum, (float)P.k)));
            const float s = __fmul_rn(__fmul_rn(__fmul_rn(lms, agg), __fdiv_rn(f, dden)), ls);
            best = (s > best || s != s) ? s : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, best, o);
        best = (other > best || other != other) ? other : best;
    }
    if (lane == 0) P.score[(size_t)b * P.N + n] = best;
}

// ---- top-k of every row ----
__device__ __forceinline__ unsigned long long topk_key(float v, int i) {
    if (v == 0.f) v = 0.f;   // -0 and +0 compare equal in torch.topk: one key for both
    return ((unsigned long long)(~f2ord(v)) << 32) | (unsigned int)i;   // ascending key = descending value, ascending index
}

__global__ __launch_bounds__(1024) void topk_rows_kernel(const float *__restrict__ score, int N, int k, int kpad,
                                                         float *__restrict__ values, int64_t *__restrict__ index) {
    extern __shared__ unsigned long long skeys[];   // [kpad]
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_count;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *row = score + (size_t)b * N;
    if (tid == 0) { s_prefix = 0ull; s_remaining = k; s_count = 0; }
    unsigned long long mask = 0ull;
    __syncthreads();
    // radix select, most significant byte first: after the 8 passes s_prefix is the k-th smallest key
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i = tid; i < N; i += blockDim.x) {
            const unsigned long long key = topk_key(row[i], i);
            if ((key & mask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int rem = s_remaining, d = 0;
            for (; d < 255; ++d) {
                if ((int)hist[d] >= rem) break;
                rem -= (int)hist[d];
            }
            s_remaining = rem;
            s_prefix = prefix | ((unsigned long long)d << shift);
        }
        mask |= 0xffull << shift;
        __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    for (int i = tid; i < kpad; i += blockDim.x) skeys[i] = ~0ull;
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
        const unsigned long long key = topk_key(row[i], i);
        if (key <= kth) {
            const int pos = atomicAdd(&s_count, 1);
            if (pos < kpad) skeys[pos] = key;
        }
    }
    __syncthreads();
    // bitonic sort, ascending
    for (int size = 2; size <= kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < kpad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = skeys[lo], c = skeys[hi];
                if ((a > c) == up) { skeys[lo] = c; skeys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += blockDim.x) {
        const unsigned long long key = skeys[i];
        const int id = (int)(unsigned int)(key & 0xffffffffull);
        index[(size_t)b * k + i] = id;
        values[(size_t)b * k + i] = row[id];
    }
}

}  // namespace

size_t keypoint_score_workspace_bytes(int B, int C, int N) {
    return ws_block((size_t)B * 3 * sizeof(unsigned int)) + ws_block((size_t)B * N * C * sizeof(float)) + 256;
}

int launch_keypoint_score(const float *feat, const float *xyz, const float *prob, const int64_t *label, const float *lw,
                          int num_class, const int64_t *idx, int idx_stride, int k, float ball_r, int B, int C, int N,
                          float *score, void *ws, size_t ws_bytes, cudaStream_t st) {
    Workspace W(ws, ws_bytes);
    unsigned int *maxima = W.take<unsigned int>((size_t)B * 3);
    float *featT = W.take<float>((size_t)B * N * C);
    if (!W.ok()) return DSIR_ERR_WORKSPACE;
    DSIR_CUDA_TRY(cudaMemsetAsync(maxima, 0, (size_t)B * 3 * sizeof(unsigned int), st));
    const long long CN = (long long)C * N;
    int blocks = (int)((CN + 256 * 8 - 1) / (256 * 8));
    blocks = blocks < 1 ? 1 : (blocks > 1024 ? 1024 : blocks);
    kp_max_kernel<<<dim3(blocks, B), 256, 0, st>>>(feat, CN, prob, label, lw, num_class, N, maxima);
    DSIR_LAUNCH_CHECK();
    kp_transpose_kernel<<<dim3(cdiv(N, 32), cdiv(C, 32), B), 256, 0, st>>>(feat, C, N, maxima, featT);
    DSIR_LAUNCH_CHECK();
    ScoreParams P{featT, xyz, prob, label, lw, num_class, idx, idx_stride, k, ball_r, maxima, B, C, N, score};
    const long long warps = (long long)B * N;
    kp_score_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(P);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_topk_rows(const float *score, int B, int N, int k, float *values, int64_t *index, cudaStream_t st) {
    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    const size_t smem = (size_t)kpad * sizeof(unsigned long long);
    if (smem > 200 * 1024) return DSIR_ERR_UNSUPPORTED;
    DSIR_CUDA_TRY(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_rows_kernel<<<B, 1024, smem, st>>>(score, N, k, kpad, values, index);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
