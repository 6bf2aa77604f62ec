// xyz k-nearest neighbours, brute-force variant: one query per thread, support points packed as float4
// and streamed through shared memory by the TMA engine (1-D bulk copies completing on mbarriers,
// double-buffered), per-thread sorted top-k kept in registers.
//
// Replaces torch_points_kernels.knn as called at dataloader/data_base.py:165,170 of the reference.
// Distance and tie rule (identical to oracle/knn_oracle.c):
//     d2 = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32,  order = (d2, support index) lexicographic.
#include "common.cuh"
#include "knn.cuh"

namespace dsir {

__global__ void pack_xyz4_kernel(const float *__restrict__ pts, int stride, long long total, float4 *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        const float *p = pts + i * stride;
        out[i] = make_float4(p[0], p[1], p[2], 0.f);
    }
}

// in-register sorted insertion; ascending; `d` is known to be < bd[KMAX-1]
template <int KMAX>
__device__ __forceinline__ void topk_insert(float (&bd)[KMAX], int (&bi)[KMAX], float d, int s) {
#pragma unroll
    for (int p = KMAX - 1; p >= 0; --p) {
        bool shift = (p > 0) && (d < bd[p > 0 ? p - 1 : 0]);
        bool here = !shift && (d < bd[p]);
        float nd = shift ? bd[p > 0 ? p - 1 : 0] : (here ? d : bd[p]);
        int ni = shift ? bi[p > 0 ? p - 1 : 0] : (here ? s : bi[p]);
        bd[p] = nd;
        bi[p] = ni;
    }
}

template <int KMAX>
__global__ __launch_bounds__(KNN_THREADS) void knn_brute_kernel(KnnBruteParams P) {
    __shared__ __align__(128) float4 tile[2][KNN_TILE];
    __shared__ __align__(8) uint64_t bar[2];

    const int b = blockIdx.y;
    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const float4 *S = P.sup4 + (size_t)b * P.sup_bs;
    const bool active = q < P.Nq;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *qp = P.query + (size_t)b * P.qry_bs + (size_t)q * P.qry_stride;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
    }
    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int p = 0; p < KMAX; ++p) { bd[p] = INFINITY; bi[p] = -1; }

    const int ntiles = (P.Ns + KNN_TILE - 1) / KNN_TILE;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bytes = (uint32_t)min(KNN_TILE, P.Ns) * 16u;
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(tile[0], S, bytes, &bar[0]);
    }
    for (int t = 0; t < ntiles; ++t) {
        if (threadIdx.x == 0 && t + 1 < ntiles) {
            int nb = (t + 1) & 1;
            uint32_t bytes = (uint32_t)min(KNN_TILE, P.Ns - (t + 1) * KNN_TILE) * 16u;
            mbar_expect_tx(&bar[nb], bytes);
            bulk_g2s(tile[nb], S + (size_t)(t + 1) * KNN_TILE, bytes, &bar[nb]);
        }
        mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) & 1));
        const int cnt = min(KNN_TILE, P.Ns - t * KNN_TILE);
        const float4 *T = tile[t & 1];
        const int base = t * KNN_TILE;
#pragma unroll 4
        for (int i = 0; i < cnt; ++i) {
            float4 s = T[i];  // warp-uniform address: broadcast LDS.128
            float dx = __fsub_rn(qx, s.x), dy = __fsub_rn(qy, s.y), dz = __fsub_rn(qz, s.z);
            float d = __fmul_rn(dx, dx);
            d = __fmaf_rn(dy, dy, d);
            d = __fmaf_rn(dz, dz, d);
            if (d < bd[KMAX - 1]) topk_insert<KMAX>(bd, bi, d, base + i);
        }
        __syncthreads();  // everyone is done with tile[t&1] before it is refilled at iteration t+1
    }
    if (active) {
        int64_t *o = P.idx + (size_t)b * P.idx_bs + (size_t)q * P.k;
#pragma unroll
        for (int p = 0; p < KMAX; ++p)
            if (p < P.k) o[p] = (int64_t)bi[p];
        if (P.idx2 != nullptr && q < P.idx2_rows) {
            int64_t *o2 = P.idx2 + (size_t)b * P.idx2_bs + (size_t)q * P.k;
#pragma unroll
            for (int p = 0; p < KMAX; ++p)
                if (p < P.k) o2[p] = (int64_t)bi[p];
        }
        if (P.dist2 != nullptr) {
            float *od = P.dist2 + (size_t)b * P.idx_bs + (size_t)q * P.k;
#pragma unroll
            for (int p = 0; p < KMAX; ++p)
                if (p < P.k) od[p] = bd[p];
        }
    }
}

int launch_pack_xyz4(const float *pts, int stride, long long total, float4 *out, cudaStream_t st) {
    if (total <= 0) return DSIR_OK;
    pack_xyz4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pts, stride, total, out);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_knn_brute(const KnnBruteParams &P, int B, cudaStream_t st) {
    if (P.k < 1 || P.k > 32) return DSIR_ERR_UNSUPPORTED;
    if (P.Ns < P.k) return DSIR_ERR_KNN_TOO_FEW;
    if (P.Nq <= 0 || B <= 0) return DSIR_OK;
    dim3 grid((P.Nq + KNN_THREADS - 1) / KNN_THREADS, B);
    if (P.k == 1) knn_brute_kernel<1><<<grid, KNN_THREADS, 0, st>>>(P);
    else if (P.k <= 4) knn_brute_kernel<4><<<grid, KNN_THREADS, 0, st>>>(P);
    else if (P.k <= 8) knn_brute_kernel<8><<<grid, KNN_THREADS, 0, st>>>(P);
    else if (P.k <= 16) knn_brute_kernel<16><<<grid, KNN_THREADS, 0, st>>>(P);
    else knn_brute_kernel<32><<<grid, KNN_THREADS, 0, st>>>(P);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

// concatenated level clouds: xyz_cat[b, off_l + i, :] = pts[b, i, :3]  (every level is a prefix)
__global__ void pyramid_xyz_kernel(const float *__restrict__ pts, int stride, int B, int N, PyramidLevels lv,
                                   float *__restrict__ xyz_cat) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * lv.sumN;
    if (t >= total) return;
    int b = (int)(t / lv.sumN);
    int r = (int)(t % lv.sumN);
    int i = r;
#pragma unroll
    for (int l = 0; l < DSIR_MAX_LEVELS; ++l)
        if (l < lv.L && r >= lv.off[l]) i = r - lv.off[l];
    const float *p = pts + ((size_t)b * N + i) * stride;
    float *o = xyz_cat + (size_t)t * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
}

int launch_pyramid_xyz(const float *pts, int stride, int B, int N, const PyramidLevels &lv, float *xyz_cat,
                       cudaStream_t st) {
    long long total = (long long)B * lv.sumN;
    pyramid_xyz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pts, stride, B, N, lv, xyz_cat);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
