// Feature-distance contraction on the fp32 CUDA cores, with three fused epilogues:
//   ARGMIN : row-wise first-index minimum          (network/model.py:558-569 of the reference)
//   DENSE  : materialised [B,J,K] matrix            (network/matchnet.py:49-192)
//   SOFT   : affinity + online row softmax + soft target, [J,K] never written
//            (network/matchnet.py:195-208,259; network/model.py:81-84)
// This kernel DEFINES the fp32 value every other path must reproduce:
//     dot_jk = fma(s_{C-1}, r_{C-1}, ... fma(s_0, r_0, 0))   (channels ascending)
//     d_jk   = ((-2 * dot_jk) + |s_j|^2) + |r_k|^2            (op order of matchnet.py:110-112)
// It is the refine/fallback stage of the tcgen05 path (match_tc.cu) and the whole path for the soft
// variant, whose 1e-4 relative tolerance rules out tf32 inputs.
#include "match.cuh"

namespace dsir {

constexpr int TM = 128, TN = 128, CK = 16, MT = 256;
constexpr int PITCH = TM + 4;

// squared norms (fma chain over channels ascending) and, optionally, the per-batch maximum (atomicMax on the float bits
// of the non-negative norms; a NaN compares above every finite value) into max_a[b] and max_b[b]
__global__ void sqnorm_kernel(dsir_feat f, int C, int N, float *__restrict__ out, int *__restrict__ max_a, int *__restrict__ max_b,
                              int *__restrict__ min_a) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    float acc = 0.f;
    if (n < N) {
        const float *p = f.ptr + (size_t)b * f.batch_stride + (size_t)n * f.point_stride;
        for (int c = 0; c < C; ++c) {
            float v = p[(size_t)c * f.chan_stride];
            acc = __fmaf_rn(v, v, acc);
        }
        out[(size_t)b * N + n] = acc;
    }
    if (max_a) {
        int bits = __float_as_int(acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
        if ((threadIdx.x & 31) == 0) {
            atomicMax(&max_a[b], bits);
            if (max_b) atomicMax(&max_b[b], bits);
        }
    }
    if (min_a) {   // per-batch minimum (lanes beyond N do not take part)
        int bits = n < N ? __float_as_int(acc) : 0x7f7f7f7f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bits = min(bits, __shfl_xor_sync(0xffffffffu, bits, o));
        if ((threadIdx.x & 31) == 0) atomicMin(&min_a[b], bits);
    }
}

// channel-major features with unit point stride and 16-byte aligned rows: four points per thread, so that a warp reads 512
// contiguous bytes of every channel row (the scalar kernel reads 128) - the pass is a pure HBM stream of the features
__global__ void sqnorm4_kernel(dsir_feat f, int C, int N, float *__restrict__ out, int *__restrict__ max_a, int *__restrict__ max_b,
                               int *__restrict__ min_a) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int b = blockIdx.y;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        const float *p = f.ptr + (size_t)b * f.batch_stride + n;
#pragma unroll 8
        for (int c = 0; c < C; ++c) {
            const float4 v = *reinterpret_cast<const float4 *>(p + (size_t)c * f.chan_stride);
            acc.x = __fmaf_rn(v.x, v.x, acc.x); acc.y = __fmaf_rn(v.y, v.y, acc.y);
            acc.z = __fmaf_rn(v.z, v.z, acc.z); acc.w = __fmaf_rn(v.w, v.w, acc.w);
        }
        *reinterpret_cast<float4 *>(out + (size_t)b * N + n) = acc;
    }
    if (max_a || min_a) {
        const bool v = n < N;
        int hi = max(max(__float_as_int(acc.x), __float_as_int(acc.y)), max(__float_as_int(acc.z), __float_as_int(acc.w)));
        int lo = v ? min(min(__float_as_int(acc.x), __float_as_int(acc.y)), min(__float_as_int(acc.z), __float_as_int(acc.w))) : 0x7f7f7f7f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (max_a) atomicMax(&max_a[b], hi);
            if (max_b) atomicMax(&max_b[b], hi);
            if (min_a) atomicMin(&min_a[b], lo);
        }
    }
}

int launch_sqnorm(dsir_feat f, int B, int C, int N, float *out, int *max_a, int *max_b, int *min_a, cudaStream_t st) {
    const bool vec = f.point_stride == 1 && N % 4 == 0 && f.batch_stride % 4 == 0 && f.chan_stride % 4 == 0 &&
                     ((uintptr_t)f.ptr & 15) == 0 && ((uintptr_t)out & 15) == 0;
    if (vec) {
        dim3 grid4(cdiv(N, 4 * 128), B);
        sqnorm4_kernel<<<grid4, 128, 0, st>>>(f, C, N, out, max_a, max_b, min_a);
        DSIR_LAUNCH_CHECK();
        return DSIR_OK;
    }
    dim3 grid(cdiv(N, 256), B);
    sqnorm_kernel<<<grid, 256, 0, st>>>(f, C, N, out, max_a, max_b, min_a);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}
int launch_sqnorm(dsir_feat f, int B, int C, int N, float *out, int *max_a, int *max_b, cudaStream_t st) {
    return launch_sqnorm(f, B, C, N, out, max_a, max_b, nullptr, st);
}
int launch_sqnorm(dsir_feat f, int B, int C, int N, float *out, cudaStream_t st) {
    return launch_sqnorm(f, B, C, N, out, nullptr, nullptr, st);
}

// INNER: 0 = dot product, 1 = sum of squared differences, 2 = sum of absolute differences
template <int MODE, int INNER>
__global__ __launch_bounds__(MT, 2) void match_fp32_kernel(MatchParams P) {
    __shared__ __align__(16) float As[CK][PITCH];
    __shared__ __align__(16) float Bs[CK][PITCH];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y;
    if (P.only_flagged != nullptr && P.only_flagged[b] == 0) return;   // (block-uniform, before any barrier)
    const int j0 = blockIdx.x * TM;
    const float *fsb = P.fs.ptr + (size_t)b * P.fs.batch_stride;
    const float *frb = P.fr.ptr + (size_t)b * P.fr.batch_stride;
    const bool src_cn = P.fs.point_stride == 1;
    const bool ref_cn = P.fr.point_stride == 1;

    int rowi[8];
    float nsv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        rowi[i] = j0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        nsv[i] = (MODE != MATCH_MODE_DENSE || INNER == 0) && rowi[i] < P.J && P.ns ? P.ns[(size_t)b * P.J + rowi[i]] : 0.f;
    }

    // per-row running state
    float best_d[8];
    int best_k[8];
    float sm[8], sl[8], sx[8], sy[8], sz[8];
    float beta = 0.f, alpha = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        best_d[i] = INFINITY; best_k[i] = 0;
        sm[i] = -INFINITY; sl[i] = 0.f; sx[i] = 0.f; sy[i] = 0.f; sz[i] = 0.f;
    }
    if (MODE == MATCH_MODE_SOFT) { beta = P.beta[b]; alpha = P.alpha[b]; }

    for (int k0 = 0; k0 < P.K; k0 += TN) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        for (int c0 = 0; c0 < P.C; c0 += CK) {
            // ---- stage a [CK x TM] slab of source and a [CK x TN] slab of reference features ----
#pragma unroll
            for (int e = tid; e < CK * TM; e += MT) {
                int c, n;
                if (src_cn) { c = e / TM; n = e % TM; } else { n = e / CK; c = e % CK; }
                float v = 0.f;
                if (c0 + c < P.C && j0 + n < P.J)
                    v = fsb[(size_t)(c0 + c) * P.fs.chan_stride + (size_t)(j0 + n) * P.fs.point_stride];
                As[c][n] = v;
            }
#pragma unroll
            for (int e = tid; e < CK * TN; e += MT) {
                int c, n;
                if (ref_cn) { c = e / TN; n = e % TN; } else { n = e / CK; c = e % CK; }
                float v = 0.f;
                if (c0 + c < P.C && k0 + n < P.K)
                    v = frb[(size_t)(c0 + c) * P.fr.chan_stride + (size_t)(k0 + n) * P.fr.point_stride];
                Bs[c][n] = v;
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                float4 a0 = *reinterpret_cast<const float4 *>(&As[c][ty * 4]);
                float4 a1 = *reinterpret_cast<const float4 *>(&As[c][64 + ty * 4]);
                float4 b0 = *reinterpret_cast<const float4 *>(&Bs[c][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4 *>(&Bs[c][64 + tx * 4]);
                float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (INNER == 0) {
                            acc[i][j] = __fmaf_rn(a[i], bb[j], acc[i][j]);
                        } else if (INNER == 1) {
                            float df = __fsub_rn(a[i], bb[j]);
                            acc[i][j] = __fmaf_rn(df, df, acc[i][j]);
                        } else {
                            acc[i][j] = __fadd_rn(acc[i][j], fabsf(__fsub_rn(a[i], bb[j])));
                        }
                    }
            }
            __syncthreads();
        }

        // ---- fused epilogue on the 8x8 register tile ----
        int colk[8];
        float nrv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            colk[j] = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            nrv[j] = (INNER == 0 && P.nr && colk[j] < P.K) ? P.nr[(size_t)b * P.K + colk[j]] : 0.f;
        }
        if (MODE == MATCH_MODE_ARGMIN) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float d = l2_from_dot(acc[i][j], nsv[i], nrv[j]);
                    if (colk[j] < P.K && d < best_d[i]) { best_d[i] = d; best_k[i] = colk[j]; }
                }
        } else if (MODE == MATCH_MODE_DENSE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (rowi[i] >= P.J) continue;
                float *o = P.dense + ((size_t)b * P.J + rowi[i]) * P.K;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (colk[j] >= P.K) continue;
                    float v;
                    if (INNER != 0) v = (P.metric == DSIR_METRIC_SQDIFF_SQRT) ? sqrtf(__fadd_rn(acc[i][j], 1e-16f)) : acc[i][j];
                    else if (P.metric == DSIR_METRIC_ACOS_DOT) v = acosf(acc[i][j]);
                    else {
                        v = l2_from_dot(acc[i][j], nsv[i], nrv[j]);
                        if (P.metric == DSIR_METRIC_EUCLIDEAN) v = sqrtf(__fadd_rn(v, 1e-16f));
                    }
                    o[colk[j]] = v;
                }
            }
        } else {  // SOFT: online softmax over this thread's 8 columns of each row
            float bias[8], rx[8], ry[8], rz[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bool ok = colk[j] < P.K;
                bias[j] = (ok && P.col_bias) ? P.col_bias[(size_t)b * P.K + colk[j]] : 0.f;
                const float *r = P.xyz_ref + ((size_t)b * P.K + (ok ? colk[j] : 0)) * 3;
                rx[j] = (ok && P.y_soft) ? r[0] : 0.f;
                ry[j] = (ok && P.y_soft) ? r[1] : 0.f;
                rz[j] = (ok && P.y_soft) ? r[2] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float lg[8];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float d = l2_from_dot(acc[i][j], nsv[i], nrv[j]);
                    float a = __fadd_rn(__fmul_rn(-beta, __fsub_rn(d, alpha)), bias[j]);
                    lg[j] = colk[j] < P.K ? a : -INFINITY;
                    mx = fmaxf(mx, lg[j]);
                }
                float mnew = fmaxf(sm[i], mx);
                if (mnew == -INFINITY) continue;
                float sc = __expf(sm[i] - mnew);
                float l = sl[i] * sc, x = sx[i] * sc, y = sy[i] * sc, z = sz[i] * sc;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float p = __expf(lg[j] - mnew);
                    l += p;
                    x = __fmaf_rn(p, rx[j], x);
                    y = __fmaf_rn(p, ry[j], y);
                    z = __fmaf_rn(p, rz[j], z);
                }
                sm[i] = mnew; sl[i] = l; sx[i] = x; sy[i] = y; sz[i] = z;
            }
        }
    }

    // ---- combine the 16 threads (tx) that share each row; they are 16 consecutive lanes ----
    if (MODE == MATCH_MODE_ARGMIN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float d = best_d[i];
            int k = best_k[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                float d2 = __shfl_xor_sync(0xffffffffu, d, o);
                int k2 = __shfl_xor_sync(0xffffffffu, k, o);
                if (d2 < d || (d2 == d && k2 < k)) { d = d2; k = k2; }
            }
            if (tx == 0 && rowi[i] < P.J) {
                P.idx[(size_t)b * P.J + rowi[i]] = (int64_t)k;
                if (P.min_d) P.min_d[(size_t)b * P.J + rowi[i]] = d;
            }
        }
    } else if (MODE == MATCH_MODE_SOFT) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float m = sm[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float sc = (sm[i] == -INFINITY) ? 0.f : __expf(sm[i] - m);
            float l = sl[i] * sc, x = sx[i] * sc, y = sy[i] * sc, z = sz[i] * sc;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                l += __shfl_xor_sync(0xffffffffu, l, o);
                x += __shfl_xor_sync(0xffffffffu, x, o);
                y += __shfl_xor_sync(0xffffffffu, y, o);
                z += __shfl_xor_sync(0xffffffffu, z, o);
            }
            if (tx == 0 && rowi[i] < P.J) {
                size_t r = (size_t)b * P.J + rowi[i];
                if (P.lse) P.lse[r] = m + logf(l);
                if (P.y_soft) {
                    // weights w = p/l sum to s = 1 (up to rounding): y = (sum w r) / (s + 1e-16)
                    float inv = 1.f / l;
                    P.y_soft[r * 3 + 0] = x * inv;
                    P.y_soft[r * 3 + 1] = y * inv;
                    P.y_soft[r * 3 + 2] = z * inv;
                }
            }
        }
    }
}

int launch_match_fp32(const MatchParams &P, int mode, cudaStream_t st) {
    if (P.B <= 0 || P.J <= 0 || P.K <= 0 || P.C <= 0) return DSIR_ERR_BAD_ARG;
    dim3 grid(cdiv(P.J, TM), P.B);
    if (mode == MATCH_MODE_ARGMIN) match_fp32_kernel<MATCH_MODE_ARGMIN, 0><<<grid, MT, 0, st>>>(P);
    else if (mode == MATCH_MODE_SOFT) match_fp32_kernel<MATCH_MODE_SOFT, 0><<<grid, MT, 0, st>>>(P);
    else if (mode == MATCH_MODE_DENSE) {
        if (P.metric == DSIR_METRIC_SQDIFF || P.metric == DSIR_METRIC_SQDIFF_SQRT) match_fp32_kernel<MATCH_MODE_DENSE, 1><<<grid, MT, 0, st>>>(P);
        else if (P.metric == DSIR_METRIC_CITYBLOCK) match_fp32_kernel<MATCH_MODE_DENSE, 2><<<grid, MT, 0, st>>>(P);
        else match_fp32_kernel<MATCH_MODE_DENSE, 0><<<grid, MT, 0, st>>>(P);
    } else return DSIR_ERR_BAD_ARG;
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// top-k soft correspondences of every row from a materialised distance chunk: a_jk = -beta (d_jk - alpha) (+ bias_k),
// the k largest a_jk (= largest weights), descending, ties to the lower index; w = exp(a - lse_j).
// One warp per row: every lane keeps the k best of its strided columns in a sorted register list, then k rounds of a
// warp arg-max over the lane heads merge the 32 lists.
// ---------------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ __launch_bounds__(256) void row_topk_kernel(const float *__restrict__ dist, int B, int Jc, int K, const float *__restrict__ beta,
                                                       const float *__restrict__ alpha, const float *__restrict__ bias, const float *__restrict__ lse,
                                                       long long lse_bs, int j0, int topk, int64_t *__restrict__ out_idx, float *__restrict__ out_w,
                                                       long long out_bs) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= (long long)B * Jc) return;
    const int b = (int)(row / Jc), jr = (int)(row % Jc);
    const float *d = dist + (size_t)row * K;
    const float nb = -beta[b], al = alpha[b];
    float bv[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int p = 0; p < KMAX; ++p) { bv[p] = -INFINITY; bi[p] = 0x7fffffff; }
    for (int k = lane; k < K; k += 32) {
        float a = nb * (d[k] - al);
        if (bias) a += bias[(size_t)b * K + k];
        if (a > bv[KMAX - 1] || (a == bv[KMAX - 1] && k < bi[KMAX - 1])) {   // sorted insertion, descending (value, -index)
#pragma unroll
            for (int p = KMAX - 1; p >= 0; --p) {
                const int pm = p > 0 ? p - 1 : 0;
                const bool shift = (p > 0) && (a > bv[pm] || (a == bv[pm] && k < bi[pm]));
                const bool here = !shift && (a > bv[p] || (a == bv[p] && k < bi[p]));
                bv[p] = shift ? bv[pm] : (here ? a : bv[p]);
                bi[p] = shift ? bi[pm] : (here ? k : bi[p]);
            }
        }
    }
    const float l = lse[(size_t)b * lse_bs + j0 + jr];
    int64_t *oi = out_idx + (size_t)b * out_bs + (size_t)(j0 + jr) * topk;
    float *ow = out_w + (size_t)b * out_bs + (size_t)(j0 + jr) * topk;
    for (int t = 0; t < topk; ++t) {
        // warp arg-max over the lane heads
        float v = bv[0];
        int i = bi[0], src = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o), s2 = __shfl_xor_sync(0xffffffffu, src, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; src = s2; }
        }
        if (lane == src) {       // pop the head
#pragma unroll
            for (int p = 0; p < KMAX - 1; ++p) { bv[p] = bv[p + 1]; bi[p] = bi[p + 1]; }
            bv[KMAX - 1] = -INFINITY; bi[KMAX - 1] = 0x7fffffff;
        }
        if (lane == 0) {
            oi[t] = i == 0x7fffffff ? (int64_t)-1 : (int64_t)i;
            ow[t] = i == 0x7fffffff ? 0.f : expf(v - l);
        }
    }
}

int launch_row_topk(const float *dist, int B, int Jc, int K, const float *beta, const float *alpha, const float *bias, const float *lse,
                    long long lse_bs, int j0, int topk, int64_t *out_idx, float *out_w, long long out_bs, cudaStream_t st) {
    const long long warps = (long long)B * Jc;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    if (topk <= 8) row_topk_kernel<8><<<grid, 256, 0, st>>>(dist, B, Jc, K, beta, alpha, bias, lse, lse_bs, j0, topk, out_idx, out_w, out_bs);
    else if (topk <= 32) row_topk_kernel<32><<<grid, 256, 0, st>>>(dist, B, Jc, K, beta, alpha, bias, lse, lse_bs, j0, topk, out_idx, out_w, out_bs);
    else return DSIR_ERR_UNSUPPORTED;
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
