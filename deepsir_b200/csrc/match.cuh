#pragma once
#include "common.cuh"

namespace dsir {

struct MatchParams {
    dsir_feat fs, fr;
    int B, C, J, K;
    const float *ns, *nr;  // squared norms [B,J], [B,K] (sequential fma chains over c)
    // argmin outputs
    int64_t *idx;
    float *min_d;
    // dense output
    float *dense;
    int metric;
    // soft
    const float *beta, *alpha, *col_bias, *xyz_ref;
    float *y_soft, *lse;
    // the caller guarantees that features and workspace are those of its previous call (iterations 2.. of the alignment
    // loop): norms, maxima and the tensor-core operand copies in the workspace are still valid and are not recomputed
    int reuse_prep;
    // optional [B,J] correspondences of the previous iteration (argmin only): a hint that tightens the filter, never changes the result
    const int64_t *prior_idx;
    // optional [B] flags: when set, only batch elements with a non-zero flag are computed (CTAs of the others leave at once)
    const unsigned char *only_flagged;
};

enum { MATCH_MODE_ARGMIN = 0, MATCH_MODE_DENSE = 1, MATCH_MODE_SOFT = 2 };

int launch_sqnorm(dsir_feat f, int B, int C, int N, float *out, cudaStream_t st);
int launch_sqnorm(dsir_feat f, int B, int C, int N, float *out, int *max_a, int *max_b, cudaStream_t st);
// min_a [B]: per-batch minimum of the norms (atomicMin on the float bits; initialise with bytes 0x7f)
int launch_sqnorm(dsir_feat f, int B, int C, int N, float *out, int *max_a, int *max_b, int *min_a, cudaStream_t st);
int launch_match_fp32(const MatchParams &P, int mode, cudaStream_t st);
int launch_row_topk(const float *dist, int B, int Jc, int K, const float *beta, const float *alpha, const float *bias, const float *lse,
                    long long lse_bs, int j0, int topk, int64_t *out_idx, float *out_w, long long out_bs, cudaStream_t st);

// the exact fp32 distance every path agrees on:  ((-2*dot) + ns) + nr,  dot = fma chain over c ascending
__device__ __forceinline__ float l2_from_dot(float dot, float ns, float nr) {
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), ns), nr);
}

}  // namespace dsir
