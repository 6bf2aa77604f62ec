#pragma once
#include "common.cuh"

namespace dsir {

constexpr int KB_NMOM = 17;  // {S|w|, Sw, Swx(3), Swy(3), Swxy(9)}

struct KabschParams {
    dsir_points src, tgt;
    const float *w;  // nullable -> unit weights
    long long w_bs;
    const int64_t *gather;  // nullable, [B,M]
    int B, M;
    int n_tgt;  // > 0: gather indices are checked against [0, n_tgt): a bad index poisons the pair (NaN moments -> status 1)
    double *partials;  // [B][nblk][17]
};

int kabsch_num_blocks(int M);
int launch_kabsch_moments(const KabschParams &P, int nblk, cudaStream_t st);
int launch_kabsch_reduce(const double *partials, int nblk, int B, double *out, cudaStream_t st);
int launch_kabsch_solve(const double *partials, int nblk, int B, float *T, int32_t *status, double *moments_out,
                        const float *compose_with, float *composed, int signed_norm, cudaStream_t st);
int launch_se3_apply(const float *T, long long T_bs, dsir_points pts, int B, int N, float *out, long long o_bs,
                     long long o_ps, long long o_cs, int rotate_only, cudaStream_t st);
int launch_se3_compose(const float *a, long long a_bs, const float *b, long long b_bs, int B, float *out, cudaStream_t st);
int launch_se3_inverse(const float *T, long long T_bs, int B, float *out, cudaStream_t st);
int launch_soft_targets(const float *W, long long w_bs, long long w_rs, const float *tgt, int B, int M, int N, float *y, float *mass,
                        cudaStream_t st);
int launch_gather_points(const float *in, int B, int C, int N, const int64_t *idx, int M, float *out, cudaStream_t st);

}  // namespace dsir
