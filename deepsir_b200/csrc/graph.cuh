#pragma once
#include "common.cuh"

namespace dsir {

int launch_gather_neighbours(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out, cudaStream_t st);
int launch_rel_pos_encoding(const float *xyz, int B, int N, const int64_t *idx, int k, float *out, cudaStream_t st);
int launch_pool_max(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out, cudaStream_t st);
int launch_sinkhorn(const float *log_alpha, int B, int J, int K, int n_iters, int slack, float *out, float *u, float *v,
                    cudaStream_t st);

int launch_log_ot(const float *scores, int B, int M, int N, const float *alpha, int iters, float *out, float *u, float *v, cudaStream_t st);

}  // namespace dsir
