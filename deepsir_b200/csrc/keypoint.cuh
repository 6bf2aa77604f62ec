#pragma once
#include "common.cuh"

namespace dsir {

size_t keypoint_score_workspace_bytes(int B, int C, int N);
int launch_keypoint_score(const float *feat, const float *xyz, const float *prob, const int64_t *label, const float *lw,
                          int num_class, const int64_t *idx, int idx_stride, int k, float ball_r, int B, int C, int N,
                          float *score, void *ws, size_t ws_bytes, cudaStream_t st);
int launch_topk_rows(const float *score, int B, int N, int k, float *values, int64_t *index, cudaStream_t st);

}  // namespace dsir
