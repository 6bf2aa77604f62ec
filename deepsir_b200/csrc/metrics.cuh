#pragma once
#include "common.cuh"

namespace dsir {

int launch_pose_errors(const float *Tp, const float *Tg, int B, float rte_thresh, float rre_thresh, float *out, int *success,
                       cudaStream_t st);
size_t correspondence_check_workspace_bytes(long long total_pos);
int launch_correspondence_check(const int32_t *pos, const int64_t *offsets, long long total_pos, const int32_t *pred, int B, int N,
                                const int64_t *seeds, uint8_t *correct, void *ws, size_t ws_bytes, cudaStream_t st);
size_t nn_sqdist_workspace_bytes(int B);
int launch_nn_sqdist_mean(const float *a, const float *b, int B, int N, int M, float *min_d, float *mean, void *ws, size_t ws_bytes,
                          cudaStream_t st);

}  // namespace dsir
