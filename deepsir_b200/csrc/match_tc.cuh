#pragma once
#include "match.cuh"

namespace dsir {

// tcgen05/TMEM filter + exact fp32 refine (match_tc.cu)
bool match_tc_supported(const dsir_feat &fs, const dsir_feat &fr, int B, int C, int J, int K);
bool match_tc_profitable(int B, int C, int J, int K);
size_t match_tc_workspace_bytes(int B, int C, int J, int K);
int launch_match_tc(const MatchParams &P, void *ws, size_t ws_bytes, cudaStream_t st);
int match_tc_rescued_rows(const void *ws, int B, int C, int J, int K, int *out, cudaStream_t st);
int match_tc_filter_trace(const void *ws, int B, int C, int J, int K, unsigned int *out, cudaStream_t st);
int match_tc_filter_timing(const void *ws, int B, int C, int J, int K, double *out, cudaStream_t st);

// top-k soft correspondences on the tensor cores: two filter sweeps + exact re-scoring of <= 4k listed columns per row
// (no score matrix, no row chunks); needs beta > 0 per batch element for the fast route and no column bias
bool match_tc_topk_supported(int B, int C, int J, int K, int topk);
size_t match_tc_topk_workspace_bytes(int B, int C, int J, int K, int topk);
int launch_match_tc_topk(const MatchParams &P, int topk, int64_t *out_idx, float *out_w, void *ws, size_t ws_bytes, cudaStream_t st);
int match_tc_topk_exhaustive_rows(const void *ws, int B, int C, int J, int K, int topk, int *out, cudaStream_t st);

// tcgen05 soft match: bf16 x3 split contraction fused with the row-wise online softmax (match_tc_soft.cu), C <= 32
bool match_tc_soft_supported(int B, int C, int J, int K);
size_t match_tc_soft_workspace_bytes(int B, int C, int J, int K);
int launch_match_tc_soft(const MatchParams &P, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace dsir
