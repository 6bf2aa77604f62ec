/*
 * deepsir_b200 — C ABI of the B200-native (sm_100a) correspondence-and-pose hot path.
 *
 * Every entry point replaces one PyTorch-level function (or inline block) of the reference
 * LeoQLi/DeepSIR; the reference file:line each one stands in for is cited next to it.  The library is
 * a drop-in for that path only: plain device pointers and sizes, caller-owned memory (inputs, outputs
 * and workspace), an explicit cudaStream_t (passed as void*), no torch types, no hidden host
 * synchronisation, no CPU fallback.  All functions return DSIR_OK (0) or a negative DSIR_ERR_* code
 * and never throw or exit.  All tensors are dense fp32 / int64 in device memory unless stated.
 *
 * How a maintainer binds this from the reference (ctypes stub): see INTEGRATION.md.
 */
#ifndef DEEPSIR_B200_H_
#define DEEPSIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *dsir_stream_t; /* a cudaStream_t; NULL = legacy default stream */

enum {
    DSIR_OK = 0,
    DSIR_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, unknown enum */
    DSIR_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels implement (e.g. k > 32) */
    DSIR_ERR_WORKSPACE = -3,    /* workspace pointer null/misaligned or too small */
    DSIR_ERR_CUDA = -4,         /* a CUDA runtime/driver call failed; see dsir_last_cuda_error() */
    DSIR_ERR_KNN_TOO_FEW = -5,  /* fewer support points than k (the reference kernel raises) */
    DSIR_ERR_NO_DEVICE = -6     /* no sm_100 device: this library has no other code path */
};

/* feature layout of the match entry points */
enum { DSIR_LAYOUT_CN = 0 /* [B,C,N], network/matchnet.py:96 */, DSIR_LAYOUT_NC = 1 /* [B,N,C], matchnet.py:49 */ };
/* metric of dsir_match_dense (network/matchnet.py:116-144) */
enum { DSIR_METRIC_L2 = 0, DSIR_METRIC_EUCLIDEAN = 1, DSIR_METRIC_ACOS_DOT = 2, DSIR_METRIC_SQDIFF = 3, DSIR_METRIC_CITYBLOCK = 4,
       DSIR_METRIC_SQDIFF_SQRT = 5 /* feat_dist 'euclidean': sqrt(sum (s-r)^2 + 1e-16) */ };
/* algorithm selectors (0 = let the library choose) */
enum { DSIR_KNN_AUTO = 0, DSIR_KNN_BRUTE = 1, DSIR_KNN_GRID = 2 /* uniform grid, any size */,
       DSIR_KNN_TREE = 3 /* bucket tree of kd-ordered leaves: clouds of <= 17408 points, k <= 16 (else the grid) */ };
enum { DSIR_MATCH_AUTO = 0, DSIR_MATCH_FP32 = 1 /* CUDA-core exact */, DSIR_MATCH_TC = 2 /* tcgen05 filter + fp32 refine */ };

int dsir_version(void);
const char *dsir_strerror(int code);
const char *dsir_last_cuda_error(void); /* text of the last CUDA error seen by this thread */
/* 0 when the current device is sm_100 (B200); DSIR_ERR_NO_DEVICE otherwise */
int dsir_device_check(void);
/* number of kernels this library has enqueued since it was loaded (all threads) */
uint64_t dsir_launch_count(void);

/* In-situ profiler (diagnostic; not for production timing): between dsir_profile_begin(stream) and
 * dsir_profile_report() every kernel launch site of the library records a CUDA event right after its launch;
 * the report (text, one line per launch site "file:line launches total_us share") is written to buf.
 * dsir_profile_report synchronises the device. */
int dsir_profile_begin(dsir_stream_t stream);
int dsir_profile_report(char *buf, size_t buf_bytes);

/* ------------------------------------------------------------------------------------------------
 * xyz k-nearest neighbours.  Replaces torch_points_kernels.knn(pos_support, pos, k) as called at
 * dataloader/data_base.py:165,170.  d2 = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32; results ascending in
 * (d2, support index) — ties go to the lower index.
 *   support [B,Ns,sup_stride>=3]  query [B,Nq,qry_stride>=3]  (strides in floats; xyz are the first 3)
 *   idx [B,Nq,k] int64            dist2 [B,Nq,k] fp32 or NULL
 * ---------------------------------------------------------------------------------------------- */
size_t dsir_knn_workspace_bytes(int B, int Ns, int Nq, int k, int algo);
int dsir_knn_xyz(const float *support, int sup_stride, const float *query, int qry_stride, int B, int Ns, int Nq,
                 int k, int64_t *idx, float *dist2, void *ws, size_t ws_bytes, int algo, dsir_stream_t stream);

/* The 4-level pyramid of DataBase.nn_search (dataloader/data_base.py:153-183) for one cloud tensor
 * pts [B,N,pt_stride]: per level self-kNN (k), pool = first N_l/ratio rows, 1-NN of every level point
 * into the first N_l/ratio points; level-local indices, levels concatenated along the point axis:
 *   xyz_cat [B,sumN,3]  neigh [B,sumN,k]  sub [B,sumSub,k]  interp [B,sumN,1]                       */
size_t dsir_knn_pyramid_workspace_bytes(int B, int N, int k, const int *ratios, int L, int algo);
int dsir_knn_pyramid(const float *pts, int pt_stride, int B, int N, const int *ratios, int L, int k, float *xyz_cat,
                     int64_t *neigh, int64_t *sub, int64_t *interp, void *ws, size_t ws_bytes, int algo,
                     dsir_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Feature distance.  d_jk = ((-2 <s_j, r_k>) + |s_j|^2) + |r_k|^2 in fp32, the op order of
 * square_distance_V2 (network/matchnet.py:110-112).  Features are addressed as
 *   f[b, c, n] = base[b * batch_stride + c * chan_stride + n * point_stride]        (strides in floats)
 * so [B,C,N], [B,N,C] and the row slices of network/model.py:565 are all passed without copies.
 * ---------------------------------------------------------------------------------------------- */
typedef struct dsir_feat {
    const float *ptr;
    int64_t batch_stride, chan_stride, point_stride;
} dsir_feat;

/* materialising form: match_features[_V2] / square_distance[_V2] / feat_dist (matchnet.py:49-192) -> dist [B,J,K] */
size_t dsir_match_dense_workspace_bytes(int B, int J, int K);
int dsir_match_dense(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int metric, float *dist, void *ws,
                     size_t ws_bytes, dsir_stream_t stream);

/* fused distance + row argmin: replaces the chunked block network/model.py:558-569.
 *   idx [B,J] int64 = first index of the row minimum; min_d [B,J] fp32 or NULL.                     */
size_t dsir_match_argmin_workspace_bytes(int B, int C, int J, int K, int algo);
int dsir_match_argmin(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int64_t *idx, float *min_d, void *ws,
                      size_t ws_bytes, int algo, dsir_stream_t stream);
/* the same with a HINT: prior_idx [B,J] int64 (may be NULL, may alias idx) = correspondences of a previous iteration of the
 * loop (network/model.py:551-601).  The distance to the prior match bounds the row minimum from the start, which spares the
 * filter its priming pass and most of its candidate updates; the result never depends on the hint. */
int dsir_match_argmin_hint(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, int64_t *idx, float *min_d,
                           const int64_t *prior_idx, void *ws, size_t ws_bytes, int algo, dsir_stream_t stream);

/* diagnostic for the tcgen05 path (synchronises `stream`): rows of the LAST dsir_match_argmin call on this workspace
 * whose candidate list saturated and were recomputed exhaustively in fp32.  host_out is a HOST int32. */
int dsir_match_argmin_rescued_rows(const void *ws, size_t ws_bytes, int B, int C, int J, int K, int32_t *host_out,
                                   dsir_stream_t stream);

/* diagnostic for the tcgen05 path (synchronises `stream`): device-side timing of the LAST filter kernel launched on this
 * workspace.  host_out[0] = kernel span in ns (%globaltimer, last CTA end - first CTA start), host_out[1] = mean SM
 * cycles per CTA, host_out[2] = mean SM cycles per 512x128 unit (tensor-pipe floor: 1280).  host_out is HOST memory. */
int dsir_match_argmin_filter_timing(const void *ws, size_t ws_bytes, int B, int C, int J, int K, double *host_out,
                                    dsir_stream_t stream);

/* diagnostic (only filled when the environment variable DSIR_TC_DEBUG has bit 1 set): SM-clock stamps of CTA 0's first
 * 256 units of the last filter launch; host_out[4096] u32, layout documented at match_tc_filter_trace (match_tc.cu). */
int dsir_match_argmin_filter_trace(const void *ws, size_t ws_bytes, int B, int C, int J, int K, uint32_t *host_out,
                                   dsir_stream_t stream);

/* fused distance + affinity + row softmax + soft target (never materialises [J,K]):
 *   a_jk = -beta_b (d_jk - alpha_b) (+ col_bias[b,k])           compute_affinity, matchnet.py:195-208
 *   lse_j = log sum_k exp(a_jk);  w_jk = exp(a_jk - lse_j)      row pass of sinkhorn, matchnet.py:259
 *   y_j = sum_k w_jk xyz_ref[b,k,:] / (sum_k w_jk + 1e-16)      soft target, network/model.py:81-84
 * outputs: y_soft [B,J,3] (or NULL), lse [B,J] (or NULL); topk > 0 additionally returns the topk largest
 * weights per row, descending, ties to the lower index: topk_idx [B,J,topk] int64, topk_w [B,J,topk] (topk <= 32;
 * workspace from dsir_match_soft_topk_workspace_bytes).  Without a column bias, C <= 64 and K >= 40 topk (roughly) the
 * top-k comes from two more tensor-core sweeps (granule minima -> per-row threshold -> <= 4 topk listed columns, re-scored
 * exactly): nothing of size J x K exists.  Otherwise row chunks of the exact fp32 distance matrix are materialised inside
 * the workspace (<= 256 MB).  Same bits either way. */
size_t dsir_match_soft_workspace_bytes(int B, int C, int J, int K);
size_t dsir_match_soft_topk_workspace_bytes(int B, int C, int J, int K, int topk);
/* 1 when dsir_match_soft(topk) with col_bias == NULL takes the fused tensor-core route for this shape, else 0 */
int dsir_match_soft_topk_fused(int B, int C, int J, int K, int topk);
/* diagnostic (synchronises `stream`): rows of the LAST fused top-k call on this workspace whose column list overflowed or
 * came up short and were therefore scanned exhaustively (mass ties, beta <= 0, NaN rows).  host_out is HOST memory. */
int dsir_match_soft_topk_exhaustive_rows(const void *ws, size_t ws_bytes, int B, int C, int J, int K, int topk,
                                         int32_t *host_out, dsir_stream_t stream);
/* one sweep of an iteration that calls the fused soft match repeatedly with the SAME features and workspace and only the
 * column bias changing (the Sinkhorn half-steps of matchnet.py:211-271 in dual form): with reuse_prep != 0 the norms and
 * the tensor-core operand copies left in `ws` by the previous sweep are used as they are. */
int dsir_match_soft_sweep(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, const float *beta, const float *alpha,
                          const float *col_bias, const float *xyz_ref, float *y_soft, float *lse, int reuse_prep, void *ws,
                          size_t ws_bytes, dsir_stream_t stream);
int dsir_match_soft(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, const float *beta, const float *alpha,
                    const float *col_bias, const float *xyz_ref, float *y_soft, float *lse, int topk,
                    int64_t *topk_idx, float *topk_w, void *ws, size_t ws_bytes, dsir_stream_t stream);

/* gather_neighbour_V3 (network/tools.py:211-221): out[b,c,m] = in[b,c,idx[b,m]]; in [B,C,N], idx [B,M] */
int dsir_gather_points(const float *in, int B, int C, int N, const int64_t *idx, int M, float *out,
                       dsir_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * KNN consumers of the RandLA-Net local aggregation (the stage right after the KNN pyramid) and Sinkhorn.
 * All tensors contiguous.  Indices outside [0,N) (the -1 padding of a short neighbour list) read as 0 / are skipped.
 * ---------------------------------------------------------------------------------------------- */
/* gather_neighbour_V2 (network/tools.py:197-209): in [B,C,N], idx [B,M,k] int64 -> out [B,C,M,k] */
int dsir_gather_neighbours(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out,
                           dsir_stream_t stream);
/* Building_block.relative_pos_encoding (network/RandLANet.py:197-212): xyz [B,3,N], idx [B,N,k] -> out [B,10,N,k] =
 * { |rel|, rel(3) = neighbour - centre, centre(3), neighbour(3) } */
int dsir_rel_pos_encoding(const float *xyz, int B, int N, const int64_t *idx, int k, float *out, dsir_stream_t stream);
/* RandLA.random_sample (network/RandLANet.py:374-391): in [B,C,N], pool_idx [B,M,k] -> out [B,C,M] = max over the k
 * gathered neighbours; the [B,C,M,k] intermediate is never written.  k <= 32. */
int dsir_pool_max(const float *in, int B, int C, int N, const int64_t *idx, int M, int k, float *out, dsir_stream_t stream);
/* sinkhorn (network/matchnet.py:211-271): log_alpha [B,J,K] -> out [B,J,K] after n_iters row/column normalisations in
 * the log domain, with or without the slack row/column; out may alias log_alpha.  ws >= dsir_sinkhorn_workspace_bytes. */
size_t dsir_sinkhorn_workspace_bytes(int B, int J, int K);
/* log_optimal_transport + log_sinkhorn_iterations (network/matchnet.py:827-856, SuperGlue's dustbin OT): scores [B,M,N],
 * learned dustbin score alpha (ONE float on the device: the reference's nn.Parameter), `iters` Sinkhorn iterations in the log domain with marginals (1 x M, N) / (1 x N, M) over
 * M + N -> out [B,M+1,N+1] = Z + u + v - norm.  The augmented coupling matrix is never built. */
size_t dsir_log_optimal_transport_workspace_bytes(int B, int M, int N);
int dsir_log_optimal_transport(const float *scores, int B, int M, int N, const float *alpha, int iters, float *out, void *ws,
                               size_t ws_bytes, dsir_stream_t stream);
int dsir_sinkhorn(const float *log_alpha, int B, int J, int K, int n_iters, int slack, float *out, void *ws, size_t ws_bytes,
                  dsir_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Weighted Kabsch.  Replaces compute_rigid_transform_2 (network/model.py:22-66) including its host
 * round trip (fp64 LAPACK SVD at :47).  Points are addressed as p[b,m,i] = base[b*bs + m*ps + i*cs]
 * so both [B,M,3] (model.py:586) and [B,3,M] (the loop's layout) work without a transpose copy.
 *   weights w[b,m] = w_base[b * w_bs + m]
 *   gather (nullable) [B,M] int64: tgt point for row m is tgt[b, gather[b,m]] (fuses model.py:571)
 *   n_tgt: points of the tgt cloud the gather indices refer to; > 0: an index outside [0, n_tgt) is never dereferenced
 *          and poisons its pair (status 1) - the reference's torch.gather raises (network/tools.py:211-221); 0: unchecked
 *   T [B,3,4] fp32;  status [B] int32: 0 ok, 1 non-finite input / bad gather index -> identity written.  Rank-deficient
 *          covariances are NOT an error (the reference does not flag them either): rank 1 -> the minimal rotation that
 *          aligns the one determined direction, rank 0 -> R = I; the centroid translation is kept.
 *   moments (nullable) [B,17] fp64: additive raw moments {S|w|, Sw, Swx(3), Swy(3), Swxy(9)}
 * ---------------------------------------------------------------------------------------------- */
typedef struct dsir_points {
    const float *ptr;
    int64_t batch_stride, point_stride, coord_stride;
} dsir_points;

size_t dsir_kabsch_workspace_bytes(int B, int M);
int dsir_kabsch(dsir_points src, dsir_points tgt, const float *w, int64_t w_batch_stride, const int64_t *gather,
                int B, int M, int n_tgt, float *T, int32_t *status, double *moments, void *ws, size_t ws_bytes,
                dsir_stream_t stream);
/* first half only (for row-block sharding across GPUs): moments [B,17] of this rank's rows */
int dsir_kabsch_moments(dsir_points src, dsir_points tgt, const float *w, int64_t w_batch_stride,
                        const int64_t *gather, int B, int M, int n_tgt, double *moments, void *ws, size_t ws_bytes,
                        dsir_stream_t stream);
/* second half: centred covariance from (all-reduced) moments, fp64 3x3 SVD, det fix, R,t */
int dsir_kabsch_from_moments(const double *moments, int B, float *T, int32_t *status, dsir_stream_t stream);
/* the first statements of compute_rigid_transform (network/model.py:81-84) for a GIVEN weight matrix:
 * weights [B,M,N] (w[b,j,k] = base[b*w_batch_stride + j*w_row_stride + k]), tgt [B,N,3] contiguous ->
 * rowmass [B,M] = sum_k w_jk,  y_soft [B,M,3] = (sum_k w_jk tgt_k) / (rowmass + 1e-16).  W is read once. */
int dsir_soft_targets(const float *weights, int64_t w_batch_stride, int64_t w_row_stride, const float *tgt, int B, int M, int N,
                      float *y_soft, float *rowmass, dsir_stream_t stream);
/* soft variant, compute_rigid_transform (network/model.py:68-116) given the fused soft targets:
 * src [B,M,3], y_soft [B,M,3], rowmass [B,M] (sum_k W_jk) */
int dsir_kabsch_soft(dsir_points src, const float *y_soft, const float *rowmass, int B, int M, float *T,
                     int32_t *status, void *ws, size_t ws_bytes, dsir_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SE(3) helpers (common/math/se3_torch.py).  T are [B,3,4] fp32 (a [B,4,4] input is passed with
 * row_stride 4 and batch_stride 16 — only the first three rows are read).
 * ---------------------------------------------------------------------------------------------- */
int dsir_se3_apply(const float *T, int64_t T_batch_stride, dsir_points pts, int B, int N, float *out,
                   int64_t out_batch_stride, int64_t out_point_stride, int64_t out_coord_stride,
                   int rotate_only, dsir_stream_t stream);                         /* se3_torch.py:51-100 */
int dsir_se3_compose(const float *a, int64_t a_bs, const float *b, int64_t b_bs, int B, float *out,
                     dsir_stream_t stream);                                        /* se3_torch.py:28-48 */
int dsir_se3_inverse(const float *T, int64_t T_bs, int B, float *out, dsir_stream_t stream); /* :10-25 */

/* ------------------------------------------------------------------------------------------------
 * The iterative re-match / re-solve loop of forward_align_4 (network/model.py:551-601) with fixed
 * features and weights (the two neural stages are outside the path): per iteration
 *   idx = argmin match(feat_src, feat_ref); tgt = xyz_ref[idx]; T = kabsch(xyz_src, tgt, w);
 *   xyz_src <- T xyz_src;  T_total <- T o T_total
 * entirely on `stream`, no host synchronisation.
 *   xyz_src [B,3,J] fp32 IN/OUT, xyz_ref [B,3,K], weights [B,J]
 *   transforms [iters,B,3,4] cumulative; pred_idx [iters,B,J] int64 (or NULL); status [iters,B] int32
 * ---------------------------------------------------------------------------------------------- */
size_t dsir_align_loop_workspace_bytes(int B, int C, int J, int K, int algo);
int dsir_align_loop(dsir_feat fs, dsir_feat fr, int B, int C, int J, int K, float *xyz_src, const float *xyz_ref,
                    const float *weights, int iters, float *transforms, int64_t *pred_idx, int32_t *status,
                    void *ws, size_t ws_bytes, int algo, dsir_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Key-point scoring and selection on the KNN graph.
 * dsir_keypoint_score = Network.score_fun (network/model.py:700-757):
 *   feat [B,C,N], xyz [B,3,N], prob [B,N] (NULL: no probability gate), label [B,N] int64 (NULL: semantic score 1),
 *   label_weights [num_class] (model.py:146-150), neigh_idx [B,N,idx_stride] int64 of which the first k (<= 32; the
 *   reference uses 16) are read, ball_r (2.0 in the reference)  ->  score [B,N]
 * dsir_topk_rows = torch.topk(score, k, dim=-1, largest=True) (model.py:692): values/index [B,k], descending, ties to the
 *   LOWER index, NaN first.  k <= 16384.
 * ---------------------------------------------------------------------------------------------- */
size_t dsir_keypoint_score_workspace_bytes(int B, int C, int N);
int dsir_keypoint_score(const float *feat, const float *xyz, const float *prob, const int64_t *label,
                        const float *label_weights, int num_class, const int64_t *neigh_idx, int idx_stride, int k,
                        float ball_r, int B, int C, int N, float *score, void *ws, size_t ws_bytes, dsir_stream_t stream);
int dsir_topk_rows(const float *score, int B, int N, int k, float *values, int64_t *index, dsir_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * On-device evaluation of a registration result.
 * dsir_pose_errors: T_pred, T_gt [B,3,4] -> out [B,4] = { err_r_deg, err_t  (common/metrics_util.py:55-61: residual of
 *   inverse(gt) o pred), rre_deg, rte (metrics_util.py:27-33 rte_rre) }, success [B] int32 (may be NULL) =
 *   err_t < rte_thresh && err_r_deg < rre_thresh (metrics_util.py:63).
 * dsir_correspondence_check = Loss.find_correct_correspondence (network/loss.py:723-749): pos_pairs = the ground-truth
 *   pair lists of the batch concatenated [total,2] int32 with pos_offsets [B+1] int64 (device), pred_pairs [B,N,2] int32
 *   (model.py:599-601), hash_seed [B] int64 (device; loss.py:738-742)  ->  correct [B,N] uint8 = np.isin(key(pred),
 *   key(pos)) with key = a0 + a1 * seed (loss.py:280-294).
 * dsir_nn_sqdist_mean: a [B,N,3], b [B,M,3] -> min_d [B,N] (may be NULL) = min_k |a_j - b_k|^2 by direct differences and
 *   mean [B] (may be NULL): one side of the modified chamfer distance (metrics_util.py:38-40, 72-74).
 * ---------------------------------------------------------------------------------------------- */
int dsir_pose_errors(const float *T_pred, const float *T_gt, int B, float rte_thresh, float rre_thresh, float *out,
                     int32_t *success, dsir_stream_t stream);
size_t dsir_correspondence_check_workspace_bytes(int64_t total_pos);
int dsir_correspondence_check(const int32_t *pos_pairs, const int64_t *pos_offsets, int64_t total_pos,
                              const int32_t *pred_pairs, int B, int N, const int64_t *hash_seed, uint8_t *correct, void *ws,
                              size_t ws_bytes, dsir_stream_t stream);
size_t dsir_nn_sqdist_workspace_bytes(int B);
int dsir_nn_sqdist_mean(const float *a, const float *b, int B, int N, int M, float *min_d, float *mean, void *ws,
                        size_t ws_bytes, dsir_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSIR_B200_H_ */
