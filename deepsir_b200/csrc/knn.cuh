#pragma once
#include "common.cuh"

#define DSIR_MAX_LEVELS 8

namespace dsir {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;  // support points per shared-memory stage (16 KB as float4)

struct KnnBruteParams {
    const float4 *sup4;  // packed support, [B][sup_bs] float4 (xyz0); the first Ns of each batch are used
    long long sup_bs;
    const float *query;  // query xyz at query[b*qry_bs + q*qry_stride + {0,1,2}]
    long long qry_bs;
    int qry_stride;
    int Ns, Nq, k;
    int64_t *idx;  // idx[b*idx_bs + q*k + p]
    long long idx_bs;
    float *dist2;   // same addressing as idx, nullable
    int64_t *idx2;  // optional second copy of rows q < idx2_rows (the pyramid's pooling indices)
    long long idx2_bs;
    int idx2_rows;
};

struct PyramidLevels {
    int L;
    int n[DSIR_MAX_LEVELS];    // points of level l
    int m[DSIR_MAX_LEVELS];    // n[l] / ratio[l]
    int off[DSIR_MAX_LEVELS];  // row offset of level l inside the concatenated [sumN] axis
    int offsub[DSIR_MAX_LEVELS];
    int sumN, sumSub;
};

int launch_pack_xyz4(const float *pts, int stride, long long total, float4 *out, cudaStream_t st);
int launch_knn_brute(const KnnBruteParams &P, int B, cudaStream_t st);
int launch_pyramid_xyz(const float *pts, int stride, int B, int N, const PyramidLevels &lv, float *xyz_cat,
                       cudaStream_t st);

}  // namespace dsir

// ---------------------------------------------------------------------------------------------------
// grid variant (knn_grid.cu)
// ---------------------------------------------------------------------------------------------------
namespace dsir {

constexpr int KNN_GRID_GMAX = 32768;     // cells per grid (shared-memory histogram of the build kernel)
constexpr int KNN_GRID_MIN_POINTS = 512;  // below this the brute-force kernel is used

struct KnnGridHeader {
    float lo[3];
    float h, inv_h;
    int gx, gy, gz;
    float slack;  // absolute rounding slack added to every search radius
    float diag;
    int pad[6];
};

struct KnnGridBuildParams {
    const float4 *pts4;  // [B][pts_bs] packed xyz0; grid g is built over the first n[g] points of every cloud
    long long pts_bs;
    int n[DSIR_MAX_LEVELS + 1];
    KnnGridHeader *hdr[DSIR_MAX_LEVELS + 1];  // [B]
    int *cell_start[DSIR_MAX_LEVELS + 1];     // [B][gmax+1]
    float4 *sorted[DSIR_MAX_LEVELS + 1];      // [B][n[g]]  (x,y,z,index bits) in cell order
    int gmax;
    float cells_per_point;
};

struct KnnGridQueryParams {
    const KnnGridHeader *hdr;  // support grid
    const int *cell_start;
    const float4 *sorted;
    int gmax, Ns;
    // queries: either another grid's sorted array (cell-coherent order, original index in .w) or a raw array
    const float4 *q_sorted;
    long long q_sorted_bs;
    const float *query;
    long long qry_bs;
    int qry_stride;
    int Nq, k;
    float r0_cells;  // first search radius in cell sizes
    int64_t *idx;
    long long idx_bs;
    float *dist2;
    int64_t *idx2;
    long long idx2_bs;
    int idx2_rows;
};

size_t knn_grid_smem_bytes(int gmax);
int launch_knn_grid_build(const KnnGridBuildParams &P, int ngrids, int B, cudaStream_t st);
int launch_knn_grid_query(const KnnGridQueryParams &P, int B, cudaStream_t st);

}  // namespace dsir

// ---------------------------------------------------------------------------------------------------
// bucket-tree variant (knn_tree.cu): Morton-ordered leaves of 32 points, one warp per query leaf
// ---------------------------------------------------------------------------------------------------
namespace dsir {

constexpr int KNN_TREE_MIN_POINTS = 512;     // below this the brute-force kernel is used
constexpr int KNN_TREE_MAX_POINTS = 17408;   // the build sorts one cloud in the shared memory of one CTA (12 bytes per point)
constexpr int KNN_TREE_MAX_K = 16;           // sorted register list of the query kernel; larger k -> grid path

struct KnnLeaf {   // 512 bytes: one cp.async.bulk per leaf
    float nx[32], ny[32], nz[32];   // NEGATED coordinates (dx = q + nx is exactly q - x); padding slots hold -inf
    int idx[32];                    // original index; 0x7fffffff in padding slots
};

struct KnnTreeView {
    const KnnLeaf *leaves;   // [B][nleaf]
    const float *box;        // [B][6][nlpad]: lo.x, lo.y, lo.z, hi.x, hi.y, hi.z per leaf
    const float *sbox;       // [B][6][32]: the same per supernode (32 consecutive leaves)
    int n, nleaf, nlpad, nsuper;
};

struct KnnTreeBuildParams {
    const float4 *pts4;   // [B][pts_bs] packed xyz0; tree g is built over the first n[g] points of every cloud
    long long pts_bs;
    int n[DSIR_MAX_LEVELS + 1];
    KnnLeaf *leaves[DSIR_MAX_LEVELS + 1];
    float *box[DSIR_MAX_LEVELS + 1];
    float *sbox[DSIR_MAX_LEVELS + 1];
    int cap;              // shared-memory capacity in points (>= every n[g])
};

struct KnnTreeQueryParams {
    KnnTreeView sup, qry;   // support tree; query tree (its leaves are the warps' query groups)
    int self;               // query tree == support tree: every warp starts with its own leaf
    int k;
    int64_t *idx;           // idx[b*idx_bs + q*k + p], q = ORIGINAL query index
    long long idx_bs;
    float *dist2;           // same addressing, nullable
    int64_t *idx2;          // optional second copy of rows q < idx2_rows (the pyramid's pooling indices)
    long long idx2_bs;
    int idx2_rows;
};

size_t knn_tree_slot_bytes(int B, int n);
bool knn_tree_take_slot(Workspace &W, int B, int n, KnnTreeView *v);
int launch_knn_tree_build(const float4 *pts4, long long pts_bs, const KnnTreeView *trees, int ntrees, int B, cudaStream_t st);
int launch_knn_tree_query(const KnnTreeQueryParams &P, int B, cudaStream_t st);

}  // namespace dsir
