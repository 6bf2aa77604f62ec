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
