/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * CPU restatement of the dense k-nearest-neighbour contract that the
 * reference consumes at dataloader/data_base.py:165 and :170
 * (`Util.knn(support, query, k)` == torch_points_kernels.knn, a third-party
 * C++/nanoflann kernel that is NOT vendored in /root/reference and whose
 * version is unpinned there).  PARITY UNPINNED against that third-party
 * kernel: there is no golden vector for it in the reference.  What IS pinned:
 * call-site semantics (argument order (support, query, k); self-query includes
 * the point itself; results ascending by squared L2; int64 indices), and the
 * tie rule that THIS project defines because nanoflann's is traversal
 * dependent:
 *
 *     d2(q, s) = fma(dz, dz, fma(dy, dy, dx * dx))   in fp32,  dx = q.x - s.x ...
 *     order    = lexicographic (d2, support index)
 *
 * Build:  gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC knn_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float d2_fp32(const float *q, const float *s)
{
    float dx = q[0] - s[0];
    float dy = q[1] - s[1];
    float dz = q[2] - s[2];
    float d = dx * dx;
    d = fmaf(dy, dy, d);
    d = fmaf(dz, dz, d);
    return d;
}

/* support [B,Ns,3], query [B,Nq,3] contiguous fp32; idx [B,Nq,k] int64; dist2 [B,Nq,k] (nullable).
 * returns 0, or -1 when Ns < k (the reference kernel raises in that case). */
int oracle_knn(const float *support, const float *query, int B, int Ns, int Nq, int k,
               int64_t *idx, float *dist2)
{
    if (k <= 0 || Ns < k) return -1;
    long total = (long)B * Nq;
#pragma omp parallel
    {
        float *bd = (float *)malloc(sizeof(float) * (size_t)k);
        int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
        for (long t = 0; t < total; ++t) {
            int b = (int)(t / Nq);
            const float *q = query + 3 * t;
            const float *S = support + (size_t)3 * Ns * b;
            int cnt = 0;
            for (int s = 0; s < Ns; ++s) {
                float d = d2_fp32(q, S + 3 * s);
                /* ascending scan of s: strict '<' keeps the lower index first on ties */
                if (cnt == k && !(d < bd[k - 1])) continue;
                int p = (cnt < k) ? cnt++ : k - 1;
                while (p > 0 && d < bd[p - 1]) {
                    bd[p] = bd[p - 1];
                    bi[p] = bi[p - 1];
                    --p;
                }
                bd[p] = d;
                bi[p] = s;
            }
            memcpy(idx + (size_t)k * t, bi, sizeof(int64_t) * (size_t)k);
            if (dist2) memcpy(dist2 + (size_t)k * t, bd, sizeof(float) * (size_t)k);
        }
        free(bd);
        free(bi);
    }
    return 0;
}

/* Pyramid driver restating DataBase.nn_search (dataloader/data_base.py:153-183) for ONE cloud
 * tensor pts [B,N,stride>=3]: per level l: self-kNN of the level cloud, pool = first N_l/ratio rows,
 * sub-cloud = first N_l/ratio points, 1-NN of every level point into the sub-cloud, then recurse on
 * the sub-cloud.  Outputs are concatenated along the point axis, indices level-local:
 *   xyz_cat [B,sumN,3]  neigh [B,sumN,k]  sub [B,sumSub,k]  interp [B,sumN,1]                      */
int oracle_knn_pyramid(const float *pts, int B, int N, int pt_stride, const int *ratios, int L, int k,
                       float *xyz_cat, int64_t *neigh, int64_t *sub, int64_t *interp)
{
    long sumN = 0, sumSub = 0;
    {
        int n = N;
        for (int l = 0; l < L; ++l) { sumN += n; sumSub += n / ratios[l]; n = n / ratios[l]; }
    }
    float *cur = (float *)malloc(sizeof(float) * 3 * (size_t)B * N);
    float *nxt = (float *)malloc(sizeof(float) * 3 * (size_t)B * N);
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i)
            for (int c = 0; c < 3; ++c)
                cur[((size_t)b * N + i) * 3 + c] = pts[((size_t)b * N + i) * pt_stride + c];
    int n = N;
    long offN = 0, offSub = 0;
    int rc = 0;
    for (int l = 0; l < L && rc == 0; ++l) {
        int m = n / ratios[l];
        int64_t *nb = (int64_t *)malloc(sizeof(int64_t) * (size_t)B * n * k);
        int64_t *up = (int64_t *)malloc(sizeof(int64_t) * (size_t)B * n);
        rc = oracle_knn(cur, cur, B, n, n, k, nb, NULL);
        for (int b = 0; b < B; ++b)
            memcpy(nxt + (size_t)b * m * 3, cur + (size_t)b * n * 3, sizeof(float) * 3 * (size_t)m);
        if (rc == 0) rc = oracle_knn(nxt, cur, B, m, n, 1, up, NULL);
        if (rc == 0)
            for (int b = 0; b < B; ++b) {
                memcpy(xyz_cat + ((size_t)b * sumN + offN) * 3, cur + (size_t)b * n * 3, sizeof(float) * 3 * (size_t)n);
                memcpy(neigh + ((size_t)b * sumN + offN) * k, nb + (size_t)b * n * k, sizeof(int64_t) * (size_t)n * k);
                memcpy(sub + ((size_t)b * sumSub + offSub) * k, nb + (size_t)b * n * k, sizeof(int64_t) * (size_t)m * k);
                memcpy(interp + ((size_t)b * sumN + offN), up + (size_t)b * n, sizeof(int64_t) * (size_t)n);
            }
        free(nb);
        free(up);
        offN += n;
        offSub += m;
        float *t = cur; cur = nxt; nxt = t;
        n = m;
    }
    free(cur);
    free(nxt);
    return rc;
}
