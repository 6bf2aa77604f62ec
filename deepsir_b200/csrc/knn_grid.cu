// Exact xyz k-nearest neighbours on a uniform grid (the fast path for level clouds of >= 512 points).
//
// Same contract and arithmetic as the brute-force kernel (knn.cu) and oracle/knn_oracle.c:
//     d2 = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32, results ascending in (d2, support index).
// Exactness does not depend on the grid: a query scans the cells overlapping the cube [q-R, q+R]; it stops only
// when its k-th distance is strictly below R^2, i.e. when every unscanned point (which differs from q by more
// than R along some axis) is provably farther than the k-th best.  Otherwise R grows (to the k-th distance found
// so far, or x2 while fewer than k points were seen) and only the NEW shell of cells is scanned.
//
//   build   one CTA per (cloud, grid): bounding box -> cell size -> shared-memory histogram -> scan ->
//           counting sort of (x,y,z,index) into cell order.  Cells are x-fastest, so one (y,z) row of a query
//           cube is ONE contiguous run of sorted points.
//   query   one thread per query, queries taken in the cell order of their own grid so that the lanes of a warp
//           are spatial neighbours (same cells -> L1 hits, little divergence); sorted top-k in registers.
#include "knn.cuh"

namespace dsir {

namespace {

constexpr int GRID_BUILD_THREADS = 1024;

__device__ __forceinline__ int cell_of(float v, float lo, float inv_h, int n) {
    int c = (int)floorf(__fmul_rn(__fsub_rn(v, lo), inv_h));
    return min(max(c, 0), n - 1);
}

__global__ __launch_bounds__(GRID_BUILD_THREADS) void knn_grid_build_kernel(KnnGridBuildParams P) {
    extern __shared__ int s_cnt[];  // [gmax + 1] histogram, then cursors
    __shared__ float s_red[6][32];
    __shared__ KnnGridHeader s_hdr;
    __shared__ int s_warp_tot[32];

    const int g = blockIdx.x, b = blockIdx.y;
    const int n = P.n[g];
    const float4 *pts = P.pts4 + (size_t)b * P.pts_bs;
    KnnGridHeader *hdr_out = P.hdr[g] + b;
    int *cell_start = P.cell_start[g] + (size_t)b * (P.gmax + 1);
    float4 *sorted = P.sorted[g] + (size_t)b * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- bounding box ----
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += GRID_BUILD_THREADS) {
        float4 p = pts[i];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = warp_min(lo[a]), h = warp_max(hi[a]);
        if (lane == 0) { s_red[a][warp] = l; s_red[3 + a][warp] = h; }
    }
    __syncthreads();
    if (tid == 0) {
        float L[3], H[3], ext[3];
        for (int a = 0; a < 3; ++a) {
            L[a] = INFINITY; H[a] = -INFINITY;
            for (int w = 0; w < GRID_BUILD_THREADS / 32; ++w) { L[a] = fminf(L[a], s_red[a][w]); H[a] = fmaxf(H[a], s_red[3 + a][w]); }
            if (!(L[a] <= H[a])) { L[a] = 0.f; H[a] = 0.f; }   // all-NaN axis: one cell
            ext[a] = H[a] - L[a];
        }
        float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
        if (!(emax > 0.f) || !isfinite(emax)) emax = 1.f;
        float e[3];
        for (int a = 0; a < 3; ++a) e[a] = fmaxf(ext[a], emax * 1e-3f);
        // target ~P.cells_per_point * n cells of equal size; shrink the target until the rounded-up grid fits
        float target = fminf((float)P.gmax * 0.9f, fmaxf(8.f, P.cells_per_point * (float)n));
        int gx = 1, gy = 1, gz = 1;
        float h = emax;
        for (int iter = 0; iter < 64; ++iter) {
            h = cbrtf(e[0] * e[1] * e[2] / target);
            gx = (int)fminf(ext[0] / h, 4.0e6f) + 1;
            gy = (int)fminf(ext[1] / h, 4.0e6f) + 1;
            gz = (int)fminf(ext[2] / h, 4.0e6f) + 1;
            if ((long long)gx * gy * gz <= (long long)P.gmax) break;
            target *= 0.8f;
        }
        if ((long long)gx * gy * gz > (long long)P.gmax) { gx = gy = gz = 1; h = emax; }
        s_hdr.lo[0] = L[0]; s_hdr.lo[1] = L[1]; s_hdr.lo[2] = L[2];
        s_hdr.h = h;
        s_hdr.inv_h = 1.0f / h;
        s_hdr.gx = gx; s_hdr.gy = gy; s_hdr.gz = gz;
        float amax = 0.f;
        for (int a = 0; a < 3; ++a) amax = fmaxf(amax, fmaxf(fabsf(L[a]), fabsf(H[a])));
        s_hdr.slack = amax * 4.8e-7f + 1e-30f;   // 4 ulp of the largest coordinate: covers the rounding of q +- R
        s_hdr.diag = sqrtf(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2]);
        *hdr_out = s_hdr;
    }
    __syncthreads();
    const KnnGridHeader H = s_hdr;
    const int G = H.gx * H.gy * H.gz;

    // ---- histogram ----
    for (int c = tid; c <= G; c += GRID_BUILD_THREADS) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += GRID_BUILD_THREADS) {
        float4 p = pts[i];
        int c = (cell_of(p.z, H.lo[2], H.inv_h, H.gz) * H.gy + cell_of(p.y, H.lo[1], H.inv_h, H.gy)) * H.gx +
                cell_of(p.x, H.lo[0], H.inv_h, H.gx);
        atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();

    // ---- exclusive scan over G cells: thread t owns a contiguous slice ----
    const int per = (G + GRID_BUILD_THREADS - 1) / GRID_BUILD_THREADS;
    const int c0 = min(tid * per, G), c1 = min(c0 + per, G);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += s_cnt[c];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp_tot[lane];
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += u;
        }
        s_warp_tot[lane] = inc2 - v;
    }
    __syncthreads();
    int run = s_warp_tot[warp] + incl - sum;
    for (int c = c0; c < c1; ++c) {
        int cnt = s_cnt[c];
        s_cnt[c] = run;          // becomes the scatter cursor
        cell_start[c] = run;
        run += cnt;
    }
    if (tid == GRID_BUILD_THREADS - 1) cell_start[G] = n;
    __syncthreads();

    // ---- counting-sort scatter (order inside a cell is arbitrary: the query orders by (d2, index) itself) ----
    for (int i = tid; i < n; i += GRID_BUILD_THREADS) {
        float4 p = pts[i];
        int c = (cell_of(p.z, H.lo[2], H.inv_h, H.gz) * H.gy + cell_of(p.y, H.lo[1], H.inv_h, H.gy)) * H.gx +
                cell_of(p.x, H.lo[0], H.inv_h, H.gx);
        int pos = atomicAdd(&s_cnt[c], 1);
        p.w = __int_as_float(i);
        sorted[pos] = p;
    }
}

// lexicographic (d, idx) sorted insertion; caller guarantees (d, s) < (bd[K-1], bi[K-1]).  Distance ties are rare
// (duplicated points), so the index comparisons live on a separate, warp-uniformly branched path.
template <int KMAX>
__device__ __forceinline__ void topk_insert_lex(float (&bd)[KMAX], int (&bi)[KMAX], float d, int s) {
    bool tie = false;
#pragma unroll
    for (int p = 0; p < KMAX; ++p) tie = tie || (d == bd[p]);
    if (!tie) {
#pragma unroll
        for (int p = KMAX - 1; p >= 0; --p) {
            const int pm = p > 0 ? p - 1 : 0;
            const bool shift = (p > 0) && (d < bd[pm]);
            const bool here = !shift && (d < bd[p]);
            bd[p] = shift ? bd[pm] : (here ? d : bd[p]);
            bi[p] = shift ? bi[pm] : (here ? s : bi[p]);
        }
    } else {
#pragma unroll
        for (int p = KMAX - 1; p >= 0; --p) {
            const int pm = p > 0 ? p - 1 : 0;
            const bool shift = (p > 0) && (d < bd[pm] || (d == bd[pm] && s < bi[pm]));
            const bool here = !shift && (d < bd[p] || (d == bd[p] && s < bi[p]));
            bd[p] = shift ? bd[pm] : (here ? d : bd[p]);
            bi[p] = shift ? bi[pm] : (here ? s : bi[p]);
        }
    }
}

// Per-thread max-heap of the k best (d, idx) pairs in SHARED memory, column layout [slot][thread]: the bank of an
// access is thread % 32 whatever the slot, so lanes that touch different slots never conflict.  A sorted register list
// costs ~100 compare/select instructions per insertion on the ALU pipe, which bounded the kernel; the heap needs
// <= log2(k) levels of two loads + two stores and keeps 2k registers free (higher occupancy).  Order is lexicographic
// (d, idx); the root is the current k-th best and is mirrored in registers.
template <int KMAX>
struct SmemHeap {
    float *d;   // [KMAX][128]
    int *i;
    int K;      // heap size (= k)
    float rd;   // root (the current k-th best), mirrored in registers
    int ri;
    __device__ __forceinline__ static bool gt(float da, int ia, float db, int ib) { return da > db || (da == db && ia > ib); }
    __device__ __forceinline__ void init(float *dcol, int *icol, int k) {
        d = dcol; i = icol; K = k;
        for (int p = 0; p < k; ++p) { d[p * 128] = INFINITY; i[p * 128] = 0x7fffffff; }
        rd = INFINITY; ri = 0x7fffffff;
    }
    // place (x, s) at `pos` and sift it down
    __device__ __forceinline__ void sift(int pos, float x, int s) {
        while (true) {
            const int l = 2 * pos + 1;
            if (l >= K) break;
            float cd = d[l * 128];
            int ci = i[l * 128], c = l;
            if (l + 1 < K) {
                const float d2 = d[(l + 1) * 128];
                const int i2 = i[(l + 1) * 128];
                if (gt(d2, i2, cd, ci)) { cd = d2; ci = i2; c = l + 1; }
            }
            if (!gt(cd, ci, x, s)) break;
            d[pos * 128] = cd; i[pos * 128] = ci;
            pos = c;
        }
        d[pos * 128] = x; i[pos * 128] = s;
    }
    __device__ __forceinline__ void heapify() {
        for (int p = K / 2 - 1; p >= 0; --p) sift(p, d[p * 128], i[p * 128]);
        rd = d[0]; ri = i[0];
    }
    // precondition: (x, s) < root.  Replace the root and sift down.
    __device__ __forceinline__ void replace_root(float x, int s) {
        sift(0, x, s);
        rd = d[0]; ri = i[0];
    }
    // remove and return the root (largest); heap shrinks by one
    __device__ __forceinline__ void pop(float &od, int &oi) {
        od = d[0]; oi = i[0];
        --K;
        if (K > 0) sift(0, d[K * 128], i[K * 128]);
    }
};

// zig-zag offset sequence 0, -1, +1, -2, +2, ...: the rows of a search box are visited centre first, so that the k-th
// distance is tight before the far rows come up - most of which are then skipped by the row bound below
__device__ __forceinline__ int zigzag(int i) { return (i & 1) ? -((i + 1) >> 1) : (i >> 1); }

// conservative distance from coordinate v to the slab of cell c along one axis (0 inside): every point of the cell is
// at least this far from v along the axis.  cell_of() rounds, so the slab is taken 2*slack wider on both sides.
__device__ __forceinline__ float slab_gap(float v, int c, float lo, float h, float slack) {
    const float a = __fmaf_rn((float)c, h, lo), b = __fmaf_rn((float)(c + 1), h, lo);
    const float g = fmaxf(fmaxf(a - v, v - b), 0.f);
    return fmaxf(__fmul_rn(g, 0.99999f) - 2.f * slack, 0.f);
}

template <int KMAX>
__global__ __launch_bounds__(128) void knn_grid_query_kernel(KnnGridQueryParams P) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.Nq) return;
    const KnnGridHeader H = P.hdr[b];
    const int *__restrict__ cs = P.cell_start + (size_t)b * (P.gmax + 1);
    const float4 *__restrict__ S = P.sorted + (size_t)b * P.Ns;

    // query t of this block: in cell order of the query grid when one is given, else in natural order
    float qx, qy, qz;
    int qi;
    if (P.q_sorted != nullptr) {
        float4 q = P.q_sorted[(size_t)b * P.q_sorted_bs + t];
        qx = q.x; qy = q.y; qz = q.z; qi = __float_as_int(q.w);
    } else {
        const float *qp = P.query + (size_t)b * P.qry_bs + (size_t)t * P.qry_stride;
        qx = qp[0]; qy = qp[1]; qz = qp[2]; qi = t;
    }

    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int p = 0; p < KMAX; ++p) { bd[p] = INFINITY; bi[p] = 0x7fffffff; }

    int px0 = 1, px1 = 0, py0 = 1, py1 = 0, pz0 = 1, pz1 = 0;  // cells already scanned (empty box)
    float R = P.r0_cells * H.h;
    const bool q_ok = (qx == qx) && (qy == qy) && (qz == qz);   // NaN query: nothing compares, scan everything once
    const int cyq = cell_of(qy, H.lo[1], H.inv_h, H.gy), czq = cell_of(qz, H.lo[2], H.inv_h, H.gz);
    for (int pass = 0; pass < 64; ++pass) {
        const float Rb = __fmaf_rn(R, 1.0001f, H.slack);
        int x0 = cell_of(qx - Rb, H.lo[0], H.inv_h, H.gx), x1 = cell_of(qx + Rb, H.lo[0], H.inv_h, H.gx);
        int y0 = cell_of(qy - Rb, H.lo[1], H.inv_h, H.gy), y1 = cell_of(qy + Rb, H.lo[1], H.inv_h, H.gy);
        int z0 = cell_of(qz - Rb, H.lo[2], H.inv_h, H.gz), z1 = cell_of(qz + Rb, H.lo[2], H.inv_h, H.gz);
        if (!q_ok || !(Rb < INFINITY)) { x0 = 0; x1 = H.gx - 1; y0 = 0; y1 = H.gy - 1; z0 = 0; z1 = H.gz - 1; }
        // never shrink (R only grows, but keep the invariant explicit)
        if (px0 <= px1) { x0 = min(x0, px0); x1 = max(x1, px1); y0 = min(y0, py0); y1 = max(y1, py1); z0 = min(z0, pz0); z1 = max(z1, pz1); }
        // rows centre first, rows provably farther than the current k-th distance skipped (see the heap kernel)
        const int nz = 2 * max(czq - z0, z1 - czq) + 1, ny = 2 * max(cyq - y0, y1 - cyq) + 1;
        for (int iz = 0; iz < nz; ++iz) {
            const int z = czq + zigzag(iz);
            if (z < z0 || z > z1) continue;
            const float gz = q_ok ? slab_gap(qz, z, H.lo[2], H.h, H.slack) : 0.f;
            const float gz2 = __fmul_rn(gz, gz);
            if (gz2 > bd[KMAX - 1]) continue;
            for (int iy = 0; iy < ny; ++iy) {
                const int y = cyq + zigzag(iy);
                if (y < y0 || y > y1) continue;
                const float gy = q_ok ? slab_gap(qy, y, H.lo[1], H.h, H.slack) : 0.f;
                const float gyz2 = __fmul_rn(__fmaf_rn(gy, gy, gz2), 0.99999f);
                if (gyz2 > bd[KMAX - 1]) continue;
                const int row = (z * H.gy + y) * H.gx;
                const bool inner = (px0 <= px1) && y >= py0 && y <= py1 && z >= pz0 && z <= pz1;
                // run A: [x0, inner ? px0-1 : x1]   run B: inner ? [px1+1, x1] : empty
                int a0 = x0, a1 = inner ? px0 - 1 : x1;
                int b0 = inner ? px1 + 1 : 1, b1 = inner ? x1 : 0;
                if (q_ok && bd[KMAX - 1] < INFINITY) {   // along x only |dx| <= sqrt(kth - gyz2) can still matter
                    const float ex = __fmaf_rn(sqrtf(fmaxf(bd[KMAX - 1] - gyz2, 0.f)), 1.00001f, 2.f * H.slack);
                    const int xa = cell_of(qx - ex, H.lo[0], H.inv_h, H.gx), xb = cell_of(qx + ex, H.lo[0], H.inv_h, H.gx);
                    a0 = max(a0, xa); a1 = min(a1, xb); b0 = max(b0, xa); b1 = min(b1, xb);
                }
#pragma unroll 1
                for (int run = 0; run < 2; ++run) {
                    const int r0 = run == 0 ? a0 : b0, r1 = run == 0 ? a1 : b1;
                    if (r0 > r1) continue;
                    const int s = cs[row + r0], e = cs[row + r1 + 1];
                    // four candidates per trip: four independent load->distance chains hide the load latency
                    for (int i = s; i < e; i += 4) {
                        float d4[4];
                        int p4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool ok = i + u < e;
                            const float4 p = S[ok ? i + u : s];
                            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                            float d = __fmul_rn(dx, dx);
                            d = __fmaf_rn(dy, dy, d);
                            d = __fmaf_rn(dz, dz, d);
                            d4[u] = ok ? d : INFINITY;
                            p4[u] = ok ? __float_as_int(p.w) : 0x7fffffff;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (d4[u] < bd[KMAX - 1] || (d4[u] == bd[KMAX - 1] && p4[u] < bi[KMAX - 1]))
                                topk_insert_lex<KMAX>(bd, bi, d4[u], p4[u]);
                    }
                }
            }
        }
        px0 = x0; px1 = x1; py0 = y0; py1 = y1; pz0 = z0; pz1 = z1;
        const bool all = x0 == 0 && y0 == 0 && z0 == 0 && x1 == H.gx - 1 && y1 == H.gy - 1 && z1 == H.gz - 1;
        const float kth = bd[P.k - 1 < KMAX ? P.k - 1 : KMAX - 1];
        if (all) break;
        if (kth < __fmul_rn(R, R)) break;   // strict: every unscanned point is farther than R along some axis
        if (kth < INFINITY) {
            // k points known: the k-th distance bounds the answer; make R^2 strictly exceed it
            float Rn = fmaxf(__fmul_rn(sqrtf(kth), 1.000001f), 1e-18f);
            while (!(kth < __fmul_rn(Rn, Rn))) Rn = __fmul_rn(Rn, 1.0001f);
            R = fmaxf(Rn, __fmul_rn(R, 1.0001f));
        } else {
            R = __fmul_rn(R, 2.f);
        }
    }

    int64_t *o = P.idx + (size_t)b * P.idx_bs + (size_t)qi * P.k;
#pragma unroll
    for (int p = 0; p < KMAX; ++p)
        if (p < P.k) o[p] = bi[p] == 0x7fffffff ? (int64_t)-1 : (int64_t)bi[p];
    if (P.idx2 != nullptr && qi < P.idx2_rows) {
        int64_t *o2 = P.idx2 + (size_t)b * P.idx2_bs + (size_t)qi * P.k;
#pragma unroll
        for (int p = 0; p < KMAX; ++p)
            if (p < P.k) o2[p] = bi[p] == 0x7fffffff ? (int64_t)-1 : (int64_t)bi[p];
    }
    if (P.dist2 != nullptr) {
        float *od = P.dist2 + (size_t)b * P.idx_bs + (size_t)qi * P.k;
#pragma unroll
        for (int p = 0; p < KMAX; ++p)
            if (p < P.k) od[p] = bd[p];
    }
}

template <int KMAX>
__global__ __launch_bounds__(128) void knn_grid_query_heap_kernel(KnnGridQueryParams P) {
    __shared__ float s_hd[KMAX * 128];
    __shared__ int s_hi[KMAX * 128];
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.Nq) return;
    const KnnGridHeader H = P.hdr[b];
    const int *__restrict__ cs = P.cell_start + (size_t)b * (P.gmax + 1);
    const float4 *__restrict__ S = P.sorted + (size_t)b * P.Ns;

    // query t of this block: in cell order of the query grid when one is given, else in natural order
    float qx, qy, qz;
    int qi;
    if (P.q_sorted != nullptr) {
        float4 q = P.q_sorted[(size_t)b * P.q_sorted_bs + t];
        qx = q.x; qy = q.y; qz = q.z; qi = __float_as_int(q.w);
    } else {
        const float *qp = P.query + (size_t)b * P.qry_bs + (size_t)t * P.qry_stride;
        qx = qp[0]; qy = qp[1]; qz = qp[2]; qi = t;
    }

    const bool q_ok = (qx == qx) && (qy == qy) && (qz == qz);   // NaN query: nothing compares, scan everything once
    const int cyq = cell_of(qy, H.lo[1], H.inv_h, H.gy), czq = cell_of(qz, H.lo[2], H.inv_h, H.gz);

    // SEED: the heap starts with k real support points - the k sorted-order neighbours of the query's own position
    // (same / adjacent cells) - stored unordered and heapified once.  Filling an empty heap through k root replacements
    // costs k full-depth sifts, more than half of all sift work of a query; and a tight k-th distance from the start makes
    // the row bound effective at once.  The seeded positions [p0, p0 + k) are skipped by the scan.
    SmemHeap<KMAX> hp;
    hp.d = s_hd + threadIdx.x; hp.i = s_hi + threadIdx.x; hp.K = P.k;
    int p0;
    {
        int pos = t;                                   // self-kNN in cell order: the query is support point t
        if (P.q_sorted + (size_t)b * P.q_sorted_bs != S) {   // not a self query (the query IS support point t otherwise)
            const int cxq = cell_of(qx, H.lo[0], H.inv_h, H.gx);
            pos = cs[(czq * H.gy + cyq) * H.gx + cxq];
        }
        p0 = min(max(pos - (P.k >> 1), 0), P.Ns - P.k);
        for (int j = 0; j < P.k; ++j) {
            const float4 p = S[p0 + j];
            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
            float d = __fmul_rn(dx, dx);
            d = __fmaf_rn(dy, dy, d);
            d = __fmaf_rn(dz, dz, d);
            const bool nan = d != d;                   // never a candidate (like the scan's comparison): placeholder
            hp.d[j * 128] = nan ? INFINITY : d;
            hp.i[j * 128] = nan ? 0x7fffffff : __float_as_int(p.w);
        }
        hp.heapify();
    }

    int px0 = 1, px1 = 0, py0 = 1, py1 = 0, pz0 = 1, pz1 = 0;  // cells already scanned (empty box)
    float R = P.r0_cells * H.h;
    for (int pass = 0; pass < 64; ++pass) {
        const float Rb = __fmaf_rn(R, 1.0001f, H.slack);
        int x0 = cell_of(qx - Rb, H.lo[0], H.inv_h, H.gx), x1 = cell_of(qx + Rb, H.lo[0], H.inv_h, H.gx);
        int y0 = cell_of(qy - Rb, H.lo[1], H.inv_h, H.gy), y1 = cell_of(qy + Rb, H.lo[1], H.inv_h, H.gy);
        int z0 = cell_of(qz - Rb, H.lo[2], H.inv_h, H.gz), z1 = cell_of(qz + Rb, H.lo[2], H.inv_h, H.gz);
        if (!q_ok || !(Rb < INFINITY)) { x0 = 0; x1 = H.gx - 1; y0 = 0; y1 = H.gy - 1; z0 = 0; z1 = H.gz - 1; }
        // never shrink (R only grows, but keep the invariant explicit)
        if (px0 <= px1) { x0 = min(x0, px0); x1 = max(x1, px1); y0 = min(y0, py0); y1 = max(y1, py1); z0 = min(z0, pz0); z1 = max(z1, pz1); }
        // rows (z, y) of the box, centre first.  A row whose slab is provably farther than the current k-th distance
        // cannot contribute now or later (the k-th distance only shrinks), so it counts as scanned without being read.
        const int nz = 2 * max(czq - z0, z1 - czq) + 1, ny = 2 * max(cyq - y0, y1 - cyq) + 1;
        for (int iz = 0; iz < nz; ++iz) {
            const int z = czq + zigzag(iz);
            if (z < z0 || z > z1) continue;
            const float gz = q_ok ? slab_gap(qz, z, H.lo[2], H.h, H.slack) : 0.f;
            const float gz2 = __fmul_rn(gz, gz);
            if (gz2 > hp.rd) continue;
            for (int iy = 0; iy < ny; ++iy) {
                const int y = cyq + zigzag(iy);
                if (y < y0 || y > y1) continue;
                const float gy = q_ok ? slab_gap(qy, y, H.lo[1], H.h, H.slack) : 0.f;
                const float gyz2 = __fmul_rn(__fmaf_rn(gy, gy, gz2), 0.99999f);   // lower bound of every d2 in the row
                if (gyz2 > hp.rd) continue;
                const int row = (z * H.gy + y) * H.gx;
                const bool inner = (px0 <= px1) && y >= py0 && y <= py1 && z >= pz0 && z <= pz1;
                // run A: [x0, inner ? px0-1 : x1]   run B: inner ? [px1+1, x1] : empty
                int a0 = x0, a1 = inner ? px0 - 1 : x1;
                int b0 = inner ? px1 + 1 : 1, b1 = inner ? x1 : 0;
                if (q_ok && hp.rd < INFINITY) {
                    // along x only |dx| <= sqrt(kth - gyz2) can still matter: trim both runs to those cells
                    const float ex = __fmaf_rn(sqrtf(fmaxf(hp.rd - gyz2, 0.f)), 1.00001f, 2.f * H.slack);
                    const int xa = cell_of(qx - ex, H.lo[0], H.inv_h, H.gx), xb = cell_of(qx + ex, H.lo[0], H.inv_h, H.gx);
                    a0 = max(a0, xa); a1 = min(a1, xb); b0 = max(b0, xa); b1 = min(b1, xb);
                }
#pragma unroll 1
                for (int run = 0; run < 2; ++run) {
                    const int r0 = run == 0 ? a0 : b0, r1 = run == 0 ? a1 : b1;
                    if (r0 > r1) continue;
                    const int s = cs[row + r0], e = cs[row + r1 + 1];
                    // four candidates per trip: four independent load->distance chains hide the load latency
                    for (int i = s; i < e; i += 4) {
                        float d4[4];
                        int p4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool ok = i + u < e && (unsigned)(i + u - p0) >= (unsigned)P.k;   // in range, not a seed
                            const float4 p = S[ok ? i + u : s];
                            const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                            float d = __fmul_rn(dx, dx);
                            d = __fmaf_rn(dy, dy, d);
                            d = __fmaf_rn(dz, dz, d);
                            d4[u] = ok ? d : INFINITY;
                            p4[u] = ok ? __float_as_int(p.w) : 0x7fffffff;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (d4[u] < hp.rd || (d4[u] == hp.rd && p4[u] < hp.ri)) hp.replace_root(d4[u], p4[u]);
                    }
                }
            }
        }
        px0 = x0; px1 = x1; py0 = y0; py1 = y1; pz0 = z0; pz1 = z1;
        const bool all = x0 == 0 && y0 == 0 && z0 == 0 && x1 == H.gx - 1 && y1 == H.gy - 1 && z1 == H.gz - 1;
        const float kth = hp.rd;
        if (all) break;
        if (kth < __fmul_rn(R, R)) break;   // strict: every unscanned point is farther than R along some axis
        if (kth < INFINITY) {
            // k points known: the k-th distance bounds the answer; make R^2 strictly exceed it
            float Rn = fmaxf(__fmul_rn(sqrtf(kth), 1.000001f), 1e-18f);
            while (!(kth < __fmul_rn(Rn, Rn))) Rn = __fmul_rn(Rn, 1.0001f);
            // (the seeds can make the k-th distance finite long before k near points are known: never more than double)
            R = fmaxf(fminf(Rn, __fmul_rn(R, 2.f)), __fmul_rn(R, 1.0001f));
        } else {
            R = __fmul_rn(R, 2.f);
        }
    }

    int64_t *o = P.idx + (size_t)b * P.idx_bs + (size_t)qi * P.k;
    int64_t *o2 = (P.idx2 != nullptr && qi < P.idx2_rows) ? P.idx2 + (size_t)b * P.idx2_bs + (size_t)qi * P.k : nullptr;
    float *od = P.dist2 != nullptr ? P.dist2 + (size_t)b * P.idx_bs + (size_t)qi * P.k : nullptr;
    {
        for (int p = P.k - 1; p >= 0; --p) {     // the heap yields the k best in descending order
            float dd;
            int ii;
            hp.pop(dd, ii);
            const int64_t v = ii == 0x7fffffff ? (int64_t)-1 : (int64_t)ii;
            o[p] = v;
            if (o2) o2[p] = v;
            if (od) od[p] = dd;
        }
    }
}

}  // namespace

size_t knn_grid_smem_bytes(int gmax) { return (size_t)(gmax + 1) * sizeof(int); }

int launch_knn_grid_build(const KnnGridBuildParams &P, int ngrids, int B, cudaStream_t st) {
    if (ngrids <= 0 || B <= 0) return DSIR_OK;
    const size_t smem = knn_grid_smem_bytes(P.gmax);
    DSIR_CUDA_TRY(cudaFuncSetAttribute(knn_grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ngrids, B);
    knn_grid_build_kernel<<<grid, GRID_BUILD_THREADS, smem, st>>>(P);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

int launch_knn_grid_query(const KnnGridQueryParams &P, int B, cudaStream_t st) {
    if (P.k < 1 || P.k > 32) return DSIR_ERR_UNSUPPORTED;
    if (P.Ns < P.k) return DSIR_ERR_KNN_TOO_FEW;
    if (P.Nq <= 0 || B <= 0) return DSIR_OK;
    dim3 grid((P.Nq + 127) / 128, B);
    if (P.k == 1) knn_grid_query_kernel<1><<<grid, 128, 0, st>>>(P);
    else if (P.k <= 4) knn_grid_query_kernel<4><<<grid, 128, 0, st>>>(P);
    else if (P.k <= 8) knn_grid_query_heap_kernel<8><<<grid, 128, 0, st>>>(P);
    else if (P.k <= 16) knn_grid_query_heap_kernel<16><<<grid, 128, 0, st>>>(P);
    else knn_grid_query_heap_kernel<32><<<grid, 128, 0, st>>>(P);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
