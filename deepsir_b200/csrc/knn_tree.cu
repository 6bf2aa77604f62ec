// Exact xyz k-nearest neighbours on a bucket tree of kd-ordered leaves — the warp-cooperative path for level
// clouds of up to 17408 points (dataloader/data_base.py:165,170 call sites; contract of knn.cu / oracle/knn_oracle.c):
//     d2 = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32, results ascending in (d2, support index).
//
//   build   one CTA per (cloud, tree): kd-ordered LEAVES of 32 points (level-by-level counting sorts in shared memory
//           along each segment's widest axis, 8 / 4-way splits at leaf multiples), stored as one 512-byte
//           structure-of-arrays block {-x[32], -y[32], -z[32], index[32]} -> bounding box per leaf and per SUPERNODE
//           (32 consecutive leaves).  A cloud of 16384 points is 512 leaves / 16 supernodes.
//   query   one WARP per query leaf, lane = query.  The 32 queries of a leaf are spatial neighbours, so they share one
//           candidate set: the warp walks supernodes and leaves nearest-first (lane = child: 32 box tests per instruction
//           sequence, no stack, no divergence), an elected lane stages each chosen leaf into shared memory with ONE
//           512-byte cp.async.bulk (TMA engine, mbarrier ring, the next leaf in flight while the current one is scanned),
//           and every lane scans the staged leaf with broadcast LDS.128 and packed f32x2 arithmetic (~5 instructions per
//           candidate instead of ~70 in the per-thread cell walk of knn_grid.cu).  Candidates that beat the lane's
//           current k-th distance are appended to a per-lane queue; the queues are drained into per-lane max-heaps
//           (shared memory, column layout) in batches, so that the insertion code runs with most lanes active.
//
// Exactness does not depend on the tree.  A leaf (or supernode) is skipped only when a LOWER BOUND of every distance
// into its box exceeds the k-th distance: the bound is evaluated with the same operation order as d2 on per-axis gaps
// g <= |q - p| (gap = max(lo - q, q - hi, 0); for the coarse test, box to box).  fp32 subtraction, multiplication and
// fma are monotone in each argument, hence bound <= d2 holds for the COMPUTED values, without any rounding slack; a
// leaf with bound == k-th distance is still scanned (an equal distance with a lower index would win).
#include "knn.cuh"

namespace dsir {

namespace {

constexpr int TREE_BUILD_THREADS = 1024;
#ifndef DSIR_TREE_QWARPS
#define DSIR_TREE_QWARPS 4
#endif
#ifndef DSIR_TREE_CHK
#define DSIR_TREE_CHK 4
#endif
constexpr int TREE_QWARPS = DSIR_TREE_QWARPS;     // query leaves (warps) per CTA
constexpr int TREE_CHK = DSIR_TREE_CHK;           // candidates between two looks at the buffers (4 or 8)
constexpr int TREE_TRIG = 16 - TREE_CHK;          // flush when a lane has more new entries than this (<= 16 new at the flush)
constexpr int TREE_NST = 2;        // staging ring depth per warp
constexpr int TREE_QCAP = 32;      // pair-buffer entries per lane: <= 16 kept + <= 16 new
constexpr unsigned FULL = 0xffffffffu;
constexpr int NOIDX = 0x7fffffff;

#ifdef DSIR_KNN_STATS   // development counters (never compiled into the shipped library)
__device__ unsigned long long g_tree_stats[8];
#define TREE_STAT(i, v) do { const unsigned long long v_ = (unsigned long long)(v); if (lane == 0) atomicAdd(&g_tree_stats[i], v_); } while (0)
#else
#define TREE_STAT(i, v) do { } while (0)
#endif

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }

// lower bound of d2 between two boxes / a point and a box, in the operation order of d2 (see the header comment)
__device__ __forceinline__ float gap1(float lo, float hi, float qlo, float qhi) {
    return fmaxf(fmaxf(__fsub_rn(lo, qhi), __fsub_rn(qlo, hi)), 0.f);
}
__device__ __forceinline__ float gap2_3(float gx, float gy, float gz) {
    float d = __fmul_rn(gx, gx);
    d = __fmaf_rn(gy, gy, d);
    d = __fmaf_rn(gz, gz, d);
    return d;
}

// ---------------------------------------------------------------------------------------------------------------
// build: kd-ordered leaves
//   The leaf order decides how many leaves the 32 queries of a warp have to scan: Morton order needs 24 leaves per warp on
//   the C2 cloud, median splits along the widest axis need 12.5 (tools/knn_leaf_sim.py).  The build therefore sorts level
//   by level: every segment (a contiguous range of leaves) is counting-sorted along ITS OWN widest axis with an 8-bit key
//   re-quantised to its own extent, and cut into 8 (first level, when the depth is odd) or 4 children at leaf multiples -
//   four levels for 512 leaves instead of nine binary ones, with the leaf quality of exact medians (the cut falls inside
//   one of 256 key bins).  Coordinates are kept as 16-bit quantised copies in shared memory; the quantisation only shapes
//   the leaves, the search is exact for any order.
// ---------------------------------------------------------------------------------------------------------------
// lanes with the same 8-bit key (among the lanes flagged valid)
__device__ __forceinline__ unsigned peers8(unsigned key, bool valid) {
    unsigned m = __ballot_sync(FULL, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool on = (key >> bit) & 1u;
        const unsigned bal = __ballot_sync(FULL, on);
        m &= on ? bal : ~bal;
    }
    return m;
}

struct BuildSmem {
    unsigned short *q[3];     // [cap] 16-bit quantised coordinates, indexed by original point index
    unsigned short *pin, *pout;   // [cap] permutation (ping / pong)
    unsigned short *hist;     // [32 warps][256]
    unsigned short *tmp;      // [cap] per position: key | rank among equal keys of the 32-block << 8 | last of its group << 13
};

// 8-bit key of element e along axis `qa`, relative to [base, base + ext]
__device__ __forceinline__ unsigned key8(const unsigned short *qa, unsigned e, int base, float mul) {
    return (unsigned)min(255, __float2int_rz((float)((int)qa[e] - base) * mul));
}

// one warp sorts positions [p0, p1) of pin into pout along the segment's widest axis (no block-level synchronisation)
__device__ void warp_sort_segment(const BuildSmem &B, int p0, int p1, const float *unit, unsigned short *wh, int lane) {
    int lo[3] = {65535, 65535, 65535}, hi[3] = {0, 0, 0};
    for (int pos = p0 + lane; pos < p1; pos += 32) {
        const unsigned e = B.pin[pos];
#pragma unroll
        for (int a = 0; a < 3; ++a) { const int v = B.q[a][e]; lo[a] = min(lo[a], v); hi[a] = max(hi[a], v); }
    }
    float best = -1.f;
    int ax = 0, base = 0, ext = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int l = __reduce_min_sync(FULL, lo[a]), h = __reduce_max_sync(FULL, hi[a]);
        const float w = (float)(h - l) * unit[a];
        if (w > best) { best = w; ax = a; base = l; ext = h - l; }
    }
    const unsigned short *qa = B.q[ax];
    const float mul = ext > 0 ? 255.99f / (float)ext : 0.f;
    for (int i = lane; i < 256; i += 32) wh[i] = 0;
    __syncwarp();
    for (int pos0 = p0; pos0 < p1; pos0 += 32) {
        const int pos = pos0 + lane;
        const bool v = pos < p1;
        const unsigned e = v ? B.pin[pos] : 0u;
        const unsigned k = v ? key8(qa, e, base, mul) : 0u;
        const unsigned m = peers8(k, v);
        const int rank = __popc(m & ((1u << lane) - 1u)), cnt = __popc(m);
        if (v && rank == 0) wh[k] = (unsigned short)(wh[k] + cnt);
        if (v) B.tmp[pos] = (unsigned short)(k | (rank << 8) | ((rank == cnt - 1) << 13));
        __syncwarp();
    }
    {   // exclusive scan of the 256 bins: lane owns bins 8 lane .. 8 lane + 7
        int loc[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { loc[i] = wh[8 * lane + i]; sum += loc[i]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += u;
        }
        int run = incl - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) { wh[8 * lane + i] = (unsigned short)run; run += loc[i]; }
    }
    __syncwarp();
    for (int pos0 = p0; pos0 < p1; pos0 += 32) {   // scatter: key / rank / group end come from the count pass
        const int pos = pos0 + lane;
        const bool v = pos < p1;
        const unsigned t = v ? B.tmp[pos] : 0u;
        const unsigned k = t & 255u;
        const int rank = (t >> 8) & 31;
        const int off = v ? (int)wh[k] : 0;
        __syncwarp();
        if (v && (t >> 13)) wh[k] = (unsigned short)(off + rank + 1);
        __syncwarp();
        if (v) B.pout[p0 + off + rank] = B.pin[pos];
    }
}

__global__ __launch_bounds__(TREE_BUILD_THREADS) void knn_tree_build_kernel(KnnTreeBuildParams P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ float s_red[6][32];
    __shared__ float s_box[8];
    __shared__ int s_wtot[32];
    __shared__ int s_seg[8];                 // CTA-wide sort of one segment: axis, base, ext
    __shared__ int s_ired[6][32];
    __shared__ unsigned short s_b0[514], s_b1[514];   // segment boundaries in leaves (current / next level)

    const int g = blockIdx.x, b = blockIdx.y;
    const int n = P.n[g];
    const int nleaf = (n + 31) >> 5, nsuper = (nleaf + 31) >> 5, nlpad = nsuper * 32;
    const float4 *pts = P.pts4 + (size_t)b * P.pts_bs;
    BuildSmem B;
    B.q[0] = (unsigned short *)sm_raw;
    B.q[1] = B.q[0] + P.cap;
    B.q[2] = B.q[1] + P.cap;
    B.pin = B.q[2] + P.cap;
    B.pout = B.pin + P.cap;
    B.hist = B.pout + P.cap;
    B.tmp = B.hist + 32 * 256;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- bounding box (NaN coordinates are ignored by fminf / fmaxf) ----
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += TREE_BUILD_THREADS) {
        const float4 p = pts[i];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l = warp_min(lo[a]), h = warp_max(hi[a]);
        if (lane == 0) { s_red[a][warp] = l; s_red[3 + a][warp] = h; }
    }
    __syncthreads();
    if (tid == 0) {
        for (int a = 0; a < 3; ++a) {
            float l = INFINITY, h = -INFINITY;
            for (int w = 0; w < TREE_BUILD_THREADS / 32; ++w) { l = fminf(l, s_red[a][w]); h = fmaxf(h, s_red[3 + a][w]); }
            if (!(l <= h) || !isfinite(l) || !isfinite(h)) { l = 0.f; h = 0.f; }
            const float ext = h - l;
            s_box[a] = l;
            s_box[3 + a] = (ext > 0.f && isfinite(ext)) ? 65535.0f / ext : 0.f;   // quantisation scale of the axis
        }
    }
    __syncthreads();
    float unit[3];                            // length of one quantisation step per axis (to compare extents across axes)
#pragma unroll
    for (int a = 0; a < 3; ++a) unit[a] = s_box[3 + a] > 0.f ? 1.0f / s_box[3 + a] : 0.f;
    for (int i = tid; i < n; i += TREE_BUILD_THREADS) {
        const float4 p = pts[i];
        B.q[0][i] = (unsigned short)min(max(__float2int_rz((p.x - s_box[0]) * s_box[3]), 0), 65535);
        B.q[1][i] = (unsigned short)min(max(__float2int_rz((p.y - s_box[1]) * s_box[4]), 0), 65535);
        B.q[2][i] = (unsigned short)min(max(__float2int_rz((p.z - s_box[2]) * s_box[5]), 0), 65535);
        B.pin[i] = (unsigned short)i;
    }
    if (tid == 0) { s_b0[0] = 0; s_b0[1] = (unsigned short)nleaf; }
    __syncthreads();

    // ---- levels: depth T = floor(log2(nleaf)) binary levels, grouped as [8 if T is odd and >= 3 | 2 if T == 1], 4, 4, ... ----
    int T = 0;
    while ((2 << T) <= nleaf) ++T;
    unsigned short *bc = s_b0, *bn = s_b1;
    int S = 1;
    while (T > 0) {
        const int f = (T == 1) ? 2 : ((T & 1) ? 8 : 4);
        T -= (f == 8) ? 3 : (f == 4 ? 2 : 1);
        if (S < 32) {
            // few, large segments: the whole CTA sorts them one after the other (warp w owns a contiguous range of the
            // segment's 32-blocks; per-warp histograms; exclusive scan in (key-major, warp-minor) order)
            for (int sgm = 0; sgm < S; ++sgm) {
                const int p0 = bc[sgm] * 32, p1 = min(bc[sgm + 1] * 32, n);
                int l3[3] = {65535, 65535, 65535}, h3[3] = {0, 0, 0};
                for (int pos = p0 + tid; pos < p1; pos += TREE_BUILD_THREADS) {
                    const unsigned e = B.pin[pos];
#pragma unroll
                    for (int a = 0; a < 3; ++a) { const int v = B.q[a][e]; l3[a] = min(l3[a], v); h3[a] = max(h3[a], v); }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int l = __reduce_min_sync(FULL, l3[a]), h = __reduce_max_sync(FULL, h3[a]);
                    if (lane == 0) { s_ired[a][warp] = l; s_ired[3 + a][warp] = h; }
                }
                for (int i = tid; i < 32 * 256; i += TREE_BUILD_THREADS) B.hist[i] = 0;
                __syncthreads();
                if (tid == 0) {
                    float best = -1.f;
                    for (int a = 0; a < 3; ++a) {
                        int l = 65535, h = 0;
                        for (int w = 0; w < 32; ++w) { l = min(l, s_ired[a][w]); h = max(h, s_ired[3 + a][w]); }
                        const float wdt = (float)(h - l) * unit[a];
                        if (wdt > best) { best = wdt; s_seg[0] = a; s_seg[1] = l; s_seg[2] = h - l; }
                    }
                }
                __syncthreads();
                const unsigned short *qa = B.q[s_seg[0]];
                const int base = s_seg[1];
                const float mul = s_seg[2] > 0 ? 255.99f / (float)s_seg[2] : 0.f;
                const int nblk = (p1 - p0 + 31) >> 5, bpw = (nblk + 31) >> 5;
                const int blk0 = min(warp * bpw, nblk), blk1 = min(blk0 + bpw, nblk);
                unsigned short *wh = B.hist + warp * 256;
                for (int blk = blk0; blk < blk1; ++blk) {
                    const int pos = p0 + blk * 32 + lane;
                    const bool v = pos < p1;
                    const unsigned e = v ? B.pin[pos] : 0u;
                    const unsigned k = v ? key8(qa, e, base, mul) : 0u;
                    const unsigned m = peers8(k, v);
                    const int rank = __popc(m & ((1u << lane) - 1u)), cnt = __popc(m);
                    if (v && rank == 0) wh[k] = (unsigned short)(wh[k] + cnt);
                    if (v) B.tmp[pos] = (unsigned short)(k | (rank << 8) | ((rank == cnt - 1) << 13));
                    __syncwarp();
                }
                __syncthreads();
                {   // entry e = key * 32 + warp; thread t owns entries 8t .. 8t+7
                    const int dg = tid >> 2, w0 = (tid & 3) * 8;
                    int loc[8], sum = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { loc[i] = B.hist[(w0 + i) * 256 + dg]; sum += loc[i]; }
                    int incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += u;
                    }
                    if (lane == 31) s_wtot[warp] = incl;
                    __syncthreads();
                    if (warp == 0) {
                        const int u = s_wtot[lane];
                        int inc2 = u;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(FULL, inc2, o);
                            if (lane >= o) inc2 += t;
                        }
                        s_wtot[lane] = inc2 - u;
                    }
                    __syncthreads();
                    int run = s_wtot[warp] + incl - sum;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { B.hist[(w0 + i) * 256 + dg] = (unsigned short)run; run += loc[i]; }
                }
                __syncthreads();
                for (int blk = blk0; blk < blk1; ++blk) {
                    const int pos = p0 + blk * 32 + lane;
                    const bool v = pos < p1;
                    const unsigned t = v ? B.tmp[pos] : 0u;
                    const unsigned k = t & 255u;
                    const int rank = (t >> 8) & 31;
                    const int off = v ? (int)wh[k] : 0;
                    __syncwarp();
                    if (v && (t >> 13)) wh[k] = (unsigned short)(off + rank + 1);
                    __syncwarp();
                    if (v) B.pout[p0 + off + rank] = B.pin[pos];
                }
                __syncthreads();
            }
        } else {
            // many small segments: one warp per segment, no block-level synchronisation inside
            for (int sgm = warp; sgm < S; sgm += TREE_BUILD_THREADS / 32) {
                const int p0 = bc[sgm] * 32, p1 = min(bc[sgm + 1] * 32, n);
                if (p1 > p0) warp_sort_segment(B, p0, p1, unit, B.hist + warp * 256, lane);
            }
            __syncthreads();
        }
        // children: segment [a, a + m) -> child c = [a + c m / f, a + (c + 1) m / f)
        for (int i = tid; i < S * f; i += TREE_BUILD_THREADS) {
            const int sgm = i / f, c = i % f;
            const int a0 = bc[sgm], m = bc[sgm + 1] - a0;
            bn[i] = (unsigned short)(a0 + (c * m) / f);
        }
        if (tid == 0) bn[S * f] = (unsigned short)nleaf;
        S *= f;
        __syncthreads();
        { unsigned short *t = bc; bc = bn; bn = t; }
        { unsigned short *t = B.pin; B.pin = B.pout; B.pout = t; }
    }
    const unsigned short *pin = B.pin;

    // ---- leaves + leaf boxes: warp w writes leaf it*32 + w ----
    KnnLeaf *leaves = P.leaves[g] + (size_t)b * nleaf;
    float *box = P.box[g] + (size_t)b * 6 * nlpad;
    for (int lf = warp; lf < nlpad; lf += TREE_BUILD_THREADS / 32) {
        float x = INFINITY, y = INFINITY, z = INFINITY;        // padding: distance +inf to every finite query
        int id = NOIDX;
        const int pos = lf * 32 + lane;
        if (pos < n) {
            id = pin[pos];
            const float4 p = pts[id];
            x = p.x; y = p.y; z = p.z;
        }
        if (lf < nleaf) {
            KnnLeaf *L = leaves + lf;
            L->nx[lane] = -x; L->ny[lane] = -y; L->nz[lane] = -z; L->idx[lane] = id;
        }
        const bool v = pos < n;
        const float lx = warp_min(v ? x : INFINITY), ly = warp_min(v ? y : INFINITY), lz = warp_min(v ? z : INFINITY);
        const float hx = warp_max(v ? x : -INFINITY), hy = warp_max(v ? y : -INFINITY), hz = warp_max(v ? z : -INFINITY);
        if (lane == 0) {
            box[0 * nlpad + lf] = lx; box[1 * nlpad + lf] = ly; box[2 * nlpad + lf] = lz;
            box[3 * nlpad + lf] = hx; box[4 * nlpad + lf] = hy; box[5 * nlpad + lf] = hz;
        }
    }
    __syncthreads();   // the leaf boxes written above (global memory) are visible to the whole CTA
    // ---- supernode boxes: warp s reduces the 32 leaf boxes of supernode s; unused supernodes are empty boxes ----
    float *sbox = P.sbox[g] + (size_t)b * 6 * 32;
    {
        const int s = warp;
        float v[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) v[a] = a < 3 ? INFINITY : -INFINITY;
        if (s < nsuper) {
#pragma unroll
            for (int a = 0; a < 6; ++a) v[a] = box[a * nlpad + s * 32 + lane];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { v[a] = warp_min(v[a]); v[3 + a] = warp_max(v[3 + a]); }
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 6; ++a) sbox[a * 32 + s] = v[a];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// query
// ---------------------------------------------------------------------------------------------------------------
// Selection.  Every lane keeps (1) the 16 smallest DISTANCES seen so far as a sorted list in registers - only values, so a
// compare-exchange is two FMNMX - whose last entry is the lane's k-th distance, and (2) the (d, idx) pairs that can still
// belong to the answer, UNSORTED, in a per-lane buffer in shared memory (column layout [slot][lane]: conflict free).
// Candidates with d <= k-th distance are appended to the buffer; when some lane has more than 8 new entries, ALL lanes
// merge their (<= 16) new values into the register list with sorting networks (Batcher 16 + bitonic merge: branch-free,
// the same instruction stream whether one lane or all have work) and compact their buffers to the entries with
// d <= new k-th distance; more than k survivors means ties AT the k-th distance, resolved by index in a rare slow loop.
// After the last leaf the <= 16 surviving pairs are sorted once, lexicographically.  (Earlier versions: per-candidate
// shared-memory heaps - data-dependent sifts at ~35 % lane efficiency, 2/3 of the kernel's instructions; sorted (d, idx)
// register lists merged by pair networks - 11 instructions per compare-exchange, 43 % of the instructions.)
constexpr int TREE_K = 16;   // list length; clouds with k > 16 are served by the grid path

struct Pair { float d; int i; };
// (a, b) -> (min, max) in lexicographic (d, idx) order
__device__ __forceinline__ void ce(float &ad, int &ai, float &bd, int &bi) {
    const bool sw = ad > bd || (ad == bd && ai > bi);
    const float td = sw ? bd : ad, ud = sw ? ad : bd;
    const int ti = sw ? bi : ai, ui = sw ? ai : bi;
    ad = td; ai = ti; bd = ud; bi = ui;
}
// Batcher's odd-even merge sort, 16 inputs, 63 compare-exchanges, ascending (network checked with the 0-1 principle)
__device__ __forceinline__ void sort16(float (&d)[16], int (&i)[16]) {
#define CE(a, b) ce(d[a], i[a], d[b], i[b])
    CE(0, 1); CE(2, 3); CE(4, 5); CE(6, 7); CE(8, 9); CE(10, 11); CE(12, 13); CE(14, 15);
    CE(0, 2); CE(1, 3); CE(4, 6); CE(5, 7); CE(8, 10); CE(9, 11); CE(12, 14); CE(13, 15);
    CE(1, 2); CE(5, 6); CE(9, 10); CE(13, 14); CE(0, 4); CE(1, 5); CE(2, 6); CE(3, 7);
    CE(8, 12); CE(9, 13); CE(10, 14); CE(11, 15); CE(2, 4); CE(3, 5); CE(10, 12); CE(11, 13);
    CE(1, 2); CE(3, 4); CE(5, 6); CE(9, 10); CE(11, 12); CE(13, 14); CE(0, 8); CE(1, 9);
    CE(2, 10); CE(3, 11); CE(4, 12); CE(5, 13); CE(6, 14); CE(7, 15); CE(4, 8); CE(5, 9);
    CE(6, 10); CE(7, 11); CE(2, 4); CE(3, 5); CE(6, 8); CE(7, 9); CE(10, 12); CE(11, 13);
    CE(1, 2); CE(3, 4); CE(5, 6); CE(7, 8); CE(9, 10); CE(11, 12); CE(13, 14);
#undef CE
}
// the same network on values only (two FMNMX per compare-exchange)
__device__ __forceinline__ void sort16v(float (&d)[16]) {
#define CE(a, b) { const float lo_ = fminf(d[a], d[b]); d[b] = fmaxf(d[a], d[b]); d[a] = lo_; }
    CE(0, 1); CE(2, 3); CE(4, 5); CE(6, 7); CE(8, 9); CE(10, 11); CE(12, 13); CE(14, 15);
    CE(0, 2); CE(1, 3); CE(4, 6); CE(5, 7); CE(8, 10); CE(9, 11); CE(12, 14); CE(13, 15);
    CE(1, 2); CE(5, 6); CE(9, 10); CE(13, 14); CE(0, 4); CE(1, 5); CE(2, 6); CE(3, 7);
    CE(8, 12); CE(9, 13); CE(10, 14); CE(11, 15); CE(2, 4); CE(3, 5); CE(10, 12); CE(11, 13);
    CE(1, 2); CE(3, 4); CE(5, 6); CE(9, 10); CE(11, 12); CE(13, 14); CE(0, 8); CE(1, 9);
    CE(2, 10); CE(3, 11); CE(4, 12); CE(5, 13); CE(6, 14); CE(7, 15); CE(4, 8); CE(5, 9);
    CE(6, 10); CE(7, 11); CE(2, 4); CE(3, 5); CE(6, 8); CE(7, 9); CE(10, 12); CE(11, 13);
    CE(1, 2); CE(3, 4); CE(5, 6); CE(7, 8); CE(9, 10); CE(11, 12); CE(13, 14);
#undef CE
}
// list (ascending) <- the 16 smallest values of list U nw (both ascending), ascending
__device__ __forceinline__ void merge16v(float (&l)[16], const float (&nw)[16]) {
#pragma unroll
    for (int a = 0; a < 16; ++a) l[a] = fminf(l[a], nw[15 - a]);   // lower half of the bitonic sequence (list, reversed nw)
#pragma unroll
    for (int k = 8; k >= 1; k >>= 1)
#pragma unroll
        for (int a = 0; a < 16; ++a)
            if ((a & k) == 0) { const float lo_ = fminf(l[a], l[a + k]); l[a + k] = fmaxf(l[a], l[a + k]); l[a] = lo_; }
}

__host__ __device__ constexpr int tree_warp_smem(bool k1) {
    // stages + mbarriers (padded to 128) + leaf-box cache [6][32] + queue
    return TREE_NST * 512 + 128 + 768 + (k1 ? 0 : TREE_QCAP * 32 * 6);   // pair buffer: fp32 distance + 16-bit index
}

template <bool K1>
__global__ __launch_bounds__(TREE_QWARPS * 32) void knn_tree_query_kernel(KnnTreeQueryParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int ql = blockIdx.x * TREE_QWARPS + warp;
    if (ql >= P.qry.nleaf) return;                           // whole warps leave; nothing below syncs across warps

    unsigned char *wb = smem + (size_t)warp * tree_warp_smem(K1);
    float *stage = (float *)wb;                               // [NST][128 floats]
    uint64_t *bar = (uint64_t *)(wb + TREE_NST * 512);
    float *bxc = (float *)(wb + TREE_NST * 512 + 128);        // boxes of the current supernode's leaves [6][32]
    float *qd = (float *)(wb + TREE_NST * 512 + 128 + 768) + lane;   // queue columns of this lane: d, then idx
    // indices as 16 bits (tree clouds hold <= 17408 points; 0xffff marks the padding slots of the last leaf): 6 instead of
    // 8 bytes per entry is two more CTAs per SM
    unsigned short *qi = (unsigned short *)((float *)(wb + TREE_NST * 512 + 128 + 768) + TREE_QCAP * 32) + lane;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < TREE_NST; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // ---- the 32 queries of this leaf ----
    const KnnLeaf *QL = P.qry.leaves + (size_t)b * P.qry.nleaf + ql;
    const float qx = -QL->nx[lane], qy = -QL->ny[lane], qz = -QL->nz[lane];
    const int qidx = QL->idx[lane];
    const bool qvalid = qidx != NOIDX;
    const bool qok = qvalid && qx == qx && qy == qy && qz == qz;   // a NaN query compares with nothing: all slots stay empty
    float qlo[3], qhi[3];
    {
        const float *qb = P.qry.box + (size_t)b * 6 * P.qry.nlpad + ql;
#pragma unroll
        for (int a = 0; a < 3; ++a) { qlo[a] = qb[a * P.qry.nlpad]; qhi[a] = qb[(3 + a) * P.qry.nlpad]; }
    }

    const KnnLeaf *SL = P.sup.leaves + (size_t)b * P.sup.nleaf;
    const float *sbx = P.sup.box + (size_t)b * 6 * P.sup.nlpad;
    const int nleaf = P.sup.nleaf, nlpad = P.sup.nlpad;
    const int k = P.k;

    // ---- selection state ----
    // the 16 smallest distances so far, ascending; for k < 16 the first 16 - k slots hold -inf sentinels that sort before
    // every candidate, so that the k-th distance is always sv[15] (no dynamic register index)
    float sv[TREE_K];
#pragma unroll
    for (int p = 0; p < TREE_K; ++p) sv[p] = p < TREE_K - k ? -INFINITY : INFINITY;
    float bd = INFINITY;                                      // K1: the best pair
    int bi = NOIDX;
    const uint32_t qbase = smem_u32(qd);
    uint32_t qp = qbase;                                      // buffer write cursor of this lane (entry j at qbase + 128 j)
    const uint32_t ibase = smem_u32(qi);
    uint32_t ip = ibase;                                      // ... and of its index column (entry j at ibase + 64 j)
    int kept = 0;                                             // entries [0, kept) survived the last compaction
    // the lane's current k-th distance (as of the last flush); -inf for padding / NaN queries: nothing is ever appended
    float thr = qok ? INFINITY : -INFINITY;
    float bound = 0.f;                                        // warp maximum of thr
    bool bound_stale = true;

    auto flush = [&]() {
        if (K1) return;
        const int cnt = (int)(qp - qbase) >> 7;
        TREE_STAT(3, 1);
        TREE_STAT(5, __reduce_add_sync(FULL, cnt - kept));
        {   // the (<= 16) new values -> sorted -> merged into the register list
            float nd[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) nd[j] = kept + j < cnt ? qd[(kept + j) * 32] : INFINITY;
            sort16v(nd);
            merge16v(sv, nd);
        }
        thr = qok ? sv[TREE_K - 1] : -INFINITY;
        // compaction: keep the pairs with d <= k-th distance
        int mx = __reduce_max_sync(FULL, cnt);
        int w = 0;
#pragma unroll 4
        for (int j = 0; j < mx; ++j) {       // branch-free: predicated loads / stores (slots beyond cnt are inside the buffer)
            const float d = qd[j * 32];
            const unsigned short id = qi[j * 32];
            const bool keep = j < cnt && d <= thr;
            if (keep) { qd[w * 32] = d; qi[w * 32] = id; }
            w += keep ? 1 : 0;
        }
        // more than k survivors: ties at the k-th distance -> drop the tied entries with the largest indices
        if (__any_sync(FULL, w > k)) {
            TREE_STAT(4, 1);
            while (w > k) {
                int worst = -1, wi = -1;
                for (int j = 0; j < w; ++j)
                    if (qd[j * 32] == thr && (int)qi[j * 32] > wi) { wi = qi[j * 32]; worst = j; }
                --w;
                qd[worst * 32] = qd[w * 32]; qi[worst * 32] = qi[w * 32];
            }
        }
        kept = w;
        qp = qbase + (uint32_t)w * 128u;
        ip = ibase + (uint32_t)w * 64u;
        bound_stale = true;
    };

    // ---- traversal state: lane = supernode (D1) / lane = leaf of the current supernode (D0) ----
    float D1 = INFINITY;
    bool sdone = true;
    if (lane < P.sup.nsuper) {
        const float *sb = P.sup.sbox + (size_t)b * 6 * 32 + lane;
        D1 = gap2_3(gap1(sb[0], sb[3 * 32], qlo[0], qhi[0]), gap1(sb[32], sb[4 * 32], qlo[1], qhi[1]),
                    gap1(sb[2 * 32], sb[5 * 32], qlo[2], qhi[2]));
        sdone = false;
    }
    float D0 = INFINITY;
    bool ldone = true;
    int cur_super = -1;
    auto load_super = [&](int s) {
        cur_super = s;
        const int lf = s * 32 + lane;
        ldone = lf >= nleaf;
        D0 = INFINITY;
        float bx[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) bx[a] = ldone ? (a < 3 ? INFINITY : -INFINITY) : sbx[a * nlpad + lf];
        if (!ldone) D0 = gap2_3(gap1(bx[0], bx[3], qlo[0], qhi[0]), gap1(bx[1], bx[4], qlo[1], qhi[1]), gap1(bx[2], bx[5], qlo[2], qhi[2]));
        __syncwarp();                                         // readers of the previous supernode's boxes are done
#pragma unroll
        for (int a = 0; a < 6; ++a) bxc[a * 32 + lane] = bx[a];
        __syncwarp();
    };
    // next leaf to visit (warp-uniform), nearest first, or -1.  `bound` = the largest k-th distance of the warp: a box
    // farther than that from the query box holds nothing for any lane.  Skipped boxes stay skipped (thr only shrinks).
    auto pick = [&]() -> int {
        if (bound_stale) {
            bound = thr;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(FULL, bound, o));
            bound_stale = false;
        }
        for (;;) {
            if (cur_super >= 0) {
                const bool c = !ldone && D0 <= bound;
                if (__ballot_sync(FULL, c)) {
                    const float m = warp_min(c ? D0 : INFINITY);
                    const int sl = __ffs(__ballot_sync(FULL, c && D0 == m)) - 1;
                    if (lane == sl) ldone = true;
                    return cur_super * 32 + sl;
                }
            }
            const bool c = !sdone && D1 <= bound;
            if (!__ballot_sync(FULL, c)) return -1;
            const float m = warp_min(c ? D1 : INFINITY);
            const int sl = __ffs(__ballot_sync(FULL, c && D1 == m)) - 1;
            if (lane == sl) sdone = true;
            load_super(sl);
        }
    };

    unsigned phase = 0;                                       // mbarrier parity per stage
    auto issue = [&](int leaf, int s) {
        if (lane == 0) {
            mbar_expect_tx(&bar[s], 512);
            bulk_g2s(stage + s * 128, SL + leaf, 512, &bar[s]);
        }
    };

    const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);

    int cur;
    if (P.self) {                                             // a self query starts with its own leaf
        load_super(ql >> 5);
        if (lane == (ql & 31)) ldone = true;
        if (lane == (ql >> 5)) sdone = true;
        cur = ql;
    } else {
        cur = pick();
    }
    int st = 0;
    if (cur >= 0) issue(cur, 0);
    while (cur >= 0) {
        // exact per-lane test of the current leaf against the lane's own query and k-th distance (its box is still in the
        // cache: pick() below may move on to another supernode)
        const int cl = cur & 31;
        const float lb = gap2_3(gap1(bxc[cl], bxc[96 + cl], qx, qx), gap1(bxc[32 + cl], bxc[128 + cl], qy, qy),
                                gap1(bxc[64 + cl], bxc[160 + cl], qz, qz));
        const int nxt = pick();
        __syncwarp();                                         // every lane is done with the stage that is refilled now
        if (nxt >= 0) issue(nxt, st ^ 1);
        TREE_STAT(0, 1);
        const bool need = qok && lb <= thr;
        mbar_wait(&bar[st], (phase >> st) & 1u);
        phase ^= 1u << st;
        // the last leaf is always walked through: the final flush sits in its loop
        if (__any_sync(FULL, need) || (!K1 && nxt < 0)) {
            TREE_STAT(1, 1);
            TREE_STAT(2, __popc(__ballot_sync(FULL, need)));
            const float *S = stage + st * 128;
#pragma unroll 1
            for (int h = 0; h < 32; h += TREE_CHK) {
#pragma unroll
                for (int c = 0; c < TREE_CHK; c += 4) {
                    const float *Sc = S + h + c;
                    const ulonglong2 X = *(const ulonglong2 *)(Sc), Y = *(const ulonglong2 *)(Sc + 32), Z = *(const ulonglong2 *)(Sc + 64);
                    const int4 I = *(const int4 *)(Sc + 96);
                    const f32x2 dx01 = add2(qx2, X.x), dx23 = add2(qx2, X.y);
                    const f32x2 dy01 = add2(qy2, Y.x), dy23 = add2(qy2, Y.y);
                    const f32x2 dz01 = add2(qz2, Z.x), dz23 = add2(qz2, Z.y);
                    f32x2 e01 = mul2(dx01, dx01), e23 = mul2(dx23, dx23);
                    e01 = fma2(dy01, dy01, e01); e23 = fma2(dy23, dy23, e23);
                    e01 = fma2(dz01, dz01, e01); e23 = fma2(dz23, dz23, e23);
                    float d[4];
                    unpack2(e01, d[0], d[1]);
                    unpack2(e23, d[2], d[3]);
                    const int id[4] = {I.x, I.y, I.z, I.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (K1) {
                            if (d[u] < bd || (d[u] == bd && id[u] < bi)) { bd = d[u]; bi = id[u]; }
                        } else {
                            if (d[u] <= thr) { sts_f32(qp, d[u]); sts_u16(ip, id[u]); qp += 128; ip += 64; }
                        }
                    }
                }
                // every TREE_CHK candidates: more than TREE_TRIG new entries somewhere?  (at most 16 are new then: one round of the merge
                // network, and kept + new <= 32 always fits the buffer; on the first leaf that is after 16 candidates, when
                // every lane holds exactly 16)
                if (!K1 && (__any_sync(FULL, qp > qbase + (uint32_t)(kept + TREE_TRIG) * 128u) || (nxt < 0 && h == 32 - TREE_CHK))) flush();
            }
            if (K1) { thr = qok ? bd : -INFINITY; bound_stale = true; }
        }
        cur = nxt;
        st ^= 1;
    }

    // ---- results ----
    if (!qvalid) return;
    int64_t *o = P.idx + (size_t)b * P.idx_bs + (size_t)qidx * k;
    int64_t *o2 = (P.idx2 != nullptr && qidx < P.idx2_rows) ? P.idx2 + (size_t)b * P.idx2_bs + (size_t)qidx * k : nullptr;
    float *od = P.dist2 != nullptr ? P.dist2 + (size_t)b * P.idx_bs + (size_t)qidx * k : nullptr;
    if (K1) {
        const int64_t v = bi == NOIDX ? (int64_t)-1 : (int64_t)bi;
        o[0] = v;
        if (o2) o2[0] = v;
        if (od) od[0] = bd;
    } else {
        // the <= k surviving pairs, sorted once in lexicographic (d, idx) order
        float fd[16];
        int fi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            fd[j] = j < kept ? qd[j * 32] : INFINITY;
            const int v16 = j < kept ? (int)qi[j * 32] : 0xffff;
            fi[j] = v16 == 0xffff ? NOIDX : v16;
        }
        sort16(fd, fi);
#pragma unroll
        for (int p = 0; p < TREE_K; ++p) {
            if (p < k) {
                const int64_t v = fi[p] == NOIDX ? (int64_t)-1 : (int64_t)fi[p];
                o[p] = v;
                if (o2) o2[p] = v;
                if (od) od[p] = fd[p];
            }
        }
    }
}

}  // namespace

size_t knn_tree_slot_bytes(int B, int n) {
    const int nleaf = (n + 31) / 32, nsuper = (nleaf + 31) / 32;
    return ws_block((size_t)B * nleaf * sizeof(KnnLeaf)) + ws_block((size_t)B * 6 * nsuper * 32 * sizeof(float)) +
           ws_block((size_t)B * 6 * 32 * sizeof(float));
}

bool knn_tree_take_slot(Workspace &W, int B, int n, KnnTreeView *v) {
    const int nleaf = (n + 31) / 32, nsuper = (nleaf + 31) / 32;
    v->n = n; v->nleaf = nleaf; v->nsuper = nsuper; v->nlpad = nsuper * 32;
    v->leaves = W.take<KnnLeaf>((size_t)B * nleaf);
    v->box = W.take<float>((size_t)B * 6 * nsuper * 32);
    v->sbox = W.take<float>((size_t)B * 6 * 32);
    return W.ok();
}

int launch_knn_tree_build(const float4 *pts4, long long pts_bs, const KnnTreeView *trees, int ntrees, int B, cudaStream_t st) {
    if (ntrees <= 0 || B <= 0) return DSIR_OK;
    if (ntrees > DSIR_MAX_LEVELS + 1) return DSIR_ERR_UNSUPPORTED;
    KnnTreeBuildParams P{};
    P.pts4 = pts4; P.pts_bs = pts_bs;
    int cap = 0;
    for (int g = 0; g < ntrees; ++g) {
        if (trees[g].n < 1 || trees[g].n > KNN_TREE_MAX_POINTS) return DSIR_ERR_UNSUPPORTED;
        P.n[g] = trees[g].n;
        P.leaves[g] = const_cast<KnnLeaf *>(trees[g].leaves);
        P.box[g] = const_cast<float *>(trees[g].box);
        P.sbox[g] = const_cast<float *>(trees[g].sbox);
        cap = trees[g].n > cap ? trees[g].n : cap;
    }
    cap = (cap + 1023) / 1024 * 1024;
    if (cap < 1024) cap = 1024;
    P.cap = cap;
    const size_t smem = (size_t)cap * 12 + 32 * 256 * 2;   // 3 x u16 coordinates + 2 x u16 permutation + u16 scratch per point, histograms
    DSIR_CUDA_TRY(cudaFuncSetAttribute(knn_tree_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ntrees, B);
    knn_tree_build_kernel<<<grid, TREE_BUILD_THREADS, smem, st>>>(P);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

#ifdef DSIR_KNN_STATS
extern "C" int dsir_knn_tree_stats(unsigned long long *out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_tree_stats, sizeof(g_tree_stats));
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(g_tree_stats, z, sizeof(z)); }
    return 0;
}
#endif

int launch_knn_tree_query(const KnnTreeQueryParams &P, int B, cudaStream_t st) {
    if (P.k < 1 || P.k > KNN_TREE_MAX_K) return DSIR_ERR_UNSUPPORTED;
    if (P.sup.n < P.k) return DSIR_ERR_KNN_TOO_FEW;
    if (P.qry.nleaf <= 0 || B <= 0) return DSIR_OK;
    if (P.sup.nsuper > 32) return DSIR_ERR_UNSUPPORTED;
    dim3 grid((P.qry.nleaf + TREE_QWARPS - 1) / TREE_QWARPS, B);
    if (P.k == 1) {
        const size_t smem = (size_t)TREE_QWARPS * tree_warp_smem(true);
        knn_tree_query_kernel<true><<<grid, TREE_QWARPS * 32, smem, st>>>(P);
    } else {
        const size_t smem = (size_t)TREE_QWARPS * tree_warp_smem(false);
        DSIR_CUDA_TRY(cudaFuncSetAttribute(knn_tree_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_tree_query_kernel<false><<<grid, TREE_QWARPS * 32, smem, st>>>(P);
    }
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace dsir
