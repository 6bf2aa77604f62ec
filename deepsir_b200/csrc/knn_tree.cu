// Exact xyz k-nearest neighbours on a bucket tree of Morton-ordered leaves — the warp-cooperative fast path for level
// clouds of 512 ... 24576 points (dataloader/data_base.py:165,170 call sites; contract of knn.cu / oracle/knn_oracle.c):
//     d2 = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32, results ascending in (d2, support index).
//
//   build   one CTA per (cloud, tree): bounding box -> 30-bit Morton keys -> stable LSD radix sort of the permutation in
//           shared memory (8-bit digits, per-warp histograms ranked with match.any) -> LEAVES of 32 consecutive points,
//           stored as one 512-byte structure-of-arrays block {-x[32], -y[32], -z[32], index[32]} -> bounding box per leaf
//           and per SUPERNODE (32 consecutive leaves).  A cloud of 16384 points is 512 leaves / 16 supernodes.
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
constexpr int TREE_QWARPS = 4;     // query leaves (warps) per CTA
constexpr int TREE_NST = 2;        // staging ring depth per warp
constexpr int TREE_QCAP = 16;      // pending-candidate queue entries per lane
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
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ unsigned part1by2(unsigned v) {   // spread the low 10 bits to every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

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
// build
// ---------------------------------------------------------------------------------------------------------------
__global__ __launch_bounds__(TREE_BUILD_THREADS) void knn_tree_build_kernel(KnnTreeBuildParams P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ float s_red[6][32];
    __shared__ float s_box[8];
    __shared__ int s_wtot[32];

    const int g = blockIdx.x, b = blockIdx.y;
    const int n = P.n[g];
    const int nleaf = (n + 31) >> 5, nsuper = (nleaf + 31) >> 5, nlpad = nsuper * 32;
    const float4 *pts = P.pts4 + (size_t)b * P.pts_bs;
    unsigned *key = (unsigned *)sm_raw;                            // [cap]
    unsigned short *pa = (unsigned short *)(key + P.cap);          // [cap] permutation (ping)
    unsigned short *pb = pa + P.cap;                               // [cap] permutation (pong)
    unsigned short *hist = pb + P.cap;                             // [32 warps][256 digits]
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
        float L[3], ext = 0.f;
        for (int a = 0; a < 3; ++a) {
            float l = INFINITY, h = -INFINITY;
            for (int w = 0; w < TREE_BUILD_THREADS / 32; ++w) { l = fminf(l, s_red[a][w]); h = fmaxf(h, s_red[3 + a][w]); }
            if (!(l <= h) || !isfinite(l) || !isfinite(h)) { l = 0.f; h = 0.f; }
            L[a] = l;
            ext = fmaxf(ext, h - l);
        }
        s_box[0] = L[0]; s_box[1] = L[1]; s_box[2] = L[2];
        s_box[3] = (ext > 0.f && isfinite(ext)) ? 1023.0f / ext : 0.f;   // cubic cells: one scale for the three axes
    }
    __syncthreads();
    const float ox = s_box[0], oy = s_box[1], oz = s_box[2], scale = s_box[3];

    // ---- Morton keys (the quantisation only shapes the leaves; the search is exact for any order) ----
    for (int i = tid; i < n; i += TREE_BUILD_THREADS) {
        const float4 p = pts[i];
        const int ix = min(max(__float2int_rz((p.x - ox) * scale), 0), 1023);
        const int iy = min(max(__float2int_rz((p.y - oy) * scale), 0), 1023);
        const int iz = min(max(__float2int_rz((p.z - oz) * scale), 0), 1023);
        key[i] = part1by2((unsigned)ix) | (part1by2((unsigned)iy) << 1) | (part1by2((unsigned)iz) << 2);
    }

    // ---- stable LSD radix sort of the permutation, 4 passes of 8 bits.  Warp w owns a contiguous range of 32-blocks;
    //      inside a block the rank of an element among equal digits comes from match.any ----
    const int bpw = (nleaf + 31) >> 5;                       // 32-blocks per warp
    const int blk0 = min(warp * bpw, nleaf), blk1 = min(blk0 + bpw, nleaf);
    unsigned short *pin = pa, *pout = pb;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        for (int i = tid; i < 32 * 256; i += TREE_BUILD_THREADS) hist[i] = 0;
        __syncthreads();                                     // also orders key[] / pout writes of the previous step
        unsigned short *wh = hist + warp * 256;
        for (int blk = blk0; blk < blk1; ++blk) {
            const int pos = blk * 32 + lane;
            const bool v = pos < n;
            const unsigned e = v ? (pass == 0 ? (unsigned)pos : (unsigned)pin[pos]) : 0u;
            const unsigned dg = v ? ((key[e] >> shift) & 255u) : 0xffffu;
            const unsigned m = __match_any_sync(FULL, dg);
            if (v && lane == __ffs(m) - 1) wh[dg] = (unsigned short)(wh[dg] + __popc(m));
            __syncwarp();
        }
        __syncthreads();
        // exclusive scan in (digit-major, warp-minor) order: entry e = digit * 32 + warp; thread t owns entries 8t .. 8t+7
        {
            const int dg = tid >> 2, w0 = (tid & 3) * 8;
            int loc[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { loc[i] = hist[(w0 + i) * 256 + dg]; sum += loc[i]; }
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
            for (int i = 0; i < 8; ++i) { hist[(w0 + i) * 256 + dg] = (unsigned short)run; run += loc[i]; }
        }
        __syncthreads();
        for (int blk = blk0; blk < blk1; ++blk) {
            const int pos = blk * 32 + lane;
            const bool v = pos < n;
            const unsigned e = v ? (pass == 0 ? (unsigned)pos : (unsigned)pin[pos]) : 0u;
            const unsigned dg = v ? ((key[e] >> shift) & 255u) : 0xffffu;
            const unsigned m = __match_any_sync(FULL, dg);
            const int rank = __popc(m & ((1u << lane) - 1u));
            const int base = v ? (int)wh[dg] : 0;
            __syncwarp();
            if (v && lane == __ffs(m) - 1) wh[dg] = (unsigned short)(base + __popc(m));
            __syncwarp();
            if (v) pout[base + rank] = (unsigned short)e;
        }
        __syncthreads();
        unsigned short *t = pin; pin = pout; pout = t;
    }

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
// Selection: every lane keeps its k best (d, idx) pairs as a SORTED LIST IN REGISTERS (16 pairs).  Candidates that pass
// the lane's threshold are appended to a per-lane queue in shared memory (column layout [slot][lane]: conflict free);
// when some queue runs full, ALL lanes merge their pending entries at once with sorting networks (Batcher odd-even merge
// sort of the 16 newest + bitonic merge with the list): branch-free, the same instruction stream for every lane, so a
// flush costs the same whether one lane or all lanes have work - the opposite of per-candidate heap insertions, whose
// data-dependent loops ran at ~35 % lane efficiency and took 2/3 of the instructions of the first version of this kernel.
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
// list (ascending) <- the 16 smallest of list U nw (both ascending), ascending
__device__ __forceinline__ void merge16(float (&ld)[16], int (&li)[16], const float (&nd)[16], const int (&ni)[16]) {
#pragma unroll
    for (int a = 0; a < 16; ++a) {   // lower half of the bitonic sequence (list, reversed nw)
        const float bd = nd[15 - a];
        const int bi = ni[15 - a];
        const bool sw = ld[a] > bd || (ld[a] == bd && li[a] > bi);
        ld[a] = sw ? bd : ld[a];
        li[a] = sw ? bi : li[a];
    }
#pragma unroll
    for (int k = 8; k >= 1; k >>= 1)
#pragma unroll
        for (int a = 0; a < 16; ++a)
            if ((a & k) == 0) ce(ld[a], li[a], ld[a + k], li[a + k]);
}

__host__ __device__ constexpr int tree_warp_smem(bool k1) {
    // stages + mbarriers (padded to 128) + leaf-box cache [6][32] + queue
    return TREE_NST * 512 + 128 + 768 + (k1 ? 0 : TREE_QCAP * 32 * 8);
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
    int *qi = (int *)(qd + TREE_QCAP * 32);

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
    // the k best so far, ascending, in slots [16 - k, 16): the first 16 - k slots hold (-inf, -1) sentinels that sort
    // before every candidate, so that the k-th best is always ld[15] (no dynamic register index)
    float ld[TREE_K];
    int li[TREE_K];
#pragma unroll
    for (int p = 0; p < TREE_K; ++p) { ld[p] = p < TREE_K - k ? -INFINITY : INFINITY; li[p] = p < TREE_K - k ? -1 : NOIDX; }
    float bd = INFINITY;                                      // K1: the best pair
    int bi = NOIDX;
    const uint32_t qbase = smem_u32(qd);
    uint32_t qp = qbase;                                      // queue write cursor of this lane (entry j at qbase + 128 j)
    // the lane's current k-th distance (as of the last flush); -inf for padding / NaN queries: nothing is ever appended
    float thr = qok ? INFINITY : -INFINITY;
    float bound = 0.f;                                        // warp maximum of thr
    bool bound_stale = true;

    auto flush = [&]() {
        if (K1) return;
        const int cnt = (int)(qp - qbase) >> 7;
        int mx = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
        TREE_STAT(3, 1);
        TREE_STAT(4, (mx + 15) >> 4);
        TREE_STAT(5, __reduce_add_sync(FULL, cnt));
        for (int base = 0; base < mx; base += 16) {           // one round unless some queue holds more than 16
            float nd[16];
            int ni[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const bool v = base + j < cnt;
                nd[j] = v ? qd[(base + j) * 32] : INFINITY;
                ni[j] = v ? qi[(base + j) * 32] : NOIDX;
            }
            sort16(nd, ni);
            merge16(ld, li, nd, ni);
        }
        qp = qbase;
        thr = qok ? ld[TREE_K - 1] : -INFINITY;
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
            for (int h = 0; h < 32; h += 8) {
#pragma unroll
                for (int c = 0; c < 8; c += 4) {
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
                            if (d[u] <= thr) { sts_f32(qp, d[u]); sts_s32(qp + TREE_QCAP * 128, id[u]); qp += 128; }
                        }
                    }
                }
                // every 8 candidates: room for 8 more in every queue?  (at most 16 entries are pending then: one round of
                // the merge network; on the first leaf that is after 16 candidates, when every lane holds exactly 16)
                if (!K1 && (__any_sync(FULL, qp > qbase + (TREE_QCAP - 8) * 128) || (nxt < 0 && h == 24))) flush();
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
#pragma unroll
        for (int s = 0; s < TREE_K; ++s) {
            const int p = s - (TREE_K - k);
            if (p >= 0) {
                const int64_t v = li[s] == NOIDX ? (int64_t)-1 : (int64_t)li[s];
                o[p] = v;
                if (o2) o2[p] = v;
                if (od) od[p] = ld[s];
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
    const size_t smem = (size_t)cap * 8 + 32 * 256 * 2;
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
