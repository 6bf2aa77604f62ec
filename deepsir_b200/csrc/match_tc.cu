// Fused feature-distance + row-argmin on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces the chunked block network/model.py:558-569 of the reference (match_features_V2 + .min(dim=2)[1]);
// the [J,K] score matrix never leaves TMEM/registers.  Bit-exactness on the indices is kept by a
// FILTER-AND-REFINE scheme:
//
//   prep    point-major (K-major) fp32 copies of both feature sets (reference side scaled by -2, exact) for the
//           refine stage, and fp16 copies for the tensor cores: every value is multiplied by a per-batch power of
//           two sigma (exact) that brings the largest feature norm below 1, then rounded to fp16.  The squared
//           reference norm is FOLDED INTO THE CONTRACTION: 16 extra channels carry 1,1,1,0.. on the source side and
//           the three-term fp16 split (hi, lo, lolo) of sigma^2 |r_k|^2 on the reference side, so the accumulator
//           already holds  x_jk = sigma^2 (|r_k|^2 - 2 <s_j, r_k>)  and the epilogue needs no add and no shared
//           memory.
//   filter  persistent warp-specialised kernel, one CTA per SM, 640 threads:
//             warp 0      TMA producer  (SWIZZLE_128B feature boxes + SWIZZLE_32B norm boxes -> mbarrier ring)
//             warp 1      one thread issues tcgen05.mma kind::f16 (M=128, N=128, K=16; 5 per tile) and
//                         tcgen05.commit; four M=128 accumulators (512 source rows per work item) fill the 512
//                         TMEM columns, so the accumulator of row block i is drained while the tensor pipe works
//                         on the other three
//             warp 2      TMEM allocation
//             warps 4-19  epilogue: tcgen05.ld 32 columns at a time, a 3-input-min tree, ONE warp vote per 32
//                         columns; only when some row sees a value within `margin` of its running minimum does the
//                         slow path re-read the flagged columns and insert them into the row's sorted candidate
//                         list.
//           margin_j = 2 eps_j, where eps_j bounds the fp16-input error of x_jk (see tc_margin), so the exact fp32
//           argmin is always among the kept candidates unless the list saturated.
//   refine  per row: candidates within margin of the approximate minimum are re-scored in exact fp32 with the
//           op order of match_fp32.cu (fma chain over channels, ((-2 dot)+ns)+nr) and the (value, index)
//           lexicographic minimum is taken -> identical indices AND minima to the fp32 kernel.
//   rescue  rows whose list saturated (or held no finite candidate) are recomputed exhaustively in fp32.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "match_tc.cuh"
#include <vector>
#include "tc_common.cuh"

namespace dsir {

namespace {

constexpr int TC_T = 4;                       // candidates kept per row per (split, column part)
// Geometry of one CTA.  The 512 TMEM columns hold RBS x ACC_STAGES accumulators of 128 columns; every accumulator is
// an independent  issue -> drain -> issue  chain, and the chains only meet in the tensor pipe.  (4 row blocks x 1 stage
// halves the L2->smem traffic of the reference tiles compared with 2 x 2 and measured faster.)
constexpr int TC_RBS = 4;                     // 128-row blocks per work item
constexpr int TC_ACC_STAGES = 1;              // accumulator stages per row block (RBS * ACC_STAGES == 4)
constexpr int TC_HALVES = 1;                  // column parts of an accumulator drained by different warps (1 or 2)
constexpr int TC_MMA_WARPS = 2;               // MMA issuer warps; issuer w owns row blocks w, w + MMA_WARPS, ..
constexpr int TC_BM = 128 * TC_RBS;           // source rows per work item
constexpr int TC_BN = 128;                    // reference rows per unit (one N=128 MMA)
constexpr int TC_STAGES = 5;                  // B ring depth
constexpr int TC_EPI_WARPS = 4 * TC_RBS * TC_HALVES;   // warp = 4 * (r * HALVES + h) + q: row block r, column part h, lane quadrant q
constexpr int TC_STEPS = TC_BN / TC_HALVES / 32;       // 32-column steps per warp per unit
static_assert(TC_RBS * TC_ACC_STAGES == 4 && TC_EPI_WARPS == 16 && TC_RBS % TC_MMA_WARPS == 0, "TMEM / warp budget");
// The issue arbiter of an SM sub-partition prefers the HIGHEST warp id, so the control warps sit above the 16 epilogue
// warps.  Warp 16 allocates TMEM and then produces (TMA); warps 17.. issue MMAs (whole warp, one elected lane).
// clock64() stamps in the MMA / epilogue loops and the DSIR_TC_DEBUG experiment switches exist only in builds made with
// -DDSIR_TC_TRACE (tools/build_variant.sh); the shipped library carries neither.
#ifdef DSIR_TC_TRACE
constexpr bool TC_TRACE = true;
#else
constexpr bool TC_TRACE = false;
#endif
constexpr int TC_WARP_TMA = TC_EPI_WARPS, TC_WARP_MMA0 = TC_EPI_WARPS + 1;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_MMA_WARPS) * 32;
constexpr int TC_LISTS = TC_HALVES;           // candidate lists per (row, split)
constexpr int TC_MAX_SPLIT = 8 / TC_HALVES;
constexpr double TC_ITEM_OVERHEAD = 24.0;     // fixed cost of a work item in units of one 512x128 unit (fitted: tools/sweep_split.py)
constexpr int TC_CH = 64;                     // fp16 channels per point in the tensor-core copy (one 128-byte swizzle row)
constexpr int TC_AUG = 16;                    // folded-norm channels (one K=16 MMA)
constexpr float TC_PAD_NORM = 60000.0f;       // folded norm of padded reference rows: larger than any real x_jk (<= 3)
constexpr uint32_t MAIN_TILE = 128 * TC_CH * 2;   // 16 KB: 128 rows x 64 halves, 128-byte swizzle
constexpr uint32_t AUG_TILE = 128 * TC_AUG * 2;   //  4 KB: 128 rows x 16 halves, 32-byte swizzle

__device__ __forceinline__ float fmin3(float a, float b, float c) {   // FMNMX3 (sm_100+)
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// instruction descriptor: D=f32, A=B=f16, both K-major, N=128, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// Bound on |x_hat_jk - x_jk| in scaled units, x = sigma^2 (|r|^2 - 2<s,r>), for fp16-rounded operands with fp32
// accumulation; a = sigma |s_j|, R = sigma max_k |r_k| (both <= 1):
//   2^-9 (1+2^-12) a R          rounding of s' and -2r' to fp16 (unit round-off 2^-11 each, Cauchy-Schwarz)
//   2^-25 sqrt(C) (a + 2R)      fp16 subnormal spacing (values below 2^-14)
//   80 * 2^-22 (2aR + R^2)      fp32 accumulation of the 80 products inside the tensor core (4x an IEEE chain)
//   2^-24 + 2^-30 R^2           three-term fp16 split of the folded norm
// margin = 2 eps (+2 %): every column whose exact value could be the row minimum lies within margin of the
// approximate minimum.
__device__ __forceinline__ float tc_margin(float ns_j, float rmax_b, float sigma, int C) {
    const float a = sigma * sqrtf(ns_j);
    const float R = sigma * sqrtf(rmax_b);
    const float eps = 1.9536e-3f * a * R + 2.9803e-8f * sqrtf((float)C) * (a + 2.f * R) + 1.9074e-5f * (2.f * a * R + R * R) +
                      5.97e-8f + 9.4e-10f * R * R;
    return 2.04f * eps + 1e-30f;
}

template <int T>
__device__ __forceinline__ void cand_insert(float (&cv)[T], int (&ci)[T], float x, int col) {
#pragma unroll
    for (int p = T - 1; p >= 0; --p) {
        bool shift = (p > 0) && (x < cv[p > 0 ? p - 1 : 0]);
        bool here = !shift && (x < cv[p]);
        float nv = shift ? cv[p > 0 ? p - 1 : 0] : (here ? x : cv[p]);
        int ni = shift ? ci[p > 0 ? p - 1 : 0] : (here ? col : ci[p]);
        cv[p] = nv;
        ci[p] = ni;
    }
}

// One observed value enters a row's candidate list (sorted ascending, T entries).  The common event is a CLEAR
// record low (x below the current minimum by more than the margin): every older entry is then outside the margin
// of the new minimum and the list collapses to {x}.  Anything else (a near tie) takes the sorted insert.
__device__ __forceinline__ void cand_update(float (&cv)[TC_T], int (&ci)[TC_T], float x, int col, float margin, float &thr) {
    if (x < cv[0] - margin) {
        cv[0] = x; ci[0] = col;
#pragma unroll
        for (int t = 1; t < TC_T; ++t) cv[t] = INFINITY;
    } else {
        cand_insert<TC_T>(cv, ci, x, col);
    }
    thr = fminf(thr, cv[0] + margin);   // never above the primed bound
}

__device__ __forceinline__ void tmem_ld4_sync(uint32_t taddr, float (&x)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    x[0] = __uint_as_float(r0); x[1] = __uint_as_float(r1); x[2] = __uint_as_float(r2); x[3] = __uint_as_float(r3);
}

// 32 accumulator columns of one row.  Fast path: minima of the 8 aligned column quads, a 3-input-min tree over them
// (20 FMNMX/FMNMX3) and ONE vote.  Whenever some lane of the warp sees a value within `margin` of its running minimum,
// the slow path re-reads only the flagged QUADS from TMEM (four values in registers with static indices) and updates
// the candidate list of the lanes concerned; the hot loop carries no per-element branches.
// LAST: this is the unit's final step, every tcgen05.ld of the accumulator has completed -> the accumulator goes back to
// the MMA issuer as soon as the vote says that nothing has to be re-read from TMEM (or right after the re-reads).
__device__ __forceinline__ void release_acc(uint64_t *empty_bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar);
}
// MODE (template of the kernel): 0 = row argmin (candidate lists), 1 = value only (minima per column granule, top-k
// sweep 1), 2 = collect (top-k sweep 2: every column at or below the row's FIXED threshold is appended to the row's list)
struct TopkRow {
    int n;        // columns seen at or below the threshold (the row has ONE owner thread: top-k plans never split K)
    int *list;    // the row's column list, `cap` entries
    int cap;
};
template <bool LAST, int MODE = 0>
__device__ __forceinline__ void filter32(const uint32_t (&v)[32], int col0, uint32_t taddr, float margin, float &thr,
                                         float (&cv)[TC_T], int (&ci)[TC_T], bool sample, float &best,
                                         uint64_t *empty_bar = nullptr, int lane = 0, TopkRow *tk = nullptr) {
#define F(i) __uint_as_float(v[i])
    const float a0 = fmin3(F(0), F(1), F(2)), a1 = fmin3(F(3), F(4), F(5)), a2 = fmin3(F(6), F(7), F(8));
    const float a3 = fmin3(F(9), F(10), F(11)), a4 = fmin3(F(12), F(13), F(14)), a5 = fmin3(F(15), F(16), F(17));
    const float a6 = fmin3(F(18), F(19), F(20)), a7 = fmin3(F(21), F(22), F(23)), a8 = fmin3(F(24), F(25), F(26));
    const float a9 = fmin3(F(27), F(28), F(29));
    const float b0 = fmin3(a0, a1, a2), b1 = fmin3(a3, a4, a5), b2 = fmin3(a6, a7, a8), b3 = fmin3(a9, F(30), F(31));
    const float mm = fminf(fmin3(b0, b1, b2), b3);
    if (sample) {              // priming pass: only the value of the running minimum, no candidates, no branches
        if (LAST) release_acc(empty_bar, lane);
        best = fminf(best, mm);
        return;
    }
    const bool slow = __any_sync(0xffffffffu, mm <= thr);   // inclusive, like the refine step's `v <= gmin + margin`
    if (LAST && !slow) release_acc(empty_bar, lane);
    if constexpr (MODE == 2) {
        // collect: tens of hits per row and sweep - no TMEM re-read, the 32 values are tested where they are (static indices)
        if (slow) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
                if (F(e) <= thr) {
                    if (tk->n < tk->cap) tk->list[tk->n] = col0 + e;
                    ++tk->n;
                }
            if (tk->n > tk->cap) thr = -INFINITY;   // the list is full: the row goes to the exhaustive pass, stop looking
            if (LAST) release_acc(empty_bar, lane);
        }
        return;
    }
    if (slow) {
        unsigned qm = 0;       // which aligned column quads hold a value below the threshold
#pragma unroll
        for (int g = 0; g < 8; ++g)
            qm |= (fminf(fmin3(F(4 * g), F(4 * g + 1), F(4 * g + 2)), F(4 * g + 3)) <= thr) ? (1u << g) : 0u;
#undef F
        unsigned um = __reduce_or_sync(0xffffffffu, qm);
#pragma unroll 1
        while (um) {
            const int g = __ffs(um) - 1;
            um &= um - 1;
            float x[4];
            tmem_ld4_sync(taddr + 4 * g, x);   // bit-identical to v[4g .. 4g+3]
            if ((qm >> g) & 1u) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (x[e] <= thr) cand_update(cv, ci, x[e], col0 + 4 * g + e, margin, thr);
            }
        }
        if (LAST) release_acc(empty_bar, lane);
    }
}

// Unit schedule of one work item: `ns` PRIMING units first (every (nu/ns)-th unit of the item's range, value-only epilogue:
// they give every row a tight upper bound of its minimum, so that the main pass meets ~ln(nu/ns) record lows per row
// instead of ~ln(32 nu)), then all `nu` units of the range.  The three roles walk the same schedule.
struct UnitSched {
    int u0, nu, ns, stride;
    __device__ __forceinline__ UnitSched(int sp, int S, int U, int prime_div) {
        u0 = (int)((long long)sp * U / S);
        nu = (int)((long long)(sp + 1) * U / S) - u0;
        ns = (prime_div > 0 && nu >= 2 * prime_div) ? nu / prime_div : 0;
        stride = ns > 0 ? nu / ns : 1;
    }
    __device__ __forceinline__ int total() const { return ns + nu; }
    __device__ __forceinline__ int unit(int t) const { return t < ns ? u0 + t * stride : u0 + (t - ns); }
};

struct TcParams {
    int B, J, K, C;
    int RB, U, S;           // row blocks (512), units (128), k-splits
    int Jpad, Kpad;
    const float *ns;        // [B,J] exact squared norms
    const float *rmax;      // [B] max reference squared norm
    const float *scale;     // [B] sigma (power of two)
    const float *xm;        // [B] >= 0: the folded-norm K-step is SKIPPED for this batch element and every margin is widened by
                            //     this amount (sigma^2 x spread of the reference norms); < 0: norm folded in (see TC_NOAUG_SPREAD)
    float *cand_val;        // [B][Jpad][S][T]  (scaled units)
    int *cand_idx;
    int prime_div;            // priming pass over every prime_div-th unit (0 = none)
    // optional prior correspondences [B,J] (the previous iteration of the alignment loop): the distance to the prior match
    // bounds the row minimum, so no priming pass is needed
    const int64_t *prior;
    const __half *a16, *b16;  // tensor-core copies [B][J][64], [B][K][64]
    const float *nr;          // [B,K] exact squared norms
    // top-k sweeps (MODE 1 / 2 of the kernel, see launch_match_tc_topk)
    float *umin;              // MODE 1 out: [B][G][Jpad] minima of x per column granule (G granules of `gran` columns)
    int gran32;               // MODE 1: granule = 32 columns (one epilogue step) instead of the 128-column unit
    int G;                    // granules per row
    const float *thr_in;      // MODE 2 in: [B][Jpad] fixed threshold per row
    int *tk_cnt;              // MODE 2: [B][Jpad] entries appended
    int *tk_list;             // MODE 2: [B][Jpad][tk_cap] column indices
    int tk_cap;
    int dbg_flags;            // experiments only (DSIR_TC_DEBUG): bit 0 = never take the slow path (wrong results)
    unsigned int *trace;      // DSIR_TC_DEBUG bit 1: block 0 logs clock stamps of its first 256 units (see match_tc_filter_trace)
    unsigned long long *dbg;  // [grid][4]: start ns, end ns, cycles, units (diagnostic, always written)
};

template <int NKS, int MODE = 0>  // 16-channel k-steps of the feature part (NKS = ceil(C / 16), C <= 64)
__global__ __launch_bounds__(TC_THREADS, 1) void match_tc_filter_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                        const __grid_constant__ CUtensorMap mapB,
                                                                        const __grid_constant__ CUtensorMap mapAaug,
                                                                        const __grid_constant__ CUtensorMap mapBaug,
                                                                        TcParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                            // [RBS][16 KB]
    uint8_t *sAaug = sA + TC_RBS * MAIN_TILE;                      // [4 KB]
    uint8_t *sB = sAaug + AUG_TILE;                                // [STAGES][16 KB]
    uint8_t *sBaug = sB + TC_STAGES * MAIN_TILE;                   // [STAGES][4 KB]
    uint64_t *bars = (uint64_t *)(sBaug + TC_STAGES * AUG_TILE);
    uint64_t *full_b = bars, *empty_b = bars + TC_STAGES;
    uint64_t *tmem_full = bars + 2 * TC_STAGES;                       // [ACC_STAGES][RBS]
    uint64_t *tmem_empty = tmem_full + TC_ACC_STAGES * TC_RBS;       // [ACC_STAGES][RBS]
    uint64_t *full_a = tmem_empty + TC_ACC_STAGES * TC_RBS, *empty_a = full_a + 1;
    uint32_t *tmem_slot = (uint32_t *)(empty_a + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform by construction
    const int total_items = P.B * P.RB * P.S;
    unsigned long long *dbg_s = (unsigned long long *)(tmem_slot + 2);   // start time / start clock, parked in smem
    if (threadIdx.x == 0) { dbg_s[0] = globaltimer_ns(); dbg_s[1] = clock64(); }

    if (warp == TC_WARP_TMA && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        prefetch_tmap(&mapAaug);
        prefetch_tmap(&mapBaug);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], TC_MMA_WARPS); }
        for (int a = 0; a < TC_ACC_STAGES * TC_RBS; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4 * TC_HALVES); }
        mbar_init(full_a, 1);
        mbar_init(empty_a, TC_MMA_WARPS);
        mbar_fence_init();
    }
    if (warp == TC_WARP_TMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC_WARP_TMA) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            PipeState pb{0, 0};
            uint32_t iphase = 0;
            bool first = true;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
                const int sp = it % P.S, rb = (it / P.S) % P.RB, b = it / (P.S * P.RB);
                const UnitSched us(sp, P.S, P.U, P.prime_div);
                mbar_wait(empty_a, iphase ^ 1u);
                mbar_expect_tx(full_a, TC_RBS * MAIN_TILE + (first ? AUG_TILE : 0u));
#pragma unroll
                for (int r = 0; r < TC_RBS; ++r)
                    tma_load_3d(sA + r * MAIN_TILE, &mapA, 0, rb * TC_BM + r * 128, b, full_a);
                if (first) tma_load_3d(sAaug, &mapAaug, 0, 0, 0, full_a);   // constant 1,1,1,0.. tile, loaded once
                first = false;
                const bool aug = P.xm[b] < 0.f;
                for (int t = 0; t < us.total(); ++t) {
                    const int u = us.unit(t);
                    while (!mbar_try_wait(&empty_b[pb.stage], pb.phase ^ 1u)) __nanosleep(32);
                    mbar_expect_tx(&full_b[pb.stage], MAIN_TILE + (aug ? AUG_TILE : 0u));
                    tma_load_3d(sB + pb.stage * MAIN_TILE, &mapB, 0, u * TC_BN, b, &full_b[pb.stage]);
                    if (aug) tma_load_3d(sBaug + pb.stage * AUG_TILE, &mapBaug, 0, u * TC_BN, b, &full_b[pb.stage]);
                    pb.advance(TC_STAGES);
                }
                iphase ^= 1u;
            }
        }
    } else if (warp >= TC_WARP_MMA0 && warp < TC_WARP_MMA0 + TC_MMA_WARPS) {
        // =========================== MMA issuer (whole warp, one elected lane issues) ===========================
        {
            const int w = warp - TC_WARP_MMA0;
            PipeState pb{0, 0}, pa{0, 0};
            uint32_t iphase = 0;
            const bool tracing = TC_TRACE && (P.dbg_flags & 2) && blockIdx.x == 0 && lane == 0;
            int useq = 0;
            const uint64_t descA0 = make_kmajor_desc(smem_u32(sA), 1024, 2);
            const uint64_t descAaug = make_kmajor_desc(smem_u32(sAaug), 256, 6);
            const uint64_t descB0 = make_kmajor_desc(smem_u32(sB), 1024, 2);
            const uint64_t descBaug0 = make_kmajor_desc(smem_u32(sBaug), 256, 6);
            for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
                const UnitSched us(it % P.S, P.S, P.U, P.prime_div);
                const bool aug = P.xm[it / (P.S * P.RB)] < 0.f;
                mbar_wait(full_a, iphase);
                for (int t = 0; t < us.total(); ++t) {
                    mbar_wait(&full_b[pb.stage], pb.phase);
                    const uint64_t descB = descB0 + (uint64_t)((uint32_t)pb.stage * (MAIN_TILE >> 4));
                    const uint64_t descBaug = descBaug0 + (uint64_t)((uint32_t)pb.stage * (AUG_TILE >> 4));
#pragma unroll
                    for (int r = w; r < TC_RBS; r += TC_MMA_WARPS) {
                        mbar_wait(&tmem_empty[pa.stage * TC_RBS + r], pa.phase ^ 1u);
                        tc_fence_after();
                        if (tracing && useq < 256 && r < 2) P.trace[(useq * 2 + r) * 2] = (unsigned int)clock64();
                        const uint32_t d_tmem = tmem_base + (uint32_t)((pa.stage * TC_RBS + r) * 128);
                        const uint64_t descA = descA0 + (uint64_t)(r * (MAIN_TILE >> 4));
#pragma unroll
                        for (int ks = 0; ks < NKS; ++ks)     // +32 bytes (16 halves) inside the 128-byte swizzle row
                            mma_f16(d_tmem, descA + (uint64_t)(ks * 2), descB + (uint64_t)(ks * 2), TC_IDESC, ks > 0 ? 1u : 0u);
                        if (aug) mma_f16(d_tmem, descAaug, descBaug, TC_IDESC, 1u);   // + sigma^2 |r_k|^2
                        tc_commit(&tmem_full[pa.stage * TC_RBS + r]);   // accumulator ready for its epilogue warps
                        if (tracing && useq < 256 && r < 2) P.trace[(useq * 2 + r) * 2 + 1] = (unsigned int)clock64();
                    }
                    tc_commit(&empty_b[pb.stage]);                   // one of the MMA_WARPS releases of this B stage
                    ++useq;
                    pb.advance(TC_STAGES);
                    pa.advance(TC_ACC_STAGES);
                }
                tc_commit(empty_a);
                iphase ^= 1u;
            }
        }
    } else if (warp < TC_EPI_WARPS) {
        // =========================== epilogue: TMEM -> registers -> candidate lists ===========================
        const int q = warp & 3;                       // TMEM lane quadrant of this warp
        const int h = (warp >> 2) % TC_HALVES;        // column part of the accumulator
        const int r = (warp >> 2) / TC_HALVES;        // row block
        const int trow = q * 32 + lane;               // row inside the 128-row block
        PipeState pa{0, 0};
        const bool tracing = TC_TRACE && (P.dbg_flags & 2) && blockIdx.x == 0 && warp == 0 && lane == 0;
        int useq = 0;
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const int sp = it % P.S, rb = (it / P.S) % P.RB, b = it / (P.S * P.RB);
            const UnitSched us(sp, P.S, P.U, P.prime_div);
            float best = INFINITY;
            float cv[TC_T];
            int ci[TC_T];
#pragma unroll
            for (int t = 0; t < TC_T; ++t) { cv[t] = INFINITY; ci[t] = -1; }
            const int j = rb * TC_BM + r * 128 + trow;
            // Rows beyond J (zero-filled by the TMA) would tie on EVERY column (x = sigma^2 |r_k|^2) and drag their warp
            // through the slow path at every step; their lists are never read, so they never look at anything.
            const bool dead = j >= P.J || (TC_TRACE && (P.dbg_flags & 1));
            float thr = dead ? -INFINITY : INFINITY;
            TopkRow tk{0, nullptr, 0};
            if constexpr (MODE == 2) {
                const size_t row = (size_t)b * P.Jpad + j;
                if (!dead) thr = P.thr_in[row];
                tk.list = P.tk_list + row * P.tk_cap; tk.cap = P.tk_cap;
            }
            float *umin_row = nullptr;
            if constexpr (MODE == 1) umin_row = P.umin + (size_t)b * P.G * P.Jpad + j;
            const float nsj = j < P.J ? P.ns[(size_t)b * P.J + j] : 0.f;
            const float xmb = P.xm[b];
            const bool aug = xmb < 0.f;
            const float margin = tc_margin(nsj, P.rmax[b], P.scale[b], P.C) + fmaxf(xmb, 0.f);
            if (MODE == 0 && P.prior != nullptr && !dead) {
                // x of the prior match from the same fp16 operands the tensor core sees.  The two fp32 accumulations (80
                // products each, different order) differ by less than the accumulation term of eps plus a quarter of it,
                // i.e. by less than margin = 2.04 eps: x' + margin bounds the accumulator value of that column, hence the
                // row minimum, from above, and everything within margin of the minimum is below x' + 2 margin.
                const long long kp = P.prior[(size_t)b * P.J + j];
                if (kp >= 0 && kp < P.K) {
                    const uint4 *pa = reinterpret_cast<const uint4 *>(P.a16 + ((size_t)b * P.J + j) * TC_CH);
                    const uint4 *pb = reinterpret_cast<const uint4 *>(P.b16 + ((size_t)b * P.K + kp) * TC_CH);
                    float acc = 0.f;
#pragma unroll
                    for (int v8 = 0; v8 < TC_CH / 8; ++v8) {
                        const uint4 ua = pa[v8], ub = pb[v8];
                        const __half2 *ha = reinterpret_cast<const __half2 *>(&ua), *hb = reinterpret_cast<const __half2 *>(&ub);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
                            acc = __fmaf_rn(fa.x, fb.x, acc);
                            acc = __fmaf_rn(fa.y, fb.y, acc);
                        }
                    }
                    const float sg = P.scale[b];
                    thr = (aug ? __fmaf_rn(P.nr[(size_t)b * P.K + kp], sg * sg, acc) : acc) + 2.0f * margin;
                }
            }
            for (int t = 0; t < us.total(); ++t) {
                const int u = us.unit(t);
                const bool sample = MODE == 1 || t < us.ns;
                if constexpr (MODE == 1) best = INFINITY;
                const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) +
                                       (uint32_t)((pa.stage * TC_RBS + r) * 128 + h * (TC_BN / TC_HALVES));
                mbar_wait(&tmem_full[pa.stage * TC_RBS + r], pa.phase);
                tc_fence_after();
                if (tracing && useq < 256) P.trace[1024 + useq * 8] = (unsigned int)clock64();
                const int col0 = u * TC_BN + h * (TC_BN / TC_HALVES);
                uint32_t va[32], vb[32];
                tmem_ld32(tbase, va);
#pragma unroll
                for (int g = 0; g < TC_STEPS; g += 2) {     // ping-pong: the next 32 columns fly during this step's math
                    tmem_wait32(va);
                    if (tracing && useq < 256 && g == 0) P.trace[1024 + useq * 8 + 1] = (unsigned int)clock64();
                    tmem_ld32(tbase + (g + 1) * 32, vb);
                    filter32<false, MODE>(va, col0 + g * 32, tbase + g * 32, margin, thr, cv, ci, sample, best, nullptr, 0, &tk);
                    if constexpr (MODE == 1)
                        if (P.gran32) { if (!dead) umin_row[(size_t)(u * TC_STEPS + g) * P.Jpad] = best; best = INFINITY; }
                    if (tracing && useq < 256 && g == 0) P.trace[1024 + useq * 8 + 2] = (unsigned int)clock64();
                    tmem_wait32(vb);
                    if (g + 2 < TC_STEPS) tmem_ld32(tbase + (g + 2) * 32, va);
                    if (tracing && useq < 256 && g == 0) P.trace[1024 + useq * 8 + 3] = (unsigned int)clock64();
                    if (g + 2 < TC_STEPS)
                        filter32<false, MODE>(vb, col0 + (g + 1) * 32, tbase + (g + 1) * 32, margin, thr, cv, ci, sample, best, nullptr,
                                              0, &tk);
                    else   // last step: the accumulator is handed back from inside (right after the vote)
                        filter32<true, MODE>(vb, col0 + (g + 1) * 32, tbase + (g + 1) * 32, margin, thr, cv, ci, sample, best,
                                             &tmem_empty[pa.stage * TC_RBS + r], lane, &tk);
                    if constexpr (MODE == 1)
                        if (P.gran32) { if (!dead) umin_row[(size_t)(u * TC_STEPS + g + 1) * P.Jpad] = best; best = INFINITY; }
                }
                if constexpr (MODE == 1)
                    if (!P.gran32 && !dead) umin_row[(size_t)u * P.Jpad] = best;
                if (tracing && useq < 256) P.trace[1024 + useq * 8 + 4] = (unsigned int)clock64();
                if (MODE == 0 && t == us.ns - 1 && !dead) thr = best + margin;   // primed: every row has seen a value <= best
                if (tracing && useq < 256) P.trace[1024 + useq * 8 + 5] = (unsigned int)clock64();
                ++useq;
                pa.advance(TC_ACC_STAGES);
            }
            if constexpr (MODE == 2) P.tk_cnt[(size_t)b * P.Jpad + j] = tk.n;
            if constexpr (MODE == 0) {
                const size_t slot = ((((size_t)b * P.Jpad + (size_t)j) * P.S + sp) * TC_LISTS + h) * TC_T;
                *reinterpret_cast<float4 *>(P.cand_val + slot) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                *reinterpret_cast<int4 *>(P.cand_idx + slot) = make_int4(ci[0], ci[1], ci[2], ci[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_TMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0 && P.dbg) {
        unsigned long long *d = P.dbg + (size_t)blockIdx.x * 4;
        d[0] = dbg_s[0]; d[1] = globaltimer_ns(); d[2] = clock64() - dbg_s[1];
        unsigned long long units = 0;
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            units += (unsigned long long)UnitSched(it % P.S, P.S, P.U, P.prime_div).nu;
        }
        d[3] = units;
    }
}

// ---------------------------------------------------------------------------------------------------------
// With (nearly) constant reference norms - the features are L2-normalised upstream, network/model.py:233-234 - the
// term sigma^2 |r_k|^2 shifts a whole row of x by a constant and cannot change the argmin: the filter then works on
// x' = -2 sigma^2 <s_j, r_k> (4 MMAs per tile instead of 5, no norm tile traffic) and widens every margin by the spread
// sigma^2 (max_k |r_k|^2 - min_k |r_k|^2): a column whose x is within margin of the row minimum of x has its x' within
// margin + spread of the minimum of x'.  Exactness is untouched (the refine step re-scores with the exact norms).
// Chosen per batch element on the device; beyond TC_NOAUG_SPREAD (2^-14, ~6 % of the smallest margin of unit features)
// the norm stays folded in - and so it does when K is not a multiple of the 128-column unit: the zero-filled reference rows
// beyond K rely on their huge padded norm to stay out of the candidate lists.  The per-batch minimum / maximum of the norms
// come out of the norm kernel (atomics); the first block of the reference-side prep kernel writes the decision.
// ---------------------------------------------------------------------------------------------------------
constexpr float TC_NOAUG_SPREAD = 6.1035e-5f;

__global__ void tc_init_kernel(int *count, float *rmax, float *amax, float *rmin, int B) {
    if (threadIdx.x < 2) count[threadIdx.x] = 0;
    if (rmax != nullptr)
        for (int b = threadIdx.x; b < B; b += blockDim.x) { rmax[b] = 0.f; amax[b] = 0.f; rmin[b] = __int_as_float(0x7f7f7f7f); }
}

// prep: [B,C,N] fp32 (any strides) -> fp16 tensor-core copy [B][N][64] = mul * sigma * f (point-major, channels C..63
// zero), and for the reference side the folded-norm channels [B][Npad][16] = {hi, lo, lolo, 0..} of sigma^2 |r|^2
// (TC_PAD_NORM beyond N).  One block = 32 points x 64 channels through a shared-memory transpose.  Block (0,0) also
// writes the constant source-side norm tile; the first block of every batch writes sigma.
__global__ __launch_bounds__(256) void tc_prep_kernel(dsir_feat f, int C, int N, int Npad, const float *__restrict__ amax,
                                                      const float *__restrict__ nrm, float mul, __half *__restrict__ out16,
                                                      __half *__restrict__ aug, __half *__restrict__ aug_const,
                                                      float *__restrict__ scale_out, const float *__restrict__ rmin = nullptr,
                                                      const float *__restrict__ rmax = nullptr, float *__restrict__ xm_out = nullptr) {
    __shared__ float tile[TC_CH][33];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float sigma = tc_sigma(amax[b]);
    const float *src = f.ptr + (size_t)b * f.batch_stride;
    if (n0 < N) {
#pragma unroll
        for (int c = ty; c < TC_CH; c += 8) {   // c: channel, tx: point (coalesced when the point stride is 1)
            const int n = n0 + tx;
            float v = 0.f;
            if (c < C && n < N) v = src[(size_t)c * f.chan_stride + (size_t)n * f.point_stride];
            tile[c][tx] = v;
        }
    }
    __syncthreads();
    const int i = threadIdx.x >> 3, g = threadIdx.x & 7;   // point inside the block, group of 8 channels
    const int n = n0 + i;
    if (n < N) {
        const float m = mul * sigma;   // power of two (times -2): the scaling is exact
        __align__(16) __half h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = __float2half_rn(tile[8 * g + k][i] * m);
        *reinterpret_cast<uint4 *>(out16 + ((size_t)b * N + n) * TC_CH + g * 8) = *reinterpret_cast<const uint4 *>(h);
    }
    if (aug && n < Npad && g < 2) {
        __align__(16) __half h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = __float2half_rn(0.f);
        if (g == 0) {
            if (n < N) {
                const float x = nrm[(size_t)b * N + n] * sigma * sigma;
                const __half hi = __float2half_rn(x);
                const float r1 = x - __half2float(hi);
                const __half lo = __float2half_rn(r1);
                const __half lolo = __float2half_rn(r1 - __half2float(lo));
                h[0] = hi; h[1] = lo; h[2] = lolo;
            } else {
                h[0] = __float2half_rn(TC_PAD_NORM);
            }
        }
        *reinterpret_cast<uint4 *>(aug + ((size_t)b * Npad + n) * TC_AUG + g * 8) = *reinterpret_cast<const uint4 *>(h);
    }
    if (aug_const && blockIdx.x == 0 && b == 0) {
        for (int t = threadIdx.x; t < 128 * TC_AUG; t += blockDim.x)
            aug_const[t] = __float2half_rn((t % TC_AUG) < 3 ? 1.f : 0.f);
    }
    if (scale_out && blockIdx.x == 0 && threadIdx.x == 0) scale_out[b] = sigma;
    if (xm_out && blockIdx.x == 0 && threadIdx.x == 0) {   // fold the norm in, or skip its K-step (see TC_NOAUG_SPREAD)
        const float spread = __fmul_rn(__fmul_rn(sigma, sigma), rmax[b] - rmin[b]) * 1.0001f;
        float x = (spread >= 0.f && spread <= TC_NOAUG_SPREAD && N % TC_BN == 0) ? spread : -1.f;   // NaN / inf norms keep the folded norm
#ifdef DSIR_TC_FORCE_AUG   // development builds: always fold the norm in (A/B of the 4-MMA mode)
        x = -1.f;
#endif
        xm_out[b] = x;
    }
}

// ---------------------------------------------------------------------------------------------------------
// refine: one thread per source row.  Candidates within margin of the row's approximate minimum are the only
// columns that can hold the exact fp32 minimum.  A single such candidate IS the answer (no arithmetic needed unless
// the caller wants the distance); several are re-scored with the exact fp32 op order of match_fp32.cu, reading the
// caller's own feature tensors (consecutive rows -> coalesced source reads; the reference side is a gather).
// ---------------------------------------------------------------------------------------------------------
struct RefineParams {
    int B, J, K, C, S, Jpad;
    dsir_feat fs, fr;
    const float *ns, *nr, *rmax, *scale, *xm;
    const float *cand_val;
    const int *cand_idx;
    int64_t *idx;
    float *min_d;
    int *rescue_count, *exact_count;   // adjacent ints
    int *rescue_rows;  // [B*J] flat row ids
    int *exact_rows;   // [B*J]
    unsigned long long *rescue_keys;  // [B*J] (ordered distance bits << 32) | index, atomicMin target
};

// dot = fma chain over channels ascending, d = ((-2 dot) + ns) + nr: bit-identical to match_fp32_kernel
__device__ __forceinline__ float exact_dist(const float *__restrict__ sp, int64_t s_cs, const float *__restrict__ rp, int64_t r_cs,
                                            int C, float nsj, float nrk) {
    float dot = 0.f;
#pragma unroll 32
    for (int c = 0; c < C; ++c) dot = __fmaf_rn(sp[(size_t)c * s_cs], rp[(size_t)c * r_cs], dot);
    return l2_from_dot(dot, nsj, nrk);
}

// classify: one thread per source row
__global__ __launch_bounds__(256) void match_tc_refine_kernel(RefineParams P) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)P.B * P.J) return;
    const int b = (int)(row / P.J), j = (int)(row % P.J);
    const int nlist = P.S * TC_LISTS;
    const int ncand = nlist * TC_T;
    const float4 *cvp = reinterpret_cast<const float4 *>(P.cand_val + ((size_t)b * P.Jpad + j) * ncand);
    const int4 *cip = reinterpret_cast<const int4 *>(P.cand_idx + ((size_t)b * P.Jpad + j) * ncand);
    const float nsj = P.ns[(size_t)b * P.J + j];
    const float margin = tc_margin(nsj, P.rmax[b], P.scale[b], P.C) + fmaxf(P.xm[b], 0.f);   // candidate values are in scaled units
    // pass 1: approximate row minimum over the valid candidates
    float gmin = INFINITY;
    for (int s = 0; s < nlist; ++s) {
        const float4 v = cvp[s];
        const int4 k = cip[s];
        const float vv[4] = {v.x, v.y, v.z, v.w};
        const int kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int t = 0; t < TC_T; ++t)
            if (kk[t] >= 0 && kk[t] < P.K && vv[t] < 1e38f) gmin = fminf(gmin, vv[t]);
    }
    // pass 2: who qualifies; a list whose LAST slot still qualifies may have dropped something -> rescue
    const float lim = gmin + margin;
    int ntake = 0, ksingle = 0;
    bool sat = false;
    for (int s = 0; s < nlist; ++s) {
        const float4 v = cvp[s];
        const int4 k = cip[s];
        const float vv[4] = {v.x, v.y, v.z, v.w};
        const int kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int t = 0; t < TC_T; ++t) {
            const bool take = kk[t] >= 0 && kk[t] < P.K && vv[t] < 1e38f && vv[t] <= lim;
            if (take) { ++ntake; ksingle = kk[t]; sat = sat || (t == TC_T - 1); }
        }
    }
    if (ntake == 0 || sat) {                       // exhaustive fp32 scan
        P.idx[row] = 0;
        if (P.min_d) P.min_d[row] = INFINITY;
        const int pos = atomicAdd(P.rescue_count, 1);
        P.rescue_rows[pos] = (int)row;
        P.rescue_keys[pos] = ~0ull;
    } else if (ntake == 1 && P.min_d == nullptr) {  // the only possible argmin: no arithmetic needed
        P.idx[row] = (int64_t)ksingle;
    } else {                                        // several candidates (or the distance is wanted): exact re-scoring
        P.idx[row] = 0;
        const int pos = atomicAdd(P.exact_count, 1);
        P.exact_rows[pos] = (int)row;
    }
}

// exact re-scoring of the listed rows: one lane per candidate slot (S x 4 <= 32 slots per row, hence 32 / LPR rows per warp:
// eight at S = 1), every qualifying lane walks its own fma chain; the (value, index) lexicographic minimum of a row's lanes
// wins, exactly like match_fp32_kernel.
template <int LPR>   // lanes per row: the candidate slots of a row (S x 4) rounded up to a power of two; 32 / LPR rows per warp
__global__ __launch_bounds__(256) void match_tc_exact_kernel(RefineParams P) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int count = *P.exact_count;
    const int nlist = P.S * TC_LISTS, ncand = nlist * TC_T;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int i0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; i0 < count; i0 += warps * RPW) {
        const int i = i0 + sub;
        const bool live = i < count;
        const int row = live ? P.exact_rows[i] : 0;
        const int b = row / P.J, j = row % P.J;
        const size_t cbase = ((size_t)b * P.Jpad + j) * ncand;
        float v = INFINITY;
        int k = -1;
        if (live && sl < ncand) { v = P.cand_val[cbase + sl]; k = P.cand_idx[cbase + sl]; }
        const bool valid = k >= 0 && k < P.K && v < 1e38f;
        float gmin = valid ? v : INFINITY;
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        const float nsj = P.ns[(size_t)b * P.J + j];
        const float margin = tc_margin(nsj, P.rmax[b], P.scale[b], P.C) + fmaxf(P.xm[b], 0.f);
        const bool take = valid && v <= gmin + margin;
        float d = INFINITY;
        int kk = 0x7fffffff;
        if (take) {
            d = exact_dist(P.fs.ptr + (size_t)b * P.fs.batch_stride + (size_t)j * P.fs.point_stride, P.fs.chan_stride,
                           P.fr.ptr + (size_t)b * P.fr.batch_stride + (size_t)k * P.fr.point_stride, P.fr.chan_stride, P.C, nsj,
                           P.nr[(size_t)b * P.K + k]);
            kk = k;
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
            const int k2 = __shfl_xor_sync(0xffffffffu, kk, o);
            if (d2 < d || (d2 == d && k2 < kk)) { d = d2; kk = k2; }
        }
        if (live && sl == 0) {
            const bool rescue = !(d < INFINITY);   // every qualifying distance was NaN/inf: let the exhaustive path decide
            P.idx[row] = rescue ? 0 : (int64_t)kk;
            if (P.min_d) P.min_d[row] = d;
            if (rescue) {
                const int pos = atomicAdd(P.rescue_count, 1);
                P.rescue_rows[pos] = row;
                P.rescue_keys[pos] = ~0ull;
            }
        }
    }
}

// rescue: exhaustive exact fp32 scan of the listed rows.  Work unit = (listed row, 2048-column chunk) so that a
// handful of rows still spreads over the whole chip; partial results meet in a 64-bit atomicMin whose key orders
// (distance, index) lexicographically.
__device__ __forceinline__ unsigned int float_order_bits(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_bits(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
constexpr int RESCUE_CHUNK = 256;    // one column per thread: the 64 strided loads of a thread are independent
constexpr int RESCUE_MAXC = 64;

__global__ __launch_bounds__(256) void match_tc_rescue_kernel(RefineParams P, unsigned long long *keys) {
    __shared__ float srow[RESCUE_MAXC];
    __shared__ unsigned long long red[8];
    const int count = *P.rescue_count;
    const int nchunk = (P.K + RESCUE_CHUNK - 1) / RESCUE_CHUNK;
    const long long units = (long long)count * nchunk;
    for (long long uidx = blockIdx.x; uidx < units; uidx += gridDim.x) {
        const int i = (int)(uidx / nchunk), ch = (int)(uidx % nchunk);
        const int row = P.rescue_rows[i];
        const int b = row / P.J, j = row % P.J;
        __syncthreads();
        for (int c = threadIdx.x; c < P.C; c += blockDim.x)
            srow[c] = P.fs.ptr[(size_t)b * P.fs.batch_stride + (size_t)c * P.fs.chan_stride + (size_t)j * P.fs.point_stride];
        __syncthreads();
        const float nsj = P.ns[(size_t)b * P.J + j];
        const float *rb = P.fr.ptr + (size_t)b * P.fr.batch_stride;
        unsigned long long best = ~0ull;
        const int kend = min(P.K, (ch + 1) * RESCUE_CHUNK);
        for (int k = ch * RESCUE_CHUNK + threadIdx.x; k < kend; k += blockDim.x) {   // coalesced over k when point stride is 1
            const float d = exact_dist(srow, 1, rb + (size_t)k * P.fr.point_stride, P.fr.chan_stride, P.C, nsj,
                                       P.nr[(size_t)b * P.K + k]);
            if (d == d) {
                unsigned long long key = ((unsigned long long)float_order_bits(d) << 32) | (unsigned int)k;
                best = key < best ? key : best;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) best = red[w] < best ? red[w] : best;
            if (best != ~0ull) atomicMin(&keys[i], best);
        }
    }
}

__global__ void match_tc_rescue_finalize_kernel(RefineParams P, const unsigned long long *keys) {
    const int count = *P.rescue_count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const int row = P.rescue_rows[i];
        const unsigned long long key = keys[i];
        const bool none = key == ~0ull;  // every distance was NaN: same answer as the fp32 kernel (index 0, +inf)
        P.idx[row] = none ? 0 : (int64_t)(unsigned int)(key & 0xffffffffull);
        if (P.min_d) P.min_d[row] = none ? INFINITY : float_from_order_bits((unsigned int)(key >> 32));
    }
}

struct TcPlan {
    int NKS, RB, U, S, Jpad, Kpad;
    size_t off_a16, off_b16, off_baug, off_aaug, off_rmax, off_amax, off_scale, off_xm, off_rmin, off_cval, off_cidx, off_count,
        off_rows, off_erows, off_keys, off_dbg, off_trace, total;
    int gran = 0, G = 0, cap = 0;                              // top-k sweeps (launch_match_tc_topk)
    size_t off_umin = 0, off_thr = 0, off_tkcnt = 0, off_tklist = 0, off_frt = 0, off_exrows = 0;
};

// top-k sweeps: column granule of the sweep-1 minima.  The k-th smallest of G granule minima bounds the k-th smallest
// element from above; with G >= 1.25 k the expected number of columns at or below it stays near 2 k (iid columns:
// (1-p)^gran = 1 - k/G), within the 4 k list.  0 = the fused path does not apply (too few columns).
int topk_granule(int K, int topk) {
    if (topk <= 0) return 0;
    if (4ll * ((K + 127) / 128) >= 5ll * topk) return 128;
    if (4ll * ((K + 31) / 32) >= 5ll * topk) return 32;
    return 0;
}
int topk_cap(int topk) { return topk * 4 < 32 ? 32 : topk * 4; }

TcPlan make_plan(int B, int C, int J, int K, int topk = 0) {
    TcPlan p;
    p.NKS = (C + 15) / 16;
    p.RB = (J + TC_BM - 1) / TC_BM;
    p.U = (K + TC_BN - 1) / TC_BN;
    p.Jpad = p.RB * TC_BM;
    p.Kpad = p.U * TC_BN;
    // K-splits: S work items per (batch element, 512-row block).  The persistent grid runs the items in rounds of one per
    // SM; a round costs what ONE item costs: its share of the units, the priming pass over an eighth of them, and a fixed
    // part (A tiles, cold threshold, list write-back - TC_ITEM_OVERHEAD units, fitted to tools/sweep_split.py).  Take the S
    // with the cheapest total; more splits only when they pay (every split re-primes and re-discovers its minima).
    long long items = (long long)B * p.RB;
    int S = 1;
    {
        const int sms = 148;
        double best = 0.0;
        for (int s = 1; s <= TC_MAX_SPLIT && s <= p.U; ++s) {
            const long long rounds = (items * s + sms - 1) / sms;
            const int nu = (p.U + s - 1) / s;
            const double cost = (double)rounds * (nu + (nu >= 16 ? nu / 8 : 0) + TC_ITEM_OVERHEAD);
            if (s == 1 || cost < best * 0.97) { best = cost; S = s; }   // a further split must win by 3 %
        }
    }
#ifdef DSIR_TC_TRACE
    { const char *e = getenv("DSIR_TC_SPLIT"); if (e && atoi(e) > 0) S = atoi(e) < p.U ? atoi(e) : p.U; if (S > TC_MAX_SPLIT) S = TC_MAX_SPLIT; }
#endif
    if (topk > 0) S = 1;   // top-k sweeps: one owner thread per row (list position in a register, no atomics)
    p.S = S;   // (splitting further to fill the last wave was measured slower: every split re-primes and re-discovers its minima)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += ws_block(bytes); return o; };
    p.off_a16 = take((size_t)B * J * TC_CH * 2);
    p.off_b16 = take((size_t)B * K * TC_CH * 2);
    p.off_baug = take((size_t)B * p.Kpad * TC_AUG * 2);
    p.off_aaug = take((size_t)128 * TC_AUG * 2);
    p.off_rmax = take((size_t)B * 4);
    p.off_amax = take((size_t)B * 4);
    p.off_scale = take((size_t)B * 4);
    p.off_xm = take((size_t)B * 4);
    p.off_rmin = take((size_t)B * 4);
    p.off_cval = take((size_t)B * p.Jpad * S * TC_LISTS * TC_T * 4);
    p.off_cidx = take((size_t)B * p.Jpad * S * TC_LISTS * TC_T * 4);
    p.off_count = take(256);
    p.off_rows = take((size_t)B * J * 4);
    p.off_erows = take((size_t)B * J * 4);
    p.off_keys = take((size_t)B * J * 8);
    p.off_dbg = take((size_t)256 * 4 * 8);
    p.off_trace = take((size_t)4096 * 4);
    p.gran = topk_granule(K, topk);
    if (p.gran) {
        p.G = p.Kpad / p.gran;
        p.cap = topk_cap(topk);
        p.off_umin = take((size_t)B * p.G * p.Jpad * 4);
        p.off_thr = take((size_t)B * p.Jpad * 4);
        p.off_tkcnt = take((size_t)B * p.Jpad * 4);
        p.off_tklist = take((size_t)B * p.Jpad * p.cap * 4);
        p.off_frt = take((size_t)B * K * ((C + 3) / 4 * 4) * 4);
        p.off_exrows = take((size_t)B * J * 4 + 256);     // [0] = count, rows from +64 ints on
    }
    p.total = off + 1024;
    return p;
}

constexpr size_t filter_smem_bytes() {
    return 1024 + (size_t)(TC_RBS + TC_STAGES) * MAIN_TILE + (size_t)(1 + TC_STAGES) * AUG_TILE + 512;
}

}  // namespace

bool match_tc_supported(const dsir_feat &fs, const dsir_feat &fr, int B, int C, int J, int K) {
    (void)fs; (void)fr;
    if (C < 1 || C > TC_CH) return false;
    if ((long long)B * J >= (1ll << 31) || (long long)B * K >= (1ll << 31)) return false;
    return tc_encode_fn() != nullptr;
}

bool match_tc_profitable(int B, int C, int J, int K) {
    (void)C;
    return (double)B * J * K >= 4.0e6;  // below this the prep/refine launches dominate
}

size_t match_tc_workspace_bytes(int B, int C, int J, int K) {
    if (C > TC_CH) return 0;
    return make_plan(B, C, J, K).total;
}

namespace {

// norms, maxima, fp16 operand copies and tensor maps of one (features, workspace) pair; fills the common part of TcParams
struct TcReady {
    CUtensorMap mapA, mapB, mapAaug, mapBaug;
    TcParams T;
    int grid, sms;
};

int tc_prepare(const MatchParams &P, const TcPlan &pl, char *base, TcReady &R, cudaStream_t st) {
    __half *a16 = (__half *)(base + pl.off_a16), *b16 = (__half *)(base + pl.off_b16);
    __half *baug = (__half *)(base + pl.off_baug), *aaug = (__half *)(base + pl.off_aaug);
    float *rmax = (float *)(base + pl.off_rmax), *amax = (float *)(base + pl.off_amax), *scale = (float *)(base + pl.off_scale);
    float *xm = (float *)(base + pl.off_xm), *rmin = (float *)(base + pl.off_rmin);
    int *count = (int *)(base + pl.off_count);
    // one tiny launch instead of three memsets: row counters, and (first call on these features) the per-batch extrema
    tc_init_kernel<<<1, 256, 0, st>>>(count, P.reuse_prep ? nullptr : rmax, amax, rmin, P.B);
    DSIR_LAUNCH_CHECK();
    int rc;
    if (!P.reuse_prep) {
        // exact squared norms (fma chains, shared with the fp32 kernel) + per-batch maxima for sigma and the margin
        if ((rc = launch_sqnorm(P.fr, P.B, P.C, P.K, const_cast<float *>(P.nr), (int *)rmax, (int *)amax, (int *)rmin, st))) return rc;
        if ((rc = launch_sqnorm(P.fs, P.B, P.C, P.J, const_cast<float *>(P.ns), (int *)amax, nullptr, st))) return rc;
        tc_prep_kernel<<<dim3(cdiv(P.J, 32), P.B), 256, 0, st>>>(P.fs, P.C, P.J, P.J, amax, nullptr, 1.0f, a16, nullptr, aaug, scale);
        DSIR_LAUNCH_CHECK();
        tc_prep_kernel<<<dim3(pl.Kpad / 32, P.B), 256, 0, st>>>(P.fr, P.C, P.K, pl.Kpad, amax, P.nr, -2.0f, b16, baug, nullptr, nullptr,
                                                                rmin, rmax, xm);
        DSIR_LAUNCH_CHECK();
    }
    if (!make_f16_tmap(&R.mapA, a16, P.B, P.J, TC_CH, TC_CH, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_f16_tmap(&R.mapB, b16, P.B, P.K, TC_CH, TC_CH, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !make_f16_tmap(&R.mapAaug, aaug, 1, 128, TC_AUG, TC_AUG, CU_TENSOR_MAP_SWIZZLE_32B) ||
        !make_f16_tmap(&R.mapBaug, baug, P.B, pl.Kpad, TC_AUG, TC_AUG, CU_TENSOR_MAP_SWIZZLE_32B))
        return DSIR_ERR_UNSUPPORTED;
    TcParams &T = R.T;
    T = TcParams{};
    T.B = P.B; T.J = P.J; T.K = P.K; T.C = P.C; T.RB = pl.RB; T.U = pl.U; T.S = pl.S;
    T.Jpad = pl.Jpad; T.Kpad = pl.Kpad; T.ns = P.ns; T.rmax = rmax; T.scale = scale; T.xm = xm;
    T.cand_val = (float *)(base + pl.off_cval); T.cand_idx = (int *)(base + pl.off_cidx);
    T.dbg = (unsigned long long *)(base + pl.off_dbg);
    T.trace = (unsigned int *)(base + pl.off_trace);
    T.a16 = a16; T.b16 = b16; T.nr = P.nr;
    const int items = P.B * pl.RB * pl.S;
    int dev = 0;
    R.sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&R.sms, cudaDevAttrMultiProcessorCount, dev);
    R.grid = items < R.sms ? items : (R.sms > 256 ? 256 : R.sms);
    return DSIR_OK;
}

template <int MODE>
int tc_launch_filter(const TcPlan &pl, const TcReady &R, cudaStream_t st) {
    const size_t smem = filter_smem_bytes();
#define DSIR_TC_LAUNCH(NKS)                                                                                                   \
    do {                                                                                                                      \
        DSIR_CUDA_TRY(cudaFuncSetAttribute(match_tc_filter_kernel<NKS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        match_tc_filter_kernel<NKS, MODE><<<R.grid, TC_THREADS, smem, st>>>(R.mapA, R.mapB, R.mapAaug, R.mapBaug, R.T);          \
    } while (0)
    switch (pl.NKS) {
        case 1: DSIR_TC_LAUNCH(1); break;
        case 2: DSIR_TC_LAUNCH(2); break;
        case 3: DSIR_TC_LAUNCH(3); break;
        default: DSIR_TC_LAUNCH(4); break;
    }
#undef DSIR_TC_LAUNCH
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

}  // namespace

int launch_match_tc(const MatchParams &P, void *ws, size_t ws_bytes, cudaStream_t st) {
    const TcPlan pl = make_plan(P.B, P.C, P.J, P.K);
    char *base = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    if (ws == nullptr || (size_t)(base - (char *)ws) + pl.total - 1024 > ws_bytes) return DSIR_ERR_WORKSPACE;
    TcReady Rdy;
    int rc;
    if ((rc = tc_prepare(P, pl, base, Rdy, st))) return rc;
    TcParams &T = Rdy.T;
    T.dbg_flags = 0;
    T.prime_div = 8;
#ifdef DSIR_TC_TRACE
    { const char *e = getenv("DSIR_TC_DEBUG"); T.dbg_flags = e ? atoi(e) : 0; }
    { const char *e = getenv("DSIR_TC_PRIME"); T.prime_div = e ? atoi(e) : 8; }
#endif
    T.prior = P.prior_idx;
    if (T.prior) T.prime_div = 0;
    if ((rc = tc_launch_filter<0>(pl, Rdy, st))) return rc;
    const int sms = Rdy.sms;
    float *rmax = (float *)(base + pl.off_rmax), *scale = (float *)(base + pl.off_scale), *xm = (float *)(base + pl.off_xm);
    float *cval = T.cand_val;
    int *cidx = T.cand_idx, *count = (int *)(base + pl.off_count), *rows = (int *)(base + pl.off_rows);

    RefineParams R{};
    R.B = P.B; R.J = P.J; R.K = P.K; R.C = P.C; R.S = pl.S; R.Jpad = pl.Jpad;
    R.fs = P.fs; R.fr = P.fr; R.ns = P.ns; R.nr = P.nr; R.rmax = rmax; R.scale = scale; R.xm = xm; R.cand_val = cval; R.cand_idx = cidx;
    R.idx = P.idx; R.min_d = P.min_d; R.rescue_count = count; R.exact_count = count + 1; R.rescue_rows = rows;
    R.exact_rows = (int *)(base + pl.off_erows);
    R.rescue_keys = (unsigned long long *)(base + pl.off_keys);
    const long long nrows = (long long)P.B * P.J;
    match_tc_refine_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(R);
    DSIR_LAUNCH_CHECK();
    {   // one lane per candidate slot: S x 4 slots per row -> 8, 4, 2 or 1 rows per warp
        const int ncand = pl.S * TC_LISTS * TC_T;
        if (ncand <= 4) match_tc_exact_kernel<4><<<sms * 8, 256, 0, st>>>(R);
        else if (ncand <= 8) match_tc_exact_kernel<8><<<sms * 8, 256, 0, st>>>(R);
        else if (ncand <= 16) match_tc_exact_kernel<16><<<sms * 8, 256, 0, st>>>(R);
        else match_tc_exact_kernel<32><<<sms * 8, 256, 0, st>>>(R);
    }
    DSIR_LAUNCH_CHECK();
    match_tc_rescue_kernel<<<sms * 16, 256, 0, st>>>(R, R.rescue_keys);   // (row, 256-column chunk) units: one per CTA for a few rows
    DSIR_LAUNCH_CHECK();
    match_tc_rescue_finalize_kernel<<<sms, 256, 0, st>>>(R, R.rescue_keys);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Top-k soft correspondences without a score matrix (north_star: "row-wise online softmax / top-k ... never materialised").
// With a_jk = -beta (d_jk - alpha) and beta > 0 the k largest weights of a row are its k smallest distances:
//   sweep 1 (MODE 1)  the tensor-core pass, value only: x-minima of every column granule (128 or 32 columns) of every row;
//   threshold         tau_j = k-th smallest granule minimum of the row: k different columns have x_hat <= tau_j, so the exact
//                     k-th smallest x is <= tau_j + eps and every column of the exact top-k set has x_hat <= tau_j + 2 eps;
//   sweep 2 (MODE 2)  the same pass again, appending every column with x_hat <= tau_j + margin (+ a round-off allowance for
//                     the map d -> a) to the row's list (4 k entries);
//   exact             one warp per row re-scores the listed columns with the fp32 op order of match_fp32.cu and takes the k
//                     largest a (ties to the lower index), w = exp(a - lse).  Rows whose list overflowed or holds fewer than
//                     k columns (ties en masse, beta <= 0, NaN rows) are scanned exhaustively by the same warp.
// Outputs are those of the materialising path (DENSE chunk + row_topk_kernel), bit for bit.
// ---------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ float warp_sort32(float v, int lane) {   // ascending over the lanes (bitonic network)
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, v, stride);
            const bool up = (lane & size) == 0, lower = (lane & stride) == 0;
            v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}
__device__ __forceinline__ float warp_merge32(float v, int lane) {  // bitonic sequence -> ascending
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, stride);
        v = (lane & stride) == 0 ? fminf(v, o) : fmaxf(v, o);
    }
    return v;
}

// one warp per row: the 32 smallest granule minima live one per lane (ascending); thr = k-th + margin + allowance
__global__ __launch_bounds__(256) void topk_thr_kernel(TcParams P, int topk, const float *__restrict__ beta,
                                                       const float *__restrict__ alpha, float *__restrict__ thr_out,
                                                       int *__restrict__ cnt_out, int *__restrict__ ex_count) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *ex_count = 0;
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= (long long)P.B * P.J) return;
    const int b = (int)(row / P.J), j = (int)(row % P.J);
    const float *u = P.umin + (size_t)b * P.G * P.Jpad + j;
    float best = INFINITY;                                  // lane l: the (l+1)-th smallest so far
    for (int g0 = 0; g0 < P.G; g0 += 32) {
        float v = g0 + lane < P.G ? u[(size_t)(g0 + lane) * P.Jpad] : INFINITY;
        v = v == v ? v : INFINITY;                          // NaN columns never count
        v = warp_sort32(v, lane);
        const float rev = __shfl_sync(0xffffffffu, v, 31 - lane);
        best = warp_merge32(fminf(best, rev), lane);        // ascending vs descending: the element-wise minimum is the 32 smallest
    }
    const float tau = __shfl_sync(0xffffffffu, best, topk - 1);
    if (lane == 0) {
        const float nsj = P.ns[(size_t)b * P.J + j], sg = P.scale[b], rm = P.rmax[b];
        const float margin = tc_margin(nsj, rm, sg, P.C) + fmaxf(P.xm[b], 0.f);
        // columns whose distances differ by a few ulps of (|alpha| + d) can tie in a = -beta (d - alpha): keep them all
        const float allow = sg * sg * 4.8e-7f * (fabsf(alpha[b]) + nsj + rm + 2.f * sqrtf(nsj * rm));
        float t = tau + margin + allow;
        if (!(beta[b] > 0.f)) t = INFINITY;                 // not a distance order: everything to the exhaustive pass
        thr_out[(size_t)b * P.Jpad + j] = t;
        cnt_out[(size_t)b * P.Jpad + j] = 0;
    }
}

// [B,C,N] (any strides) -> point-major fp32 copy [B][N][C4] (C4 = C rounded up to 4, zero filled): a listed column's
// channels become 16-byte loads of one or two lines instead of C scattered sectors
__global__ __launch_bounds__(256) void feat_point_major_kernel(dsir_feat f, int C, int N, int C4, float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = f.ptr + (size_t)b * f.batch_stride;
    for (int c = ty; c < 32; c += 8) {
        const int n = n0 + tx;
        tile[c][tx] = (c0 + c < C && n < N) ? src[(size_t)(c0 + c) * f.chan_stride + (size_t)n * f.point_stride] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, c = c0 + tx;
        if (n < N && c < C4) out[((size_t)b * N + n) * C4 + c] = tile[tx][i];
    }
}

// <s_j, r_k> over the point-major copy: the fma chain over ascending channels of match_fp32.cu (the zero-filled tail
// channels add exact zeros)
__device__ __forceinline__ float dot_point_major(const float *srow, const float4 *__restrict__ rp, int C4) {
    float dot = 0.f;
#pragma unroll 8      // eight 16-byte loads in flight; the fma chain stays in channel order
    for (int c4 = 0; c4 < C4 / 4; ++c4) {
        const float4 r = rp[c4];
        const float4 sv = *reinterpret_cast<const float4 *>(srow + 4 * c4);
        dot = __fmaf_rn(sv.x, r.x, dot); dot = __fmaf_rn(sv.y, r.y, dot);
        dot = __fmaf_rn(sv.z, r.z, dot); dot = __fmaf_rn(sv.w, r.w, dot);
    }
    return dot;
}

// rows whose list holds between k and cap columns (nearly all): every lane re-scores at most LPL listed columns (fp32 op
// order of match_fp32.cu: fma chain over the channels, ((-2 dot) + |s|^2) + |r|^2), keeps them sorted, and k rounds of a
// warp arg-max over the lane heads emit the row's top-k (a descending, ties to the lower index).
template <int LPL>
__global__ __launch_bounds__(256) void topk_listed_kernel(RefineParams P, const int *__restrict__ cnt, const int *__restrict__ list,
                                                          int cap, int topk, const float *__restrict__ frt, int C4,
                                                          const float *__restrict__ beta, const float *__restrict__ alpha,
                                                          const float *__restrict__ lse, int64_t *__restrict__ out_idx,
                                                          float *__restrict__ out_w, int *__restrict__ ex) {
    __shared__ __align__(16) float srow[8][TC_CH];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= (long long)P.B * P.J) return;
    const int b = (int)(row / P.J), j = (int)(row % P.J);
    const size_t prow = (size_t)b * P.Jpad + j;
    const int n = cnt[prow];
    if (n > cap || n < topk) {                             // topk_exact_kernel scans these exhaustively
        if (lane == 0) ex[64 + atomicAdd(ex, 1)] = (int)row;
        return;
    }
    const float *sp = P.fs.ptr + (size_t)b * P.fs.batch_stride + (size_t)j * P.fs.point_stride;
    for (int c = lane; c < C4; c += 32) srow[w][c] = c < P.C ? sp[(size_t)c * P.fs.chan_stride] : 0.f;
    __syncwarp();
    const int *lst = list + prow * cap;
    const float nb = -beta[b], al = alpha[b];
    const float nsj = P.ns[(size_t)b * P.J + j];
    float av[LPL];
    int ai[LPL];
#pragma unroll
    for (int q = 0; q < LPL; ++q) {
        const int p = lane + 32 * q;
        av[q] = -INFINITY; ai[q] = 0x7fffffff;
        if (p < n) {
            const int k = lst[p];
            const float4 *rp = reinterpret_cast<const float4 *>(frt + ((size_t)b * P.K + k) * C4);
            const float dot = dot_point_major(srow[w], rp, C4);
            const float a = nb * (l2_from_dot(dot, nsj, P.nr[(size_t)b * P.K + k]) - al);
            if (a == a) { av[q] = a + 0.f; ai[q] = k; }    // NaN scores are never selected (like the materialising route); -0 -> +0: one bit pattern per value
        }
    }
    // order-preserving bits: the arg-max over the lane heads is one REDUX for the value and one for the (lower) index
    unsigned int hv[LPL];
#pragma unroll
    for (int q = 0; q < LPL; ++q) hv[q] = float_order_bits(av[q]);
    // sort the lane's LPL entries, descending (a, -index)
#pragma unroll
    for (int x = 0; x < LPL; ++x)
#pragma unroll
        for (int y = 0; y + 1 < LPL - x; ++y) {
            const bool sw = hv[y + 1] > hv[y] || (hv[y + 1] == hv[y] && ai[y + 1] < ai[y]);
            const unsigned int tv = sw ? hv[y] : hv[y + 1]; const int ti = sw ? ai[y] : ai[y + 1];
            hv[y] = sw ? hv[y + 1] : hv[y]; ai[y] = sw ? ai[y + 1] : ai[y];
            hv[y + 1] = tv; ai[y + 1] = ti;
        }
    const float l = lse[(size_t)b * P.J + j];
    int64_t *oi = out_idx + ((size_t)b * P.J + j) * topk;
    float *ow = out_w + ((size_t)b * P.J + j) * topk;
    unsigned int myv = 0;
    int myi = 0x7fffffff;
    constexpr unsigned int PAD = 0x007fffffu;              // float_order_bits(-inf)
    for (int t = 0; t < topk; ++t) {
        const unsigned int m = __reduce_max_sync(0xffffffffu, hv[0]);
        const int i = __reduce_min_sync(0xffffffffu, hv[0] == m ? ai[0] : 0x7fffffff);
        if (hv[0] == m && ai[0] == i) {                    // the winner pops its head (padding entries may pop together: harmless)
#pragma unroll
            for (int q = 0; q < LPL - 1; ++q) { hv[q] = hv[q + 1]; ai[q] = ai[q + 1]; }
            hv[LPL - 1] = PAD; ai[LPL - 1] = 0x7fffffff;
        }
        if (lane == t) { myv = m; myi = i; }               // lane t keeps result t: one coalesced store per row
    }
    if (lane < topk) {
        oi[lane] = myi == 0x7fffffff ? (int64_t)-1 : (int64_t)myi;
        ow[lane] = myi == 0x7fffffff ? 0.f : expf(float_from_order_bits(myv) - l);
    }
}

// the rows topk_listed_kernel set aside (list overflowed or short: mass ties, beta <= 0, NaN rows): one CTA per row, the
// grid strides over the list.  Every warp scans an eighth of the columns (per-lane sorted lists of KMAX, the only bound that
// holds for any distribution of the winners over the lanes), emits its own top-k by warp arg-max rounds, warp 0 merges the
// eight sorted lists.  Same order as everywhere: a descending, ties to the lower index; NaN scores never enter.
template <int KMAX>
__global__ __launch_bounds__(256) void topk_exact_kernel(RefineParams P, int topk, const float *__restrict__ frt, int C4,
                                                         const float *__restrict__ beta, const float *__restrict__ alpha,
                                                         const float *__restrict__ lse, int64_t *__restrict__ out_idx,
                                                         float *__restrict__ out_w, const int *__restrict__ ex) {
    __shared__ __align__(16) float srow[TC_CH];
    __shared__ float sm_a[8][32];
    __shared__ int sm_i[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_ex = ex[0];
    for (int it = blockIdx.x; it < n_ex; it += gridDim.x) {
        const int row = ex[64 + it];
        const int b = row / P.J, j = row % P.J;
        __syncthreads();                                   // the previous row is finished with the shared arrays
        const float *sp = P.fs.ptr + (size_t)b * P.fs.batch_stride + (size_t)j * P.fs.point_stride;
        for (int c = threadIdx.x; c < C4; c += blockDim.x) srow[c] = c < P.C ? sp[(size_t)c * P.fs.chan_stride] : 0.f;
        __syncthreads();
        const float nb = -beta[b], al = alpha[b];
        const float nsj = P.ns[(size_t)b * P.J + j];
        float bv[KMAX];
        int bi[KMAX];
#pragma unroll
        for (int p = 0; p < KMAX; ++p) { bv[p] = -INFINITY; bi[p] = 0x7fffffff; }
        for (int k = threadIdx.x; k < P.K; k += 256) {
            const float dot = dot_point_major(srow, reinterpret_cast<const float4 *>(frt + ((size_t)b * P.K + k) * C4), C4);
            const float a = nb * (l2_from_dot(dot, nsj, P.nr[(size_t)b * P.K + k]) - al);
            if (a > bv[KMAX - 1] || (a == bv[KMAX - 1] && k < bi[KMAX - 1])) {   // sorted insertion, descending (value, -index)
#pragma unroll
                for (int q = KMAX - 1; q >= 0; --q) {
                    const int qm = q > 0 ? q - 1 : 0;
                    const bool shift = (q > 0) && (a > bv[qm] || (a == bv[qm] && k < bi[qm]));
                    const bool here = !shift && (a > bv[q] || (a == bv[q] && k < bi[q]));
                    bv[q] = shift ? bv[qm] : (here ? a : bv[q]);
                    bi[q] = shift ? bi[qm] : (here ? k : bi[q]);
                }
            }
        }
        for (int t = 0; t < topk; ++t) {      // this warp's t-th best: warp arg-max over the lane heads
            float v = bv[0];
            int i = bi[0], src = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
                const int i2 = __shfl_xor_sync(0xffffffffu, i, o), s2 = __shfl_xor_sync(0xffffffffu, src, o);
                if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; src = s2; }
            }
            if (lane == src) {
#pragma unroll
                for (int q = 0; q < KMAX - 1; ++q) { bv[q] = bv[q + 1]; bi[q] = bi[q + 1]; }
                bv[KMAX - 1] = -INFINITY; bi[KMAX - 1] = 0x7fffffff;
            }
            if (lane == 0) { sm_a[w][t] = v; sm_i[w][t] = i; }
        }
        __syncthreads();
        if (w == 0) {                          // merge the eight sorted lists: lane l < 8 walks list l
            const float l = lse[(size_t)b * P.J + j];
            int64_t *oi = out_idx + ((size_t)b * P.J + j) * topk;
            float *ow = out_w + ((size_t)b * P.J + j) * topk;
            int pos = 0;
            for (int t = 0; t < topk; ++t) {
                const bool have = lane < 8 && pos < topk;
                float v = have ? sm_a[lane][pos] : -INFINITY;
                int i = have ? sm_i[lane][pos] : 0x7fffffff, src = lane;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
                    const int i2 = __shfl_xor_sync(0xffffffffu, i, o), s2 = __shfl_xor_sync(0xffffffffu, src, o);
                    if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; src = s2; }
                }
                if (lane == src) ++pos;        // (lanes 0..7 agree on the winner: the xor tree over 4, 2, 1 stays inside them)
                if (lane == 0) {
                    oi[t] = i == 0x7fffffff ? (int64_t)-1 : (int64_t)i;
                    ow[t] = i == 0x7fffffff ? 0.f : expf(v - l);
                }
            }
        }
    }
}

}  // namespace

bool match_tc_topk_supported(int B, int C, int J, int K, int topk) {
    if (topk < 1 || topk > 32 || C < 1 || C > TC_CH) return false;
    if ((long long)B * J >= (1ll << 31) || (long long)B * K >= (1ll << 31)) return false;
    if ((double)B * J * K < 4.0e6) return false;      // tiny problems: the materialising path is a handful of launches
    return topk_granule(K, topk) != 0 && tc_encode_fn() != nullptr;
}

size_t match_tc_topk_workspace_bytes(int B, int C, int J, int K, int topk) {
    return make_plan(B, C, J, K, topk).total + ws_block((size_t)B * J * 4) + ws_block((size_t)B * K * 4);
}

// P: fs, fr, B, C, J, K, beta, alpha, lse (given, [B,J]); out_idx / out_w [B,J,topk]
int launch_match_tc_topk(const MatchParams &P0, int topk, int64_t *out_idx, float *out_w, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!match_tc_topk_supported(P0.B, P0.C, P0.J, P0.K, topk)) return DSIR_ERR_UNSUPPORTED;
    const TcPlan pl = make_plan(P0.B, P0.C, P0.J, P0.K, topk);
    char *base = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    const size_t need = pl.total - 1024 + ws_block((size_t)P0.B * P0.J * 4) + ws_block((size_t)P0.B * P0.K * 4);
    if (ws == nullptr || (size_t)(base - (char *)ws) + need > ws_bytes) return DSIR_ERR_WORKSPACE;
    MatchParams P = P0;
    float *ns = (float *)(base + pl.total - 1024);
    float *nr = (float *)((char *)ns + ws_block((size_t)P.B * P.J * 4));
    P.ns = ns; P.nr = nr; P.reuse_prep = 0; P.prior_idx = nullptr;
    TcReady Rdy;
    int rc;
    if ((rc = tc_prepare(P, pl, base, Rdy, st))) return rc;
    TcParams &T = Rdy.T;
    T.prime_div = 0;
    T.umin = (float *)(base + pl.off_umin); T.gran32 = pl.gran == 32 ? 1 : 0; T.G = pl.G;
    float *thr = (float *)(base + pl.off_thr);
    int *ex = (int *)(base + pl.off_exrows);
    T.thr_in = thr; T.tk_cnt = (int *)(base + pl.off_tkcnt); T.tk_list = (int *)(base + pl.off_tklist); T.tk_cap = pl.cap;
    if ((rc = tc_launch_filter<1>(pl, Rdy, st))) return rc;
    const long long nrows = (long long)P.B * P.J;
    const unsigned wgrid = (unsigned)((nrows + 7) / 8);
    topk_thr_kernel<<<wgrid, 256, 0, st>>>(T, topk, P.beta, P.alpha, thr, T.tk_cnt, ex);
    DSIR_LAUNCH_CHECK();
    if ((rc = tc_launch_filter<2>(pl, Rdy, st))) return rc;
    RefineParams R{};
    R.B = P.B; R.J = P.J; R.K = P.K; R.C = P.C; R.S = pl.S; R.Jpad = pl.Jpad;
    R.fs = P.fs; R.fr = P.fr; R.ns = P.ns; R.nr = P.nr;
    const int C4 = (P.C + 3) / 4 * 4;
    float *frt = (float *)(base + pl.off_frt);
    feat_point_major_kernel<<<dim3(cdiv(P.K, 32), cdiv(C4, 32), P.B), 256, 0, st>>>(P.fr, P.C, P.K, C4, frt);
    DSIR_LAUNCH_CHECK();
    if (pl.cap <= 32) topk_listed_kernel<1><<<wgrid, 256, 0, st>>>(R, T.tk_cnt, T.tk_list, pl.cap, topk, frt, C4, P.beta, P.alpha, P.lse, out_idx, out_w, ex);
    else if (pl.cap <= 64) topk_listed_kernel<2><<<wgrid, 256, 0, st>>>(R, T.tk_cnt, T.tk_list, pl.cap, topk, frt, C4, P.beta, P.alpha, P.lse, out_idx, out_w, ex);
    else topk_listed_kernel<4><<<wgrid, 256, 0, st>>>(R, T.tk_cnt, T.tk_list, pl.cap, topk, frt, C4, P.beta, P.alpha, P.lse, out_idx, out_w, ex);
    DSIR_LAUNCH_CHECK();
    if (topk <= 8) topk_exact_kernel<8><<<Rdy.sms * 2, 256, 0, st>>>(R, topk, frt, C4, P.beta, P.alpha, P.lse, out_idx, out_w, ex);
    else topk_exact_kernel<32><<<Rdy.sms, 256, 0, st>>>(R, topk, frt, C4, P.beta, P.alpha, P.lse, out_idx, out_w, ex);
    DSIR_LAUNCH_CHECK();
    return DSIR_OK;
}

// diagnostic: rows of the last top-k launch on this workspace that took the exhaustive pass (synchronises the stream)
int match_tc_topk_exhaustive_rows(const void *ws, int B, int C, int J, int K, int topk, int *out, cudaStream_t st) {
    const TcPlan pl = make_plan(B, C, J, K, topk);
    if (!pl.gran) return DSIR_ERR_UNSUPPORTED;
    const char *base = (const char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    DSIR_CUDA_TRY(cudaMemcpyAsync(out, base + pl.off_exrows, 4, cudaMemcpyDeviceToHost, st));
    DSIR_CUDA_TRY(cudaStreamSynchronize(st));
    return DSIR_OK;
}

// diagnostic: number of rows the last launch sent to the exhaustive rescue path (synchronises the stream)
int match_tc_rescued_rows(const void *ws, int B, int C, int J, int K, int *out, cudaStream_t st) {
    const TcPlan pl = make_plan(B, C, J, K);
    const char *base = (const char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    DSIR_CUDA_TRY(cudaMemcpyAsync(out, base + pl.off_count, 4, cudaMemcpyDeviceToHost, st));
    DSIR_CUDA_TRY(cudaStreamSynchronize(st));
    return DSIR_OK;
}

// diagnostic: device-side timing of the LAST filter launch on this workspace (synchronises the stream):
// out[0] = kernel span in ns (last CTA end - first CTA start, %globaltimer), out[1] = mean SM cycles per CTA,
// out[2] = mean cycles per 256x128 unit (the tensor-pipe floor is 2 x 5 x 64 = 640)
int match_tc_filter_timing(const void *ws, int B, int C, int J, int K, double *out, cudaStream_t st) {
    const TcPlan pl = make_plan(B, C, J, K);
    const char *base = (const char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    static thread_local unsigned long long h[256 * 4];
    DSIR_CUDA_TRY(cudaMemcpyAsync(h, base + pl.off_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    DSIR_CUDA_TRY(cudaStreamSynchronize(st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = B * pl.RB * pl.S;
    const int grid = items < sms ? items : (sms > 256 ? 256 : sms);
    unsigned long long t0 = ~0ull, t1 = 0, cyc = 0, units = 0;
    for (int i = 0; i < grid; ++i) {
        t0 = h[4 * i] < t0 ? h[4 * i] : t0;
        t1 = h[4 * i + 1] > t1 ? h[4 * i + 1] : t1;
        cyc += h[4 * i + 2];
        units += h[4 * i + 3];
    }
    out[0] = (double)(t1 - t0);
    out[1] = (double)cyc / grid;
    out[2] = units ? (double)cyc / (double)units : 0.0;
    return DSIR_OK;
}


// diagnostic (DSIR_TC_DEBUG=2): clock stamps of block 0's first 256 units.  out[4096] u32:
//   [ (useq*2+r)*2 + {0,1} ]   MMA issuer of row block r: accumulator free seen / MMAs of the tile issued+committed
//   [ 1024 + useq*8 + k ]      epilogue warp 0: k=0 full seen, 1 first 32 columns in registers, 2 first step done,
//                              3 second 32 columns in registers, 4 second step done, 5 accumulator released
int match_tc_filter_trace(const void *ws, int B, int C, int J, int K, unsigned int *out, cudaStream_t st) {
    if (!TC_TRACE) return DSIR_ERR_UNSUPPORTED;   // only in -DDSIR_TC_TRACE builds
    const TcPlan pl = make_plan(B, C, J, K);
    const char *base = (const char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    DSIR_CUDA_TRY(cudaMemcpyAsync(out, base + pl.off_trace, 4096 * 4, cudaMemcpyDeviceToHost, st));
    DSIR_CUDA_TRY(cudaStreamSynchronize(st));
    return DSIR_OK;
}

}  // namespace dsir
