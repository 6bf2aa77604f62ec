// Fused feature-distance + row-argmin on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces the chunked block network/model.py:558-569 of the reference (match_features_V2 + .min(dim=2)[1]),
// the [J,K] score matrix never leaves TMEM/registers.  Bit-exactness on the indices is kept by a
// FILTER-AND-REFINE scheme:
//
//   prep    K-major (point-major) fp32 copies of both feature sets; the reference copy is scaled by -2
//           (exact), padded reference norms, per-batch max reference norm.
//   filter  persistent warp-specialised kernel: TMA (SWIZZLE_128B boxes) -> tcgen05.mma kind::tf32,
//           M=128 x N=128 accumulators in TMEM (2 row blocks x 2 stages = 512 columns) -> epilogue warps
//           read them back with tcgen05.ld and keep, per source row, the T smallest approximate values
//           x_jk = nr_k - 2<s_j,r_k>_tf32 that ever came within `margin_j` of the running minimum.
//           margin_j = 2 * eps_j where eps_j bounds the tf32 input-truncation error of x_jk, so the exact fp32
//           argmin is always among the kept candidates unless the list saturated.
//   refine  per row: candidates within margin of the approximate minimum are re-scored in exact fp32 with the
//           op order of match_fp32.cu (fma chain over channels, ((-2 dot)+ns)+nr) and the (value, index)
//           lexicographic minimum is taken -> identical indices AND minima to the fp32 kernel.
//   rescue  rows whose list saturated (or held no finite candidate) are recomputed exhaustively in fp32.
#include <cuda.h>

#include "match_tc.cuh"

namespace dsir {

namespace {

constexpr int TC_T = 4;             // candidates kept per row per split
constexpr int TC_BM = 256;          // source rows per work item (two M=128 accumulators)
constexpr int TC_BN = 128;          // reference rows per unit (one N=128 MMA)
constexpr int TC_STAGES = 4;        // B ring depth
constexpr int TC_THREADS = 384;     // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-11 epilogue
constexpr int TC_EPI_WARPS = 8;     // warps 4-7 own row block 0, warps 8-11 row block 1; TMEM lane quadrant = warp % 4
constexpr int TC_MAX_SPLIT = 8;
constexpr float TC_PAD_NORM = 3.0e38f;
constexpr uint32_t TILE_BYTES = 128 * 128;  // one TMA box: 128 rows x 32 floats (128 B, swizzled)

// ---------------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no libcuda link dependency)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [B][N][Cp] fp32, box = 32 channels x 128 rows, 128-byte swizzle; rows/batches beyond the extent read as zero
bool make_feat_tmap(CUtensorMap *m, const float *base, int B, int N, int Cp) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)Cp * 4, (cuuint64_t)N * Cp * 4};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// the wait names the destination registers so that no consumer can be scheduled above it
__device__ __forceinline__ void tmem_wait8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ float tmem_ld1_sync(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r) : "r"(taddr) : "memory");
    return __uint_as_float(r);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart (SBO), LBO unused (=1)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, N=128, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct PipeState {
    int stage;
    uint32_t phase;
    __device__ __forceinline__ void advance(int n) {
        if (++stage == n) { stage = 0; phase ^= 1u; }
    }
};

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

// 32 accumulator columns of one row: x = acc + nr.  Fast path: 32 independent adds, a min tree, ONE vote.  Whenever
// some lane of the warp sees a value within `margin` of its running minimum, the (rare) slow path re-reads exactly
// the flagged columns from TMEM one at a time, so the hot loop stays small enough for the instruction cache and
// carries no per-element branches.
__device__ __forceinline__ void filter32(const uint32_t (&v)[32], const float *nr32, int col0, uint32_t taddr, float margin,
                                         float &thr, float (&cv)[TC_T], int (&ci)[TC_T]) {
    float x[32];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float4 n4 = *reinterpret_cast<const float4 *>(nr32 + 4 * e);
        x[4 * e + 0] = __fadd_rn(__uint_as_float(v[4 * e + 0]), n4.x);
        x[4 * e + 1] = __fadd_rn(__uint_as_float(v[4 * e + 1]), n4.y);
        x[4 * e + 2] = __fadd_rn(__uint_as_float(v[4 * e + 2]), n4.z);
        x[4 * e + 3] = __fadd_rn(__uint_as_float(v[4 * e + 3]), n4.w);
    }
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = fminf(fminf(x[4 * e], x[4 * e + 1]), fminf(x[4 * e + 2], x[4 * e + 3]));
    const float mm = fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])), fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
    if (__any_sync(0xffffffffu, mm < thr)) {
        unsigned mask = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) mask |= (x[e] < thr) ? (1u << e) : 0u;
        unsigned um = __reduce_or_sync(0xffffffffu, mask);
#pragma unroll 1
        while (um) {
            const int e = __ffs(um) - 1;
            um &= um - 1;
            const float xe = __fadd_rn(tmem_ld1_sync(taddr + e), nr32[e]);   // bit-identical to x[e]
            if (xe < thr) {
                cand_insert<TC_T>(cv, ci, xe, col0 + e);
                thr = cv[0] + margin;
            }
        }
    }
}

struct TcParams {
    int B, J, K, C, Cp;
    int RB, U, S;           // row blocks (256), units (128), k-splits
    int Jpad, Kpad;
    const float *ns;        // [B,J] exact squared norms
    const float *nr_pad;    // [B,Kpad] exact squared norms, TC_PAD_NORM beyond K
    const float *rmax;      // [B] max reference squared norm
    float *cand_val;        // [B][Jpad][S][T]
    int *cand_idx;
};

template <int KB>  // 32-channel blocks (Cp = 32*KB)
__global__ __launch_bounds__(TC_THREADS, 1) void match_tc_filter_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                        const __grid_constant__ CUtensorMap mapB,
                                                                        TcParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                            // [2][KB][16 KB]
    uint8_t *sB = sA + 2 * KB * TILE_BYTES;                        // [STAGES][KB][16 KB]
    float *sNr = (float *)(sB + TC_STAGES * KB * TILE_BYTES);      // [STAGES][128]
    uint64_t *bars = (uint64_t *)(sNr + TC_STAGES * TC_BN);
    uint64_t *full_b = bars, *empty_b = bars + TC_STAGES;
    uint64_t *tmem_full = bars + 2 * TC_STAGES, *tmem_empty = tmem_full + 2;
    uint64_t *full_a = tmem_empty + 2, *empty_a = full_a + 1;
    uint32_t *tmem_slot = (uint32_t *)(empty_a + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_items = P.B * P.RB * P.S;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1 + TC_EPI_WARPS); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TC_EPI_WARPS * 32); }
        mbar_init(full_a, 1);
        mbar_init(empty_a, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            PipeState pb{0, 0};
            uint32_t iphase = 0;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
                const int sp = it % P.S, rb = (it / P.S) % P.RB, b = it / (P.S * P.RB);
                const int u0 = (int)((long long)sp * P.U / P.S), u1 = (int)((long long)(sp + 1) * P.U / P.S);
                mbar_wait(empty_a, iphase ^ 1u);
                mbar_expect_tx(full_a, 2 * KB * TILE_BYTES);
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_3d(sA + (a * KB + kb) * TILE_BYTES, &mapA, kb * 32, rb * TC_BM + a * 128, b, full_a);
                for (int u = u0; u < u1; ++u) {
                    while (!mbar_try_wait(&empty_b[pb.stage], pb.phase ^ 1u)) __nanosleep(64);
                    mbar_expect_tx(&full_b[pb.stage], KB * TILE_BYTES + TC_BN * 4);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_3d(sB + (pb.stage * KB + kb) * TILE_BYTES, &mapB, kb * 32, u * TC_BN, b, &full_b[pb.stage]);
                    bulk_g2s(sNr + pb.stage * TC_BN, P.nr_pad + (size_t)b * P.Kpad + (size_t)u * TC_BN, TC_BN * 4,
                             &full_b[pb.stage]);
                    pb.advance(TC_STAGES);
                }
                iphase ^= 1u;
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            PipeState pb{0, 0}, pa{0, 0};
            uint32_t iphase = 0;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
                const int sp = it % P.S;
                const int u0 = (int)((long long)sp * P.U / P.S), u1 = (int)((long long)(sp + 1) * P.U / P.S);
                mbar_wait(full_a, iphase);
                for (int u = u0; u < u1; ++u) {
                    mbar_wait(&full_b[pb.stage], pb.phase);
                    mbar_wait(&tmem_empty[pa.stage], pa.phase ^ 1u);
                    tc_fence_after();
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(pa.stage * 256 + a * 128);
#pragma unroll
                        for (int ks = 0; ks < KB * 4; ++ks) {
                            const int kb = ks >> 2, kk = ks & 3;
                            uint64_t da = make_kmajor_sw128_desc(smem_u32(sA + (a * KB + kb) * TILE_BYTES) + kk * 32);
                            uint64_t db = make_kmajor_sw128_desc(smem_u32(sB + (pb.stage * KB + kb) * TILE_BYTES) + kk * 32);
                            mma_tf32(d_tmem, da, db, TC_IDESC, ks > 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty_b[pb.stage]);     // B stage reusable once these MMAs retire (and the epilogue is done with nr)
                    tc_commit(&tmem_full[pa.stage]);   // accumulators ready for the epilogue
                    pb.advance(TC_STAGES);
                    pa.advance(2);
                }
                tc_commit(empty_a);
                iphase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue: TMEM -> registers -> candidate lists ===========================
        const int q = warp & 3;                       // TMEM lane quadrant of this warp
        const int a = (warp - 4) >> 2;                // row block (accumulator half) of this warp
        const int trow = q * 32 + lane;               // row inside the 128-row block
        PipeState pb{0, 0}, pa{0, 0};
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const int sp = it % P.S, rb = (it / P.S) % P.RB, b = it / (P.S * P.RB);
            const int u0 = (int)((long long)sp * P.U / P.S), u1 = (int)((long long)(sp + 1) * P.U / P.S);
            float cv[TC_T];
            int ci[TC_T];
#pragma unroll
            for (int t = 0; t < TC_T; ++t) { cv[t] = INFINITY; ci[t] = -1; }
            float thr = INFINITY;
            const int j = rb * TC_BM + a * 128 + trow;
            const float nsj = j < P.J ? P.ns[(size_t)b * P.J + j] : 0.f;
            // 2 * eps, eps = 2 * (2^-9 + 2^-20) |s||r|  (tf32 truncation of both operands of -2<s,r>), +5 %
            const float margin = 8.2e-3f * sqrtf(nsj) * sqrtf(P.rmax[b]) + 1e-30f;
            for (int u = u0; u < u1; ++u) {
                mbar_wait(&tmem_full[pa.stage], pa.phase);
                tc_fence_after();
                const float *nr = sNr + pb.stage * TC_BN;
                const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pa.stage * 256 + a * 128);
                const int col0 = u * TC_BN;
                uint32_t va[32], vb[32];
                tmem_ld32(tbase, va);
#pragma unroll 1
                for (int g = 0; g < TC_BN / 32; g += 2) {     // ping-pong: the next 32 columns fly during this step's math
                    tmem_wait32(va);
                    tmem_ld32(tbase + (g + 1) * 32, vb);
                    filter32(va, nr + g * 32, col0 + g * 32, tbase + g * 32, margin, thr, cv, ci);
                    tmem_wait32(vb);
                    if (g + 2 < TC_BN / 32) tmem_ld32(tbase + (g + 2) * 32, va);
                    filter32(vb, nr + (g + 1) * 32, col0 + (g + 1) * 32, tbase + (g + 1) * 32, margin, thr, cv, ci);
                }
                tc_fence_before();
                mbar_arrive(&tmem_empty[pa.stage]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_b[pb.stage]);
                pb.advance(TC_STAGES);
                pa.advance(2);
            }
            const size_t row = (size_t)b * P.Jpad + (size_t)j - (size_t)0;
            float *ov = P.cand_val + (((size_t)b * P.Jpad + (size_t)rb * TC_BM + a * 128 + trow) * P.S + sp) * TC_T;
            int *oi = P.cand_idx + (((size_t)b * P.Jpad + (size_t)rb * TC_BM + a * 128 + trow) * P.S + sp) * TC_T;
            (void)row;
#pragma unroll
            for (int t = 0; t < TC_T; ++t) { ov[t] = cv[t]; oi[t] = ci[t]; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------
// prep: [B,C,N] (any strides) -> K-major copy [B][N][Cp] * scale, channels C..Cp zero
// ---------------------------------------------------------------------------------------------------------
__global__ void tc_transpose_kernel(dsir_feat f, int C, int N, int Cp, float scale, float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *src = f.ptr + (size_t)b * f.batch_stride;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // i: channel, threadIdx.x: point
        int c = c0 + i, n = n0 + threadIdx.x;
        float v = 0.f;
        if (c < C && n < N) v = src[(size_t)c * f.chan_stride + (size_t)n * f.point_stride];
        tile[i][threadIdx.x] = v * scale;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // i: point, threadIdx.x: channel
        int n = n0 + i, c = c0 + threadIdx.x;
        if (n < N && c < Cp) out[((size_t)b * N + n) * Cp + c] = tile[threadIdx.x][i];
    }
}

__global__ void tc_pad_norms_kernel(const float *__restrict__ nr, int K, int Kpad, float *__restrict__ nr_pad,
                                    int *__restrict__ rmax_bits) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float v = TC_PAD_NORM;
    float m = 0.f;
    if (k < K) { v = nr[(size_t)b * K + k]; m = v; }
    if (k < Kpad) nr_pad[(size_t)b * Kpad + k] = v;
    m = warp_max(m);  // NaN-free maxima only; a NaN norm sends its rows to the rescue path via NaN candidates
    if ((threadIdx.x & 31) == 0) atomicMax(&rmax_bits[b], __float_as_int(m));
}

// ---------------------------------------------------------------------------------------------------------
// refine: one warp per source row, one lane per candidate
// ---------------------------------------------------------------------------------------------------------
struct RefineParams {
    int B, J, K, C, Cp, S, Jpad;
    const float *a_copy;  // [B][J][Cp]
    const float *b_copy;  // [B][K][Cp]  (= -2 r)
    const float *ns, *nr, *rmax;
    const float *cand_val;
    const int *cand_idx;
    int64_t *idx;
    float *min_d;
    int *rescue_count;
    int *rescue_rows;  // [B*J] flat row ids
    unsigned long long *rescue_keys;  // [B*J] (ordered distance bits << 32) | index, atomicMin target
};

__device__ __forceinline__ float exact_dist(const float *__restrict__ srow, const float *__restrict__ brow, int C, float nsj, float nrk) {
    float dot = 0.f;
    for (int c = 0; c < C; c += 4) {  // Cp is a multiple of 32 and channels >= C are zero: fma(0,0,dot) == dot
        float4 s4 = *reinterpret_cast<const float4 *>(srow + c);
        float4 b4 = *reinterpret_cast<const float4 *>(brow + c);
        dot = __fmaf_rn(s4.x, __fmul_rn(-0.5f, b4.x), dot);
        dot = __fmaf_rn(s4.y, __fmul_rn(-0.5f, b4.y), dot);
        dot = __fmaf_rn(s4.z, __fmul_rn(-0.5f, b4.z), dot);
        dot = __fmaf_rn(s4.w, __fmul_rn(-0.5f, b4.w), dot);
    }
    return l2_from_dot(dot, nsj, nrk);
}

__global__ void match_tc_refine_kernel(RefineParams P) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)P.B * P.J) return;
    const int b = (int)(row / P.J), j = (int)(row % P.J);
    const int ncand = P.S * TC_T;
    const size_t cbase = ((size_t)b * P.Jpad + j) * ncand;
    float v = INFINITY;
    int k = -1;
    if (lane < ncand) { v = P.cand_val[cbase + lane]; k = P.cand_idx[cbase + lane]; }
    const bool valid = k >= 0 && k < P.K && v < 1e38f;
    const float gmin = warp_min(valid ? v : INFINITY);
    const float nsj = P.ns[(size_t)b * P.J + j];
    const float margin = 8.2e-3f * sqrtf(nsj) * sqrtf(P.rmax[b]) + 1e-30f;
    const bool take = valid && v <= gmin + margin;
    // saturation: the last slot of some split is still within the margin -> something may have been dropped
    const bool sat = valid && ((lane % TC_T) == TC_T - 1) && take;
    const unsigned any_take = __ballot_sync(0xffffffffu, take);
    const unsigned any_sat = __ballot_sync(0xffffffffu, sat);
    float d = INFINITY;
    int kk = 0x7fffffff;
    if (take) {
        d = exact_dist(P.a_copy + ((size_t)b * P.J + j) * P.Cp, P.b_copy + ((size_t)b * P.K + k) * P.Cp, P.Cp, nsj,
                       P.nr[(size_t)b * P.K + k]);
        kk = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float d2 = __shfl_xor_sync(0xffffffffu, d, o);
        int k2 = __shfl_xor_sync(0xffffffffu, kk, o);
        if (d2 < d || (d2 == d && k2 < kk)) { d = d2; kk = k2; }
    }
    const bool rescue = (any_take == 0u) || (any_sat != 0u) || !(d < INFINITY);
    if (lane == 0) {
        P.idx[row] = rescue ? 0 : (int64_t)kk;
        if (P.min_d) P.min_d[row] = d;
        if (rescue) {
            int pos = atomicAdd(P.rescue_count, 1);
            P.rescue_rows[pos] = (int)row;
            P.rescue_keys[pos] = ~0ull;
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
constexpr int RESCUE_CHUNK = 2048;

__global__ __launch_bounds__(256) void match_tc_rescue_kernel(RefineParams P, unsigned long long *keys) {
    __shared__ float srow[128];
    __shared__ unsigned long long red[8];
    const int count = *P.rescue_count;
    const int nchunk = (P.K + RESCUE_CHUNK - 1) / RESCUE_CHUNK;
    const long long units = (long long)count * nchunk;
    for (long long uidx = blockIdx.x; uidx < units; uidx += gridDim.x) {
        const int i = (int)(uidx / nchunk), ch = (int)(uidx % nchunk);
        const int row = P.rescue_rows[i];
        const int b = row / P.J, j = row % P.J;
        __syncthreads();
        for (int c = threadIdx.x; c < P.Cp; c += blockDim.x) srow[c] = P.a_copy[((size_t)b * P.J + j) * P.Cp + c];
        __syncthreads();
        const float nsj = P.ns[(size_t)b * P.J + j];
        unsigned long long best = ~0ull;
        const int kend = min(P.K, (ch + 1) * RESCUE_CHUNK);
        for (int k = ch * RESCUE_CHUNK + threadIdx.x; k < kend; k += blockDim.x) {
            float d = exact_dist(srow, P.b_copy + ((size_t)b * P.K + k) * P.Cp, P.Cp, nsj, P.nr[(size_t)b * P.K + k]);
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
    int Cp, KB, RB, U, S, Jpad, Kpad;
    size_t off_a, off_b, off_nrpad, off_rmax, off_cval, off_cidx, off_count, off_rows, off_keys, total;
};

TcPlan make_plan(int B, int C, int J, int K) {
    TcPlan p;
    p.Cp = (C + 31) / 32 * 32;
    p.KB = p.Cp / 32;
    p.RB = (J + TC_BM - 1) / TC_BM;
    p.U = (K + TC_BN - 1) / TC_BN;
    p.Jpad = p.RB * TC_BM;
    p.Kpad = p.U * TC_BN;
    long long items = (long long)B * p.RB;
    int S = 1;
    if (items < 2 * 148) {
        S = (int)((2 * 148 + items - 1) / items);
        if (S > TC_MAX_SPLIT) S = TC_MAX_SPLIT;
        if (S > p.U) S = p.U;
        if (S < 1) S = 1;
    }
    p.S = S;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += ws_block(bytes); return o; };
    p.off_a = take((size_t)B * J * p.Cp * 4);
    p.off_b = take((size_t)B * K * p.Cp * 4);
    p.off_nrpad = take((size_t)B * p.Kpad * 4);
    p.off_rmax = take((size_t)B * 4);
    p.off_cval = take((size_t)B * p.Jpad * S * TC_T * 4);
    p.off_cidx = take((size_t)B * p.Jpad * S * TC_T * 4);
    p.off_count = take(256);
    p.off_rows = take((size_t)B * J * 4);
    p.off_keys = take((size_t)B * J * 8);
    p.total = off + 1024;
    return p;
}

size_t filter_smem_bytes(int KB) {
    return 1024 + (size_t)(2 + TC_STAGES) * KB * TILE_BYTES + TC_STAGES * TC_BN * 4 + 256;
}

}  // namespace

bool match_tc_supported(const dsir_feat &fs, const dsir_feat &fr, int B, int C, int J, int K) {
    (void)fs; (void)fr;
    if (C < 1 || C > 64) return false;
    if ((long long)B * J >= (1ll << 31) || (long long)B * K >= (1ll << 31)) return false;
    return get_encode_fn() != nullptr;
}

bool match_tc_profitable(int B, int C, int J, int K) {
    (void)C;
    return (double)B * J * K >= 4.0e6;  // below this the prep/refine launches dominate
}

size_t match_tc_workspace_bytes(int B, int C, int J, int K) {
    if (C > 64) return 0;
    return make_plan(B, C, J, K).total;
}

int launch_match_tc(const MatchParams &P, void *ws, size_t ws_bytes, cudaStream_t st) {
    const TcPlan pl = make_plan(P.B, P.C, P.J, P.K);
    char *base = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    if (ws == nullptr || (size_t)(base - (char *)ws) + pl.total - 1024 > ws_bytes) return DSIR_ERR_WORKSPACE;
    float *a_copy = (float *)(base + pl.off_a), *b_copy = (float *)(base + pl.off_b);
    float *nr_pad = (float *)(base + pl.off_nrpad), *rmax = (float *)(base + pl.off_rmax);
    float *cval = (float *)(base + pl.off_cval);
    int *cidx = (int *)(base + pl.off_cidx), *count = (int *)(base + pl.off_count), *rows = (int *)(base + pl.off_rows);

    DSIR_CUDA_TRY(cudaMemsetAsync(rmax, 0, (size_t)P.B * 4, st));
    DSIR_CUDA_TRY(cudaMemsetAsync(count, 0, 4, st));
    {
        dim3 blk(32, 8);
        dim3 ga(cdiv(P.J, 32), pl.Cp / 32, P.B), gb(cdiv(P.K, 32), pl.Cp / 32, P.B);
        tc_transpose_kernel<<<ga, blk, 0, st>>>(P.fs, P.C, P.J, pl.Cp, 1.0f, a_copy);
        DSIR_LAUNCH_CHECK();
        tc_transpose_kernel<<<gb, blk, 0, st>>>(P.fr, P.C, P.K, pl.Cp, -2.0f, b_copy);
        DSIR_LAUNCH_CHECK();
        dim3 gn(cdiv(pl.Kpad, 256), P.B);
        tc_pad_norms_kernel<<<gn, 256, 0, st>>>(P.nr, P.K, pl.Kpad, nr_pad, (int *)rmax);
        DSIR_LAUNCH_CHECK();
    }
    CUtensorMap mapA, mapB;
    if (!make_feat_tmap(&mapA, a_copy, P.B, P.J, pl.Cp) || !make_feat_tmap(&mapB, b_copy, P.B, P.K, pl.Cp))
        return DSIR_ERR_UNSUPPORTED;

    TcParams T{};
    T.B = P.B; T.J = P.J; T.K = P.K; T.C = P.C; T.Cp = pl.Cp; T.RB = pl.RB; T.U = pl.U; T.S = pl.S;
    T.Jpad = pl.Jpad; T.Kpad = pl.Kpad; T.ns = P.ns; T.nr_pad = nr_pad; T.rmax = rmax; T.cand_val = cval; T.cand_idx = cidx;
    const int items = P.B * pl.RB * pl.S;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = items < sms ? items : sms;
    const size_t smem = filter_smem_bytes(pl.KB);
    if (pl.KB == 1) {
        DSIR_CUDA_TRY(cudaFuncSetAttribute(match_tc_filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        match_tc_filter_kernel<1><<<grid, TC_THREADS, smem, st>>>(mapA, mapB, T);
    } else {
        DSIR_CUDA_TRY(cudaFuncSetAttribute(match_tc_filter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        match_tc_filter_kernel<2><<<grid, TC_THREADS, smem, st>>>(mapA, mapB, T);
    }
    DSIR_LAUNCH_CHECK();

    RefineParams R{};
    R.B = P.B; R.J = P.J; R.K = P.K; R.C = P.C; R.Cp = pl.Cp; R.S = pl.S; R.Jpad = pl.Jpad;
    R.a_copy = a_copy; R.b_copy = b_copy; R.ns = P.ns; R.nr = P.nr; R.rmax = rmax; R.cand_val = cval; R.cand_idx = cidx;
    R.idx = P.idx; R.min_d = P.min_d; R.rescue_count = count; R.rescue_rows = rows;
    R.rescue_keys = (unsigned long long *)(base + pl.off_keys);
    const long long nrows = (long long)P.B * P.J;
    match_tc_refine_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(R);
    DSIR_LAUNCH_CHECK();
    match_tc_rescue_kernel<<<sms * 4, 256, 0, st>>>(R, R.rescue_keys);
    DSIR_LAUNCH_CHECK();
    match_tc_rescue_finalize_kernel<<<sms, 256, 0, st>>>(R, R.rescue_keys);
    DSIR_LAUNCH_CHECK();
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

}  // namespace dsir
