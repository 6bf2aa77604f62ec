// Micro-benchmarks that size the tcgen05 match kernel (run on a B200 through gpurun):
//   1. tcgen05.ld throughput per SM with 4 / 8 / 16 reader warps (x32 loads of the warp's own lane quadrant)
//   2. tcgen05.mma kind::tf32 issue rate: SS M128xN128, SS M128xN256, A-in-TMEM M128xN128 (operands = zeros in smem)
//   3. both at once (readers + MMA) to see whether TMEM reads slow the tensor pipe
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_tc tools/ubench_tc.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <type_traits>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
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
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 : : "memory");
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t swz) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swz << 61;
    return d;
}

// mode bit 0: run MMA issuer; bit 1: run readers.  variant: 0 SS N128, 1 SS N256, 2 TS N128, 3 SS N128 K-ext (9 k-steps)
__global__ __launch_bounds__(640, 1) void ubench(int mode, int variant, int reader_warps, int iters, int reader_iters, int do_min,
                                                 long long *out, float *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) tmem_alloc(&slot, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    long long t0 = clock64();
    if (warp == 0) {
        if (lane == 0 && (mode & 1)) {
            auto issue = [&](auto NV, auto F16V, auto KSV, auto TSV, int special) {
                constexpr int N = decltype(NV)::value;
                constexpr bool f16 = decltype(F16V)::value;
                constexpr int ksteps = decltype(KSV)::value;
                constexpr bool ts = decltype(TSV)::value;
                constexpr uint32_t idesc = (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint64_t da0 = make_desc(smem_u32(smem), 1024, 2);
                const uint64_t db0 = make_desc(smem_u32(smem + 32 * 1024), 1024, 2);
#pragma unroll 1
                for (int it = 0; it < iters; ++it) {
                    const uint32_t d = tbase + (uint32_t)((it & 1) * 256);
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const int kb = (ks >> 2) & 1, kk = ks & 3;
                        const uint64_t db = db0 + (uint64_t)((kb * (N * 128) + kk * 32) >> 4);
                        const uint64_t da = da0 + (uint64_t)((kb * 16384 + kk * 32) >> 4);
                        if (special == 1 && ks == 4) {   // fifth k-step from a SWIZZLE_32B tile (SBO 256, layout 6)
                            mma_ss_f16(d, make_desc(smem_u32(smem + 80 * 1024), 256, 6), make_desc(smem_u32(smem + 84 * 1024), 256, 6), idesc, 1);
                        } else if (special == 2) {       // alternate between two accumulators MMA by MMA
                            mma_ss_f16(tbase + (uint32_t)((ks & 1) * 128), da, db, idesc, ks > 1);
                        } else if (ts) mma_ts(d, tbase + 384 + ks * 8, db, idesc, ks > 0);
                        else if (f16) mma_ss_f16(d, da, db, idesc, ks > 0);
                        else mma_ss(d, da, db, idesc, ks > 0);
                    }
                }
            };
            using std::integral_constant;
            switch (variant) {
                case 0: issue(integral_constant<int, 128>{}, integral_constant<bool, false>{}, integral_constant<int, 8>{}, integral_constant<bool, false>{}, 0); break;
                case 1: issue(integral_constant<int, 256>{}, integral_constant<bool, false>{}, integral_constant<int, 8>{}, integral_constant<bool, false>{}, 0); break;
                case 2: issue(integral_constant<int, 128>{}, integral_constant<bool, false>{}, integral_constant<int, 8>{}, integral_constant<bool, true>{}, 0); break;
                case 3: issue(integral_constant<int, 128>{}, integral_constant<bool, false>{}, integral_constant<int, 9>{}, integral_constant<bool, false>{}, 0); break;
                case 4: issue(integral_constant<int, 128>{}, integral_constant<bool, true>{}, integral_constant<int, 4>{}, integral_constant<bool, false>{}, 0); break;
                case 5: issue(integral_constant<int, 256>{}, integral_constant<bool, true>{}, integral_constant<int, 4>{}, integral_constant<bool, false>{}, 0); break;
                case 7: issue(integral_constant<int, 128>{}, integral_constant<bool, true>{}, integral_constant<int, 5>{}, integral_constant<bool, false>{}, 1); break;
                case 8: issue(integral_constant<int, 128>{}, integral_constant<bool, true>{}, integral_constant<int, 5>{}, integral_constant<bool, false>{}, 2); break;
                case 9: issue(integral_constant<int, 64>{}, integral_constant<bool, true>{}, integral_constant<int, 5>{}, integral_constant<bool, false>{}, 0); break;
                case 10: issue(integral_constant<int, 32>{}, integral_constant<bool, true>{}, integral_constant<int, 5>{}, integral_constant<bool, false>{}, 0); break;
                default: issue(integral_constant<int, 128>{}, integral_constant<bool, true>{}, integral_constant<int, 5>{}, integral_constant<bool, false>{}, 0); break;
            }
            tc_commit(&bar);
            while (!mbar_try_wait(&bar, 0)) {}
        }
    } else if (warp >= 4 && warp < 4 + reader_warps && (mode & 2)) {
        const int q = warp & 3;
        const int part = (warp - 4) >> 2;  // which 128-column slab
        const uint32_t ta = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)((part & 3) * 128);
        uint32_t va[32], vb[32];
        float acc = 3e38f, acc4[4] = {3e38f, 3e38f, 3e38f, 3e38f};
        tmem_ld32(ta, va);
        for (int it = 0; it < reader_iters; ++it) {
#pragma unroll
            for (int g = 0; g < 4; g += 2) {
                tmem_wait32(va);
                tmem_ld32(ta + (g + 1) * 32, vb);
                if (do_min) {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) acc4[(e >> 1) & 3] = min3(acc4[(e >> 1) & 3], __uint_as_float(va[e]), __uint_as_float(va[e + 1]));
                } else acc = fminf(acc, __uint_as_float(va[0]) + __uint_as_float(va[31]));
                tmem_wait32(vb);
                tmem_ld32(ta + ((g + 2) & 3) * 32, va);
                if (do_min) {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) acc4[(e >> 1) & 3] = min3(acc4[(e >> 1) & 3], __uint_as_float(vb[e]), __uint_as_float(vb[e + 1]));
                } else acc = fminf(acc, __uint_as_float(vb[0]) + __uint_as_float(vb[31]));
            }
        }
        tmem_wait32(va);
        acc = fminf(fminf(acc, acc4[0]), fminf(fminf(acc4[1], acc4[2]), acc4[3]));
        if (acc == 12345.f) sink[threadIdx.x] = acc;
    }
    long long t1 = clock64();
    // per-role elapsed cycles: warp 0 lane 0 = MMA, warp 4 lane 0 = readers
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4)) out[warp == 0 ? 0 : 1] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); tmem_dealloc(tbase, 512); }
}

int main() {
    long long *out;
    float *sink;
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&sink, 4096));
    const size_t smem = 1024 + 96 * 1024 + 1024;
    CK(cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto run = [&](const char *name, int mode, int variant, int rw, int iters, int riters, int do_min, int grid) -> int {
        long long h[2] = {0, 0};
        CK(cudaMemset(out, 0, 16));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        ubench<<<grid, 640, smem>>>(mode, variant, rw, iters, riters, do_min, out, sink);  // warm
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        ubench<<<grid, 640, smem>>>(mode, variant, rw, iters, riters, do_min, out, sink);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
        printf("%-44s grid=%3d  %.3f ms", name, grid, ms);
        if (mode & 1) {
            const int ksteps = variant == 3 ? 9 : (variant >= 6 ? 5 : (variant >= 4 ? 4 : 8));
            printf("  mma: %lld clk, %.1f clk per MMA", h[0], (double)h[0] / ((double)iters * ksteps));
        }
        if (mode & 2) {
            double bytes = (double)riters * 4 * 32 * 32 * 4 * rw;
            printf("  rd: %lld clk, %.1f B/clk/SM (%d warps)", h[1], bytes / (double)h[1], rw);
        }
        printf("\n");
        return 0;
    };
    for (int grid : {148}) {
        run("ld only, 4 warps", 2, 0, 4, 0, 2000, 0, grid);
        run("ld only, 8 warps", 2, 0, 8, 0, 2000, 0, grid);
        run("ld only, 16 warps", 2, 0, 16, 0, 2000, 0, grid);
        run("ld+min3, 4 warps", 2, 0, 4, 0, 2000, 1, grid);
        run("ld+min3, 8 warps", 2, 0, 8, 0, 2000, 1, grid);
        run("ld+min3, 16 warps", 2, 0, 16, 0, 2000, 1, grid);
        run("mma SS M128 N128 K64", 1, 0, 0, 4000, 0, 0, grid);
        run("mma SS M128 N256 K64", 1, 1, 0, 2000, 0, 0, grid);
        run("mma TS M128 N128 K64", 1, 2, 0, 4000, 0, 0, grid);
        run("mma SS M128 N128 K72", 1, 3, 0, 4000, 0, 0, grid);
        run("mma f16 SS M128 N128 K64", 1, 4, 0, 8000, 0, 0, grid);
        run("mma f16 SS M128 N256 K64", 1, 5, 0, 4000, 0, 0, grid);
        run("mma f16 SS M128 N128 K80", 1, 6, 0, 8000, 0, 0, grid);
        run("mma f16 K80, 5th step SW32 tile", 1, 7, 0, 8000, 0, 0, grid);
        run("mma f16 K80, alternating accumulators", 1, 8, 0, 8000, 0, 0, grid);
        run("mma f16 SS M128 N64 K80", 1, 9, 0, 8000, 0, 0, grid);
        run("mma f16 SS M128 N32 K80", 1, 10, 0, 8000, 0, 0, grid);
        run("mma f16 N64 K80 + ld+min3 16 warps", 3, 9, 16, 8000, 350, 1, grid);
        run("mma f16 N128 K80 + ld+min3 16 warps", 3, 6, 16, 8000, 700, 1, grid);
        run("mma f16 N128 K80 + ld+min3 8 warps", 3, 6, 8, 8000, 1400, 1, grid);
        run("mma f16 N256 K64 + ld+min3 16 warps", 3, 5, 16, 4000, 700, 1, grid);
        run("mma SS N128 + ld+min3 8 warps", 3, 0, 8, 4000, 2000, 1, grid);
        run("mma SS N128 + ld+min3 16 warps", 3, 0, 16, 4000, 1000, 1, grid);
        run("mma TS N128 + ld+min3 8 warps", 3, 2, 8, 4000, 2000, 1, grid);
        run("mma SS N256 + ld+min3 8 warps", 3, 1, 8, 2000, 2000, 1, grid);
    }
    return 0;
}
