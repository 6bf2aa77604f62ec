// Micro-benchmark: handoff latencies of the tcgen05.commit -> mbarrier -> epilogue -> mbarrier -> issuer ping-pong.
//   one issuer warp: { 5 MMAs (M128 N128 K16 f16); commit(full); wait(empty) }      (single accumulator, no overlap)
//   W consumer warps: { wait(full); [tcgen05.ld x32 + wait]; fence; arrive(empty) }
// period - 5*64 = commit->consumer wake + consumer->issuer wake (+ the optional TMEM read)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t a) { uint64_t d = 0; d |= (uint64_t)((a >> 4) & 0x3FFF); d |= 1ull << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= 1ull << 46; d |= 2ull << 61; return d; }

__global__ __launch_bounds__(640, 1) void ping(int iters, int consumers, int do_ld, int nmma, long long *out, float *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full, bar_empty;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((float *)smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_empty, consumers); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 18) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    long long t0 = clock64();
    if (warp == 17) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 32768));
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&bar_empty, ph ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int k = 0; k < nmma; ++k) mma_f16(tbase, da + (uint64_t)((k & 3) * 2), db + (uint64_t)((k & 3) * 2), idesc, k > 0);
            tc_commit(&bar_full);
            ph ^= 1u;
        }
        if (lane == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
    } else if (warp < consumers) {
        uint32_t ph = 0;
        float acc = 0.f;
        const uint32_t ta = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&bar_full, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (do_ld) { uint32_t v[32]; tmem_ld32(ta, v); acc += __uint_as_float(v[0]) + __uint_as_float(v[31]); }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty);
            ph ^= 1u;
        }
        if (acc == 1234.5f) sink[threadIdx.x] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 18) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory"); }
}

int main() {
    long long *out; float *sink;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 4096));
    const size_t smem = 1024 + 64 * 1024;
    CK(cudaFuncSetAttribute(ping, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4000;
    for (int nmma : {0, 1, 5}) for (int consumers : {1, 4, 8}) for (int do_ld : {0, 1}) {
        ping<<<148, 640, smem>>>(iters, consumers, do_ld, nmma, out, sink);
        CK(cudaDeviceSynchronize());
        long long h = 0; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        printf("nmma=%d consumers=%d ld=%d: %.0f clk per round trip (MMA floor %d)\n", nmma, consumers, do_ld, (double)h / iters, nmma * 64);
    }
    return 0;
}
