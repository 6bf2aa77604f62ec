// Shared helpers for the deepsir_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/deepsir_b200.h"

namespace dsir {

void set_last_cuda_error(cudaError_t e);
void count_launch();  // every kernel this library enqueues is counted (dsir_launch_count)

#define DSIR_CUDA_TRY(expr)                               \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) {                          \
            ::dsir::set_last_cuda_error(_e);              \
            return DSIR_ERR_CUDA;                         \
        }                                                 \
    } while (0)

// In-situ profiling (diagnostic, off by default): when enabled through dsir_profile_begin(), every launch site records
// a CUDA event on its stream right after the launch; dsir_profile_report() turns consecutive events into per-site times.
extern bool g_profile_on;
void profile_mark(cudaStream_t st, const char *file, int line);

// every launch site has the stream in a local named `st`
#define DSIR_LAUNCH_CHECK()                                                  \
    do {                                                                     \
        ::dsir::count_launch();                                              \
        DSIR_CUDA_TRY(cudaGetLastError());                                   \
        if (::dsir::g_profile_on) ::dsir::profile_mark(st, __FILE__, __LINE__); \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over the caller's workspace; every block is 256-byte aligned.
struct Workspace {
    char *base;
    size_t size, off;
    Workspace(void *p, size_t n) : base((char *)p), size(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (base == nullptr || off + bytes > size) { off = size + 1; return nullptr; }
        T *r = (T *)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};
static inline size_t ws_block(size_t bytes) { return align_up(bytes, 256); }

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, 1-D bulk async copy (TMA engine, SASS UBLKCP), tensor-map TMA, tcgen05.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dsir
