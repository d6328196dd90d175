// Shared device/host helpers for the vnext_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/vnext_b200.h"

namespace vnext {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char *fmt, ...);
void set_variant(const char *name);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int num_sms();

#define VNEXT_CHECK_ARG(cond, ...)              \
    do {                                        \
        if (!(cond)) {                          \
            ::vnext::set_error(__VA_ARGS__);    \
            return VNEXT_EINVAL;                \
        }                                       \
    } while (0)

#define VNEXT_CUDA(call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::vnext::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),         \
                               __FILE__, __LINE__);                                             \
            return VNEXT_ECUDA;                                                                 \
        }                                                                                       \
    } while (0)

// ---------------------------------------------------------------- device: mbarrier + 1-D TMA (cp.async.bulk)
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (TMA, non-tensor form).  bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 4 consecutive channels to fp32 (16 B) or bf16 (8 B) storage
__device__ __forceinline__ void store4(float *dst, const float4 &v) { *reinterpret_cast<float4 *>(dst) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16 *dst, const float4 &v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t *>(&lo);
    u.y = *reinterpret_cast<const uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(dst) = u;
}
__device__ __forceinline__ float4 load4(const float *src) { return *reinterpret_cast<const float4 *>(src); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16 *src) {
    const uint2 u = *reinterpret_cast<const uint2 *>(src);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
// vectorised fp32 reduction to global memory (sm_90+): one RED for four channels.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
#endif

}  // namespace vnext
