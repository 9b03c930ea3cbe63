// bulk_common.cuh -- mbarrier / bulk-copy (TMA engine) primitives and explicit shared-space loads shared by the
// staged step kernels (kernels_bulk.cu).
#pragma once
#include "step_common.cuh"

namespace pbk {
namespace {

// ---- mbarrier / bulk-copy primitives ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
/// global -> shared bulk copy (16-byte granularity); signals `bar` with the number of bytes delivered
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src)), "r"(bytes) : "memory");
}

/// A warp hands a ring stage back to the producer.  The stage was read with ld.shared (generic proxy) and will be
/// overwritten by cp.async.bulk (async proxy): program order + mbarrier release / acquire do not order accesses of two
/// different proxies, so every lane issues fence.proxy.async before the warp's arrival.  Without it the refill can overtake
/// reads that are still in flight -- seen as run-to-run differences of ~1e-6 in the moments of few-lane passes on the
/// 38 M-site system (profiles/r02_diag_r4c.log; compute-sanitizer racecheck had flagged exactly these pairs in round 1).
__device__ __forceinline__ void release_stage(uint32_t empty_bar, uint32_t tid) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if ((tid & 31u) == 0) mbar_arrive(empty_bar);
}

// ---- explicit shared-space loads (32-bit addresses: no generic-address arithmetic in the hot loop) ----
template<class CH> __device__ __forceinline__ CH lds_chunk(uint32_t addr) {
    static_assert(sizeof(CH) == 16, "16-byte chunks");
    int4 t;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
    return *reinterpret_cast<CH*>(&t);
}
__device__ __forceinline__ int32_t lds_i32(uint32_t addr) { int32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void lds_val(uint32_t addr, float& v) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); }
__device__ __forceinline__ void lds_val(uint32_t addr, float2& v) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)); }
__device__ __forceinline__ void lds_val(uint32_t addr, double& v) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); }
__device__ __forceinline__ void lds_val(uint32_t addr, double2& v) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); }

} // anonymous namespace
} // namespace pbk
