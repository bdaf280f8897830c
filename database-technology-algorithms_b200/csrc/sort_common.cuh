// sort_common.cuh -- device helpers shared by the onesweep kernels (kernels_sort.cu).
#pragma once
#include "dbt_internal.cuh"

namespace dbt {

constexpr int kRadix = 256;
constexpr uint32_t kFlagAgg = 0x40000000u; // tile aggregate available
constexpr uint32_t kFlagInc = 0x80000000u; // inclusive prefix available
constexpr uint32_t kValMask = 0x3FFFFFFFu;

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// tile-state words carry flag and value together, so relaxed gpu-scope accesses are enough
// (ld.volatile would be system scope: LDG.E.STRONG.SYS, measurably slower in the look-back loop)
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Tile-state words: two flag bits on top of the value.  32-bit words hold 30-bit values (n < 2^30 elements per sort
// call: the instantiation every hot path uses); 64-bit words lift that to the 2^32 - 2^16 elements u32 row ids allow.
template <typename S> struct TileState;
template <> struct TileState<uint32_t> {
    static constexpr uint32_t kAgg = 0x40000000u, kInc = 0x80000000u, kMask = 0x3FFFFFFFu;
    static constexpr int kFlagShift = 30;
    __device__ static __forceinline__ uint32_t ld(const uint32_t *p) { return ld_volatile(p); }
    __device__ static __forceinline__ void st(uint32_t *p, uint32_t v) { st_volatile(p, v); }
};
template <> struct TileState<uint64_t> {
    static constexpr uint64_t kAgg = 1ull << 62, kInc = 1ull << 63, kMask = (1ull << 62) - 1;
    static constexpr int kFlagShift = 62;
    __device__ static __forceinline__ uint64_t ld(const uint64_t *p) {
        uint64_t v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ static __forceinline__ void st(uint64_t *p, uint64_t v) {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


} // namespace dbt
