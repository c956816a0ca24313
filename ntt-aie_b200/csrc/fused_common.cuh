// fused_common.cuh -- device helpers shared by the register-radix kernels
// (kernels_fused.cu, kernels_multi.cu): TMA / mbarrier PTX wrappers, shared-memory
// accessors and the lazy Harvey butterfly.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "modarith.cuh"

namespace nttb200 {

constexpr int kF_Team = 64;               // threads per 4096-coefficient tile
constexpr int kF_PolyBytes = 4096 * 4;    // one tile

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(64) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                       uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
                 "r"(d)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ------------------------------------------------------------------ butterfly
// Lazy GS butterfly on values in [0, 2q).  `zero` is an opaque runtime 0 that keeps
// the add a 3-input IADD3 on the ALU pipe (ptxas would otherwise turn half of the
// plain adds into IMAD.IADD on the FMA pipe, which the three multiplies saturate).
template <bool REDUCE>
__device__ __forceinline__ void gs_bfly(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp,
                                        uint32_t q, uint32_t two_q, uint32_t zero) {
    uint32_t s = x + y + zero;
    uint32_t d = x - y + two_q;
    if (REDUCE) s = min(s - two_q, s);
    uint32_t h = __umulhi(d, wp);
    x = s;
    y = d * w - h * q;
}


// Lazy Harvey Cooley-Tukey butterfly on values in [0, 4q): x is brought to [0, 2q),
// v = y*w in [0, 2q) by Shoup; outputs x+v and x-v+2q, both in [0, 4q).  4q < 2^32.
template <bool REDUCE_X>
__device__ __forceinline__ void ct_bfly(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wp,
                                        uint32_t q, uint32_t two_q, uint32_t zero) {
    uint32_t xr = REDUCE_X ? min(x - two_q, x) : x;
    uint32_t h = __umulhi(y, wp);
    uint32_t v = y * w - h * q;
    x = xr + v + zero;
    y = xr - v + two_q;
}


__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1,
                                             int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
        "r"(src), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}


}  // namespace nttb200
