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


// The same butterfly for moduli below 2^29, where 8q fits a word: sums may stay in [0, 4q).
// bin is the compile-time bound of both inputs in units of q (1: canonical, 2, 4).  A sum of
// two values below 2q is left alone (it is below 4q), a sum of two values below 4q takes one
// conditional subtraction of 4q; the difference is offset by 2q or 4q so it stays positive
// and below 8q <= 2^32, which is all Shoup's estimate needs for a result in [0, 2q).  After
// the second stage of a round only the butterflies whose inputs were sums of the previous
// stage reduce at all -- half of them -- which takes 144 of the kernel's 416 VIADDMNMX away.
// The final values are canonicalised as before, so the output is unchanged bit for bit.
// (bin folds to a constant once the stage loops are unrolled)
__device__ __forceinline__ void gs_bfly_l4(const int bin, uint32_t &x, uint32_t &y, uint32_t w,
                                           uint32_t wp, uint32_t q, uint32_t two_q, uint32_t four_q,
                                           uint32_t zero) {
    uint32_t s = x + y + zero;
    uint32_t d = x - y + (bin == 4 ? four_q : two_q);
    if (bin == 4) s = min(s - four_q, s);
    uint32_t h = __umulhi(d, wp);
    x = s;
    y = d * w - h * q;
}
// input bound of the butterfly on registers (i0, i0 + 2^S) in stage S of a round whose inputs
// are bounded by BIN0 (1 or 2: nothing reduces before stage 2; 4: the round starts reduced)
__host__ __device__ constexpr int l4_bound(int S, int i0, int BIN0) {
    if (S == 0) return BIN0;
    if (S == 1 && BIN0 == 1) return 2;   // sums of canonical values are below 2q like the products
    return ((i0 >> (S - 1)) & 1) ? 2 : 4;
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


// The forward butterfly for moduli below 2^29: x may grow to 8q.  bin = bound of x in units
// of q; x + v and x - v + 2q are below (bin + 2) q, so x only needs its conditional
// subtraction (of 4q) when bin > 6 -- every other stage after the first three -- and y can be
// any word (Shoup's estimate holds for every 32-bit operand).
__host__ __device__ constexpr int ct_l4_out(int bin) { return (bin > 6 ? 4 : bin) + 2; }
__host__ __device__ constexpr int ct_l4_out_n(int bin, int stages) {
    for (int k = 0; k < stages; k++) bin = ct_l4_out(bin);
    return bin;
}
// The last stage of a transform (bin < 0, bound -bin) brings x below 2q instead, so that both
// outputs are below 4q and canonicalise with two conditional subtractions.
__device__ __forceinline__ void ct_bfly_l4(const int bin, uint32_t &x, uint32_t &y, uint32_t w,
                                           uint32_t wp, uint32_t q, uint32_t two_q, uint32_t four_q,
                                           uint32_t zero) {
    uint32_t xr = (bin > 6 || bin < -4) ? min(x - four_q, x) : x;
    if (bin < -2) xr = min(xr - two_q, xr);
    uint32_t h = __umulhi(y, wp);
    uint32_t v = y * w - h * q;
    x = xr + v + zero;
    y = xr - v + two_q;
}
// value below bin * q -> canonical
__device__ __forceinline__ uint32_t canon_l4(const int bin, uint32_t r, uint32_t q, uint32_t two_q,
                                             uint32_t four_q) {
    if (bin > 4) r = min(r - four_q, r);
    if (bin > 2) r = min(r - two_q, r);
    return min(r - q, r);
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



// ------------------------------------------------------------------ Tensor Memory
// TMEM (256 KiB per SM: 128 lanes x 512 columns x 32 bit) as per-thread storage.  A warp
// reaches the 32 lanes of its quadrant (warp id mod 4); with the 32x32b shape lane l of
// the warp reads/writes `x N` consecutive columns of TMEM lane 32*(warp%4) + l.  The
// NTT kernels keep each thread's PRIVATE twiddle pairs there (128 words per table: the
// 32 uint4 slots of gs_stage_t / ct_stage_t), which frees 32 KiB of shared memory per
// table and takes the twiddle reads off the shared-memory crossbar (measured on B200,
// profiles/microbench/tmem.cu: 326 B/clk/SM for x16 loads against 128 B/clk/SM for
// LDS.128), and park a transformed operand there while the second one is transformed
// (polymul).  No tensor-core instruction is involved.  SASS: LDTM / STTM.
__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_slot) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_slot)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t base) {    // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15])
        : "r"(taddr));
}
// The loaded registers pass THROUGH the wait ("+r"), so every use is ordered behind it.
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
                   "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Fill one private-twiddle table (128 columns) of the calling warp's lanes from global
// memory.  `src` points at element (slot 0, this thread's j) of a [32 slots][row_stride]
// uint4 layout.  All 16 warps of the CTA take part: the four warps that share a lane quadrant
// (warp, warp+4, ..) split the eight 16-column groups, two each, with eight loads in flight --
// the prologue of a 150 us launch must not cost 30 us of serial L2 round trips.
__device__ __forceinline__ void tmem_fill_table(uint32_t taddr_table, const uint4 *src, int row_stride,
                                                int warp) {
    const int s = warp >> 2;   // which two groups this warp writes
    uint32_t r[2][16];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int g = s + 4 * h;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint4 x = __ldg(src + (size_t) (4 * g + e) * row_stride);
            r[h][4 * e + 0] = x.x;
            r[h][4 * e + 1] = x.y;
            r[h][4 * e + 2] = x.z;
            r[h][4 * e + 3] = x.w;
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) tmem_st16(taddr_table + 16u * (s + 4 * h), r[h]);
}

// NB consecutive blocks B0.. of stage S (pairs i, i + 2^S) with the (w, w') pairs of those
// blocks in t[0..2*NB): the unit of work between two TMEM loads.
template <int S, int B0, int NB, bool REDUCE>
__device__ __forceinline__ void gs_blocks(uint32_t (&v)[64], const uint32_t *t, uint32_t q,
                                          uint32_t two_q, uint32_t zero) {
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            const int i0 = (B0 + k) * 2 * kStride + e;
            gs_bfly<REDUCE>(v[i0], v[i0 + kStride], t[2 * k], t[2 * k + 1], q, two_q, zero);
        }
    }
}
template <int S, int B0, int NB, bool REDUCE_X>
__device__ __forceinline__ void ct_blocks(uint32_t (&v)[64], const uint32_t *t, uint32_t q,
                                          uint32_t two_q, uint32_t zero) {
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            const int i0 = (B0 + k) * 2 * kStride + e;
            ct_bfly<REDUCE_X>(v[i0], v[i0 + kStride], t[2 * k], t[2 * k + 1], q, two_q, zero);
        }
    }
}

template <int S, int B0, int NB, int BIN0>
__device__ __forceinline__ void gs_blocks_l4(uint32_t (&v)[64], const uint32_t *t, uint32_t q,
                                             uint32_t two_q, uint32_t four_q, uint32_t zero) {
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            const int i0 = (B0 + k) * 2 * kStride + e;
            gs_bfly_l4(l4_bound(S, e, BIN0), v[i0], v[i0 + kStride], t[2 * k], t[2 * k + 1], q, two_q,
                       four_q, zero);
        }
    }
}
template <int S, int B0, int NB, int BIN>
__device__ __forceinline__ void ct_blocks_l4(uint32_t (&v)[64], const uint32_t *t, uint32_t q,
                                             uint32_t two_q, uint32_t four_q, uint32_t zero) {
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            const int i0 = (B0 + k) * 2 * kStride + e;
            ct_bfly_l4(BIN, v[i0], v[i0 + kStride], t[2 * k], t[2 * k + 1], q, two_q, four_q, zero);
        }
    }
}

// GS stages 0..5 on the thread's 64 registers, private twiddles from the thread's
// 128-word TMEM table at `taddr` (slot layout of fused_prepare / tile_table_kernel:
// slots 0-15 stage 0, 16-23 stage 1, 24-27 stage 2, 28-29 stage 3, 30 stage 4, 31 stage 5;
// one x16 load = 4 slots = 8 pairs).  The next group's load is in flight while the current
// group's butterflies run.
template <bool REDUCE0>
__device__ __forceinline__ void gs_round_tmem(uint32_t (&v)[64], uint32_t taddr, uint32_t q,
                                              uint32_t two_q, uint32_t zero) {
    uint32_t ta[16], tb[16];
    tmem_ld16(taddr, ta);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 16, tb);
    gs_blocks<0, 0, 8, REDUCE0>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 32, ta);
    gs_blocks<0, 8, 8, REDUCE0>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 48, tb);
    gs_blocks<0, 16, 8, REDUCE0>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 64, ta);
    gs_blocks<0, 24, 8, REDUCE0>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 80, tb);
    gs_blocks<1, 0, 8, true>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 96, ta);
    gs_blocks<1, 8, 8, true>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 112, tb);
    gs_blocks<2, 0, 8, true>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    gs_blocks<3, 0, 4, true>(v, tb, q, two_q, zero);
    gs_blocks<4, 0, 2, true>(v, tb + 8, q, two_q, zero);
    gs_blocks<5, 0, 1, true>(v, tb + 12, q, two_q, zero);
}

// CT stages 5..0 (stride 32 -> 1 inside the thread's 64 contiguous coefficients), same table
template <bool REDUCE_FIRST>
__device__ __forceinline__ void ct_round_tmem(uint32_t (&v)[64], uint32_t taddr, uint32_t q,
                                              uint32_t two_q, uint32_t zero) {
    uint32_t ta[16], tb[16];
    tmem_ld16(taddr + 112, tb);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 96, ta);
    ct_blocks<5, 0, 1, REDUCE_FIRST>(v, tb + 12, q, two_q, zero);
    ct_blocks<4, 0, 2, true>(v, tb + 8, q, two_q, zero);
    ct_blocks<3, 0, 4, true>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 80, tb);
    ct_blocks<2, 0, 8, true>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 64, ta);
    ct_blocks<1, 8, 8, true>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 48, tb);
    ct_blocks<1, 0, 8, true>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 32, ta);
    ct_blocks<0, 24, 8, true>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 16, tb);
    ct_blocks<0, 16, 8, true>(v, ta, q, two_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr, ta);
    ct_blocks<0, 8, 8, true>(v, tb, q, two_q, zero);
    tmem_wait_ld16(ta);
    ct_blocks<0, 0, 8, true>(v, ta, q, two_q, zero);
}

// 4q-lazy forms of the two rounds above (q < 2^29).  GS: inputs bounded by BIN0 * q, outputs
// below 4q (sums) / 2q (products).  CT: inputs below BIN0 * q, outputs below
// ct_l4_out_n(BIN0, 6) * q -- or below 4q when the round is the LAST of the transform.
template <int BIN0>
__device__ __forceinline__ void gs_round_tmem_l4(uint32_t (&v)[64], uint32_t taddr, uint32_t q,
                                                 uint32_t two_q, uint32_t four_q, uint32_t zero) {
    uint32_t ta[16], tb[16];
    tmem_ld16(taddr, ta);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 16, tb);
    gs_blocks_l4<0, 0, 8, BIN0>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 32, ta);
    gs_blocks_l4<0, 8, 8, BIN0>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 48, tb);
    gs_blocks_l4<0, 16, 8, BIN0>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 64, ta);
    gs_blocks_l4<0, 24, 8, BIN0>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 80, tb);
    gs_blocks_l4<1, 0, 8, BIN0>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 96, ta);
    gs_blocks_l4<1, 8, 8, BIN0>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 112, tb);
    gs_blocks_l4<2, 0, 8, BIN0>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    gs_blocks_l4<3, 0, 4, BIN0>(v, tb, q, two_q, four_q, zero);
    gs_blocks_l4<4, 0, 2, BIN0>(v, tb + 8, q, two_q, four_q, zero);
    gs_blocks_l4<5, 0, 1, BIN0>(v, tb + 12, q, two_q, four_q, zero);
}
template <int BIN0, bool LAST = false>
__device__ __forceinline__ void ct_round_tmem_l4(uint32_t (&v)[64], uint32_t taddr, uint32_t q,
                                                 uint32_t two_q, uint32_t four_q, uint32_t zero) {
    constexpr int B5 = BIN0, B4 = ct_l4_out(B5), B3 = ct_l4_out(B4), B2 = ct_l4_out(B3),
                  B1 = ct_l4_out(B2), B0 = ct_l4_out(B1);
    uint32_t ta[16], tb[16];
    tmem_ld16(taddr + 112, tb);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 96, ta);
    ct_blocks_l4<5, 0, 1, B5>(v, tb + 12, q, two_q, four_q, zero);
    ct_blocks_l4<4, 0, 2, B4>(v, tb + 8, q, two_q, four_q, zero);
    ct_blocks_l4<3, 0, 4, B3>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 80, tb);
    ct_blocks_l4<2, 0, 8, B2>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 64, ta);
    ct_blocks_l4<1, 8, 8, B1>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 48, tb);
    ct_blocks_l4<1, 0, 8, B1>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 32, ta);
    constexpr int BL = LAST ? -B0 : B0;
    ct_blocks_l4<0, 24, 8, BL>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 16, tb);
    ct_blocks_l4<0, 16, 8, BL>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr, ta);
    ct_blocks_l4<0, 8, 8, BL>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    ct_blocks_l4<0, 0, 8, BL>(v, ta, q, two_q, four_q, zero);
}

// ---- pieces of the forward kernels that finish a row in two halves (tile_ct_h_kernel,
// polyt_ct_kernel, tilecol_ct_kernel): classic or 4q-lazy blocks, and stages 4..0 of one half
// with the private pairs streaming from tensor memory
template <int S, int B0, int NB, int BIN, bool L4>
__device__ __forceinline__ void ct_blocks_sel(uint32_t (&v)[64], const uint32_t *t, uint32_t q,
                                              uint32_t two_q, uint32_t four_q, uint32_t zero) {
    if (L4) {
        ct_blocks_l4<S, B0, NB, BIN>(v, t, q, two_q, four_q, zero);
    } else {
        ct_blocks<S, B0, NB, true>(v, t, q, two_q, zero);
    }
}
// stages 4..0 of half H (registers 32 H .. 32 H + 31); ts2 / ts345: the x16 groups at
// columns 96 and 112 (stage 2, stages 3-5), already loaded
template <int H, int B4, bool L4>
__device__ __forceinline__ void ct_half_tmem(uint32_t (&v)[64], uint32_t taddr, const uint32_t *ts2,
                                             const uint32_t *ts345, uint32_t q, uint32_t two_q,
                                             uint32_t four_q, uint32_t zero) {
    constexpr int B3 = ct_l4_out(B4), B2 = ct_l4_out(B3), B1 = ct_l4_out(B2), B0 = ct_l4_out(B1);
    uint32_t ta[16], tb[16];
    tmem_ld16(taddr + 64 + 16 * H, ta);                       // stage 1, blocks 8H .. 8H+7
    ct_blocks_sel<4, H, 1, B4, L4>(v, ts345 + 8 + 2 * H, q, two_q, four_q, zero);
    ct_blocks_sel<3, 2 * H, 2, B3, L4>(v, ts345 + 4 * H, q, two_q, four_q, zero);
    ct_blocks_sel<2, 4 * H, 4, B2, L4>(v, ts2 + 8 * H, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    tmem_ld16(taddr + 32 * H + 16, tb);                       // stage 0, blocks 16H+8 .. 16H+15
    ct_blocks_sel<1, 8 * H, 8, B1, L4>(v, ta, q, two_q, four_q, zero);
    tmem_wait_ld16(tb);
    tmem_ld16(taddr + 32 * H, ta);                            // stage 0, blocks 16H .. 16H+7
    ct_blocks_sel<0, 16 * H + 8, 8, B0, L4>(v, tb, q, two_q, four_q, zero);
    tmem_wait_ld16(ta);
    ct_blocks_sel<0, 16 * H, 8, B0, L4>(v, ta, q, two_q, four_q, zero);
}

}  // namespace nttb200
