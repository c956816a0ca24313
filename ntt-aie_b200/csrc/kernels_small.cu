// kernels_small.cu -- whole-transform register-radix kernels for N = 2^6 .. 2^11,
// which includes the reference's own default N = 2048 (reference src/test.cpp:66,
// src/aie2.py:14).
//
// One WARP owns a block of 2048 consecutive coefficients = 2^(11-logn) polynomials:
//   TMA (128B swizzle) HBM -> shared, 8 KiB per warp
//   round 1  lane j owns a[64j .. 64j+63]: golden stages 0-5 in registers, its 63
//            private (w, w') pairs from a shared-memory table (conflict-free LDS.128)
//   exchange through the same buffer, warp-synchronous (__syncwarp, no block barrier)
//   round 2  lane j owns columns 2j, 2j+1 of all 32 rows (LDS.64): the remaining
//            logn-6 stages pair rows; twiddles depend only on the register index, so
//            they are kernel parameters (constant-bank / uniform-register operands)
//   store    canonicalise, one 256 B row per STG.64 warp instruction
// 16 warps per CTA, one persistent CTA per SM; the next block's TMA load is issued as
// soon as round 2 has pulled the current one out of shared memory.
// Butterflies and laziness as in kernels_fused.cu; bit-exact against the golden.
#include <cuda.h>

#include <vector>

#include "fused_common.cuh"
#include "plan.h"

namespace nttb200 {

constexpr int kS_Warps = 16;
constexpr int kS_Threads = kS_Warps * 32;
constexpr int kS_BlockBytes = 2048 * 4;                 // one warp's block
constexpr int kS_TwBytes = 32 * 32 * 16;                // [32 slots][32 lanes] uint4
constexpr int kS_SmemBytes = kS_TwBytes + kS_Warps * kS_BlockBytes + 16 * 8 + 1024;
constexpr int kS_SmemBytesCt = kS_SmemBytes + kS_Warps * (kS_BlockBytes / 2) + 1024;   // + staging slots

struct SmallParams {
    uint32_t *out;
    const uint4 *tw_r1;   // [32 slots][32 lanes]
    uint64_t blocks;      // 2048-coefficient blocks in the batch
    uint32_t q;
    uint32_t zero;
    uint32_t qinv;        // DUAL: q^-1 mod 2^32 (Montgomery product of the two operands)
    uint32_t scale;       // DUAL: every output times this constant (N^-1 * 2^32 mod q) ...
    uint32_t scale_shoup; //       ... and its Shoup companion
    uint32_t four_q;      // opaque 4q for the 4q-lazy butterflies (q < 2^29)
};

// round-1 stage (same slot scheme as kernels_fused.cu, 32 lanes per slot)
// BIN0 > 0: 4q-lazy form (fused_common.cuh gs_bfly_l4), inputs of the round bounded by BIN0 * q
template <int S, bool REDUCE0 = false, int BIN0 = 0>
__device__ __forceinline__ void small_r1_stage(uint32_t (&v)[64], uint32_t tw_addr, uint32_t q,
                                               uint32_t two_q, uint32_t zero, uint32_t four_q = 0) {
    constexpr int kBlocks = 32 >> S;
    constexpr int kSlot0 = 32 - (kBlocks >= 2 ? kBlocks : 1);
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int b = 0; b < kBlocks; b += 2) {
        uint4 t = lds128(tw_addr + (kSlot0 + b / 2) * (32 * 16));
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            int i0 = b * 2 * kStride + e;
            if (BIN0 > 0) {
                gs_bfly_l4(l4_bound(S, e, BIN0), v[i0], v[i0 + kStride], t.x, t.y, q, two_q, four_q, zero);
                continue;
            }
            gs_bfly<(S > 0 || REDUCE0)>(v[i0], v[i0 + kStride], t.x, t.y, q, two_q, zero);
        }
        if (kBlocks >= 2) {
#pragma unroll
            for (int e = 0; e < kStride; e++) {
                int i0 = (b + 1) * 2 * kStride + e;
                if (BIN0 > 0) {
                    gs_bfly_l4(l4_bound(S, e, BIN0), v[i0], v[i0 + kStride], t.z, t.w, q, two_q, four_q, zero);
                    continue;
                }
                gs_bfly<(S > 0 || REDUCE0)>(v[i0], v[i0 + kStride], t.z, t.w, q, two_q, zero);
            }
        }
    }
}

// round-2 stage K on registers v[2*row + col]: pairs rows i and i + 2^K inside each
// polynomial (RP = 2^(LOGN-6) rows per polynomial); twiddle table[(RP >> (K+1)) + blk]
template <int LOGN, int K, bool L4 = false>
__device__ __forceinline__ void small_r2_stage(uint32_t (&v)[64], const UniformTw &u, uint32_t q,
                                               uint32_t two_q, uint32_t zero, uint32_t four_q = 0) {
    constexpr int RP = 1 << (LOGN - 6);
#pragma unroll
    for (int i = 0; i < 32; i++) {
        if (i & (1 << K)) continue;
        const int blk = (i & (RP - 1)) >> (K + 1);
        const uint32_t w = u.w[(RP >> (K + 1)) + blk], wp = u.wp[(RP >> (K + 1)) + blk];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (L4) {   // rows come out of round 1 below 4q whatever the lane
                gs_bfly_l4(l4_bound(K, i, 4), v[2 * i + c], v[2 * (i + (1 << K)) + c], w, wp, q, two_q,
                           four_q, zero);
            } else {
                gs_bfly<true>(v[2 * i + c], v[2 * (i + (1 << K)) + c], w, wp, q, two_q, zero);
            }
        }
    }
}

// DUAL: the block's input is the pointwise product of two buffers (Montgomery product
// a*b*2^-32, second operand through the same shared buffer) and every output is
// multiplied by `scale` -- the tail of a negacyclic multiplication in one kernel.
template <int LOGN, bool PERMUTE, bool DUAL, bool L4 = false>
__global__ void __launch_bounds__(kS_Threads, 1)
fused_gs_small_kernel(const __grid_constant__ CUtensorMap map_lo,
                      const __grid_constant__ CUtensorMap map_hi,
                      const __grid_constant__ CUtensorMap map_b_lo,
                      const __grid_constant__ CUtensorMap map_b_hi,
                      const __grid_constant__ UniformTw uni, const SmallParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t tw_base = smem_base;
    const uint32_t data_base = smem_base + kS_TwBytes;
    const uint32_t bar_base = data_base + kS_Warps * kS_BlockBytes;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
    const int j = tid & 31;
    const uint32_t q = prm.q, two_q = 2u * prm.q, four_q = prm.four_q, zero = prm.zero;

    for (int i = tid; i < 32 * 32; i += kS_Threads) {
        uint4 t = __ldg(prm.tw_r1 + i);
        sts128(tw_base + i * 16, t.x, t.y, t.z, t.w);
    }
    if (tid < kS_Warps) mbar_init(bar_base + tid * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t buf = data_base + warp * kS_BlockBytes;
    const uint32_t bar = bar_base + warp * 8;
    const uint64_t stride = (uint64_t) gridDim.x * kS_Warps;
    uint64_t blk = (uint64_t) warp * gridDim.x + blockIdx.x;   // consecutive blocks on different SMs
    uint32_t parity = 0;
    if (j == 0 && blk < prm.blocks) {
        mbar_expect_tx(bar, kS_BlockBytes);
        tma_load_3d(buf, &map_lo, bar, 0, 0, (int) blk);
        tma_load_3d(buf + kS_BlockBytes / 2, &map_hi, bar, 0, 0, (int) blk);
    }
    // buffer layout: two halves of [32 rows][32 words]; row r / half h holds
    // a[64r + 32h .. +31]; 16-byte chunk index XOR (r & 7)
    const uint32_t r1_row = buf + j * 128;
    const uint32_t r1_xor = (j & 7) << 4;
    // round 2: lane j owns words 2j, 2j+1 of every 64-word row: half j>>4, word 2(j&15)
    const uint32_t r2_col = buf + (j >> 4) * (kS_BlockBytes / 2) + (j & 1) * 8;
    const uint32_t r2_chunk = ((j & 15) >> 1) << 4;
    const uint32_t tw_addr = tw_base + j * 16;

    for (; blk < prm.blocks; blk += stride) {
        uint32_t v[64];
        mbar_wait(bar, parity);
        parity ^= 1;
#pragma unroll
        for (int c = 0; c < 16; c++) {
            uint4 t = lds128(r1_row + (c >> 3) * (kS_BlockBytes / 2) + (((c & 7) << 4) ^ r1_xor));
            v[4 * c + 0] = t.x;
            v[4 * c + 1] = t.y;
            v[4 * c + 2] = t.z;
            v[4 * c + 3] = t.w;
        }
        if (DUAL) {
            __syncwarp();
            if (j == 0) {
                mbar_expect_tx(bar, kS_BlockBytes);
                tma_load_3d(buf, &map_b_lo, bar, 0, 0, (int) blk);
                tma_load_3d(buf + kS_BlockBytes / 2, &map_b_hi, bar, 0, 0, (int) blk);
            }
            mbar_wait(bar, parity);
            parity ^= 1;
#pragma unroll
            for (int c = 0; c < 16; c++) {
                uint4 t = lds128(r1_row + (c >> 3) * (kS_BlockBytes / 2) + (((c & 7) << 4) ^ r1_xor));
                const uint32_t bb[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    uint64_t prod = (uint64_t) v[4 * c + e] * bb[e];
                    uint32_t m = (uint32_t) prod * prm.qinv;
                    v[4 * c + e] = (uint32_t) (prod >> 32) - __umulhi(m, q) + q;
                }
            }
        }
        constexpr int kB0 = L4 ? (DUAL ? 2 : 1) : 0;
        small_r1_stage<0, DUAL, kB0>(v, tw_addr, q, two_q, zero, four_q);
        small_r1_stage<1, false, kB0>(v, tw_addr, q, two_q, zero, four_q);
        small_r1_stage<2, false, kB0>(v, tw_addr, q, two_q, zero, four_q);
        small_r1_stage<3, false, kB0>(v, tw_addr, q, two_q, zero, four_q);
        small_r1_stage<4, false, kB0>(v, tw_addr, q, two_q, zero, four_q);
        small_r1_stage<5, false, kB0>(v, tw_addr, q, two_q, zero, four_q);
#pragma unroll
        for (int c = 0; c < 16; c++) {
            sts128(r1_row + (c >> 3) * (kS_BlockBytes / 2) + (((c & 7) << 4) ^ r1_xor), v[4 * c],
                   v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i++) {
            uint32_t addr = r2_col + i * 128 + (r2_chunk ^ ((i & 7) << 4));
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                         : "=r"(v[2 * i]), "=r"(v[2 * i + 1])
                         : "r"(addr));
        }
        fence_proxy_async();
        __syncwarp();
        const uint64_t next = blk + stride;
        if (j == 0 && next < prm.blocks) {
            mbar_expect_tx(bar, kS_BlockBytes);
            tma_load_3d(buf, &map_lo, bar, 0, 0, (int) next);
            tma_load_3d(buf + kS_BlockBytes / 2, &map_hi, bar, 0, 0, (int) next);
        }
        if (LOGN > 6) small_r2_stage<LOGN, 0, L4>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 7) small_r2_stage<LOGN, 1, L4>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 8) small_r2_stage<LOGN, 2, L4>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 9) small_r2_stage<LOGN, 3, L4>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 10) small_r2_stage<LOGN, 4, L4>(v, uni, q, two_q, zero, four_q);

        // register (row i, col c) is coefficient 64 i + 2j + c of the block
        uint32_t *dst = prm.out + blk * 2048 + 2 * j;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            int row = i;
            if (PERMUTE && LOGN == 11) {
                // ans_order on the top four of the 11 index bits = row bits 1..4
                int b = i >> 1;
                b = ((b & 0x5) << 1) | ((b & 0xA) >> 1);
                row = (b << 1) | (i & 1);
            }
            uint32_t r0 = v[2 * i], r1 = v[2 * i + 1];
            if (DUAL) {   // any word in, [0, 2q) out
                r0 = shoup_mul_lazy(r0, prm.scale, prm.scale_shoup, q);
                r1 = shoup_mul_lazy(r1, prm.scale, prm.scale_shoup, q);
            } else if (L4 && (LOGN == 6 || !((i >> (LOGN - 7)) & 1))) {
                // a sum of the last stage (or anything out of round 1): below 4q
                r0 = min(r0 - two_q, r0);
                r1 = min(r1 - two_q, r1);
            }
            uint2 o;
            o.x = min(r0 - q, r0);
            o.y = min(r1 - q, r1);
            *reinterpret_cast<uint2 *>(dst + row * 64) = o;
        }
    }
}


// round-2-type CT stage K (rows i, i + 2^K of both columns), uniform twiddles
// BIN > 0: 4q-lazy butterfly with input bound BIN (fused_common.cuh ct_bfly_l4)
template <int LOGN, int K, bool REDUCE_X, int BIN = 0>
__device__ __forceinline__ void small_ct_col_stage(uint32_t (&v)[64], const UniformTw &u, uint32_t q,
                                                   uint32_t two_q, uint32_t zero, uint32_t four_q = 0) {
    constexpr int RP = 1 << (LOGN - 6);
#pragma unroll
    for (int i = 0; i < 32; i++) {
        if (i & (1 << K)) continue;
        const int blk = (i & (RP - 1)) >> (K + 1);
        const uint32_t w = u.w[(RP >> (K + 1)) + blk], wp = u.wp[(RP >> (K + 1)) + blk];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (BIN > 0) {
                ct_bfly_l4(BIN, v[2 * i + c], v[2 * (i + (1 << K)) + c], w, wp, q, two_q, four_q, zero);
            } else {
                ct_bfly<REDUCE_X>(v[2 * i + c], v[2 * (i + (1 << K)) + c], w, wp, q, two_q, zero);
            }
        }
    }
}

// round-1-type CT stage S on 64 contiguous coefficients, thread-private twiddles
template <int S, bool REDUCE_X, int BIN = 0>
__device__ __forceinline__ void small_ct_row_stage(uint32_t (&v)[64], uint32_t tw_addr, uint32_t q,
                                                   uint32_t two_q, uint32_t zero, uint32_t four_q = 0) {
    constexpr int kBlocks = 32 >> S;
    constexpr int kSlot0 = 32 - (kBlocks >= 2 ? kBlocks : 1);
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int b = 0; b < kBlocks; b += 2) {
        uint4 t = lds128(tw_addr + (kSlot0 + b / 2) * (32 * 16));
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            int i0 = b * 2 * kStride + e;
            if (BIN > 0) {
                ct_bfly_l4(BIN, v[i0], v[i0 + kStride], t.x, t.y, q, two_q, four_q, zero);
                continue;
            }
            ct_bfly<REDUCE_X>(v[i0], v[i0 + kStride], t.x, t.y, q, two_q, zero);
        }
        if (kBlocks >= 2) {
#pragma unroll
            for (int e = 0; e < kStride; e++) {
                int i0 = (b + 1) * 2 * kStride + e;
                if (BIN > 0) {
                    ct_bfly_l4(BIN, v[i0], v[i0 + kStride], t.z, t.w, q, two_q, four_q, zero);
                    continue;
                }
                ct_bfly<REDUCE_X>(v[i0], v[i0 + kStride], t.z, t.w, q, two_q, zero);
            }
        }
    }
}

// stage S < 5 restricted to the registers [32 H, 32 H + 32) (after stage 5 the two halves of a
// 64-coefficient row are independent)
template <int S, int H, int BIN>
__device__ __forceinline__ void small_ct_row_half(uint32_t (&v)[64], uint32_t tw_addr, uint32_t q,
                                                  uint32_t two_q, uint32_t zero, uint32_t four_q) {
    constexpr int kBlocks = 32 >> S;
    constexpr int kSlot0 = 32 - kBlocks;
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int b = H * (kBlocks / 2); b < (H + 1) * (kBlocks / 2); b++) {
        const uint4 t = lds128(tw_addr + (kSlot0 + b / 2) * (32 * 16));
        const uint32_t w = (b & 1) ? t.z : t.x, wp = (b & 1) ? t.w : t.y;
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            const int i0 = b * 2 * kStride + e;
            if (BIN > 0) {
                ct_bfly_l4(BIN, v[i0], v[i0 + kStride], w, wp, q, two_q, four_q, zero);
            } else {
                ct_bfly<true>(v[i0], v[i0 + kStride], w, wp, q, two_q, zero);
            }
        }
    }
}

// Forward (Cooley-Tukey) partner, one warp per 2048-coefficient block: columns first
// (stages logn-1 .. 6, uniform twiddles), warp-synchronous exchange, rows (stages 5 .. 0);
// the canonical rows leave through a TMA store in two halves through a 4 KiB staging slot per
// warp (see tile_ct_h_kernel, kernels_multi.cu): the block buffer takes the prefetch of the
// warp's next block as soon as the rows are in registers, and the left half's store drains
// while the right half's five stages run.
template <int LOGN, bool L4 = false>
__global__ void __launch_bounds__(kS_Threads, 1)
fused_ct_small_kernel(const __grid_constant__ CUtensorMap map_lo,
                      const __grid_constant__ CUtensorMap map_hi,
                      const __grid_constant__ CUtensorMap out_lo,
                      const __grid_constant__ CUtensorMap out_hi,
                      const __grid_constant__ UniformTw uni, const SmallParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t tw_base = smem_base;
    const uint32_t data_base = smem_base + kS_TwBytes;
    const uint32_t bar_base = data_base + kS_Warps * kS_BlockBytes;
    const uint32_t stage_base = (bar_base + 16 * 8 + 1023u) & ~1023u;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int j = tid & 31;
    const uint32_t q = prm.q, two_q = 2u * prm.q, four_q = prm.four_q, zero = prm.zero;
    // 4q-lazy bounds entering column stage K (4 .. 0, present for K < LOGN - 6) and the rows
    constexpr int kC4 = 1, kC3 = LOGN > 10 ? ct_l4_out(kC4) : 1, kC2 = LOGN > 9 ? ct_l4_out(kC3) : 1,
                  kC1 = LOGN > 8 ? ct_l4_out(kC2) : 1, kC0 = LOGN > 7 ? ct_l4_out(kC1) : 1,
                  kR5 = LOGN > 6 ? ct_l4_out(kC0) : 1, kR4 = ct_l4_out(kR5), kR3 = ct_l4_out(kR4),
                  kR2 = ct_l4_out(kR3), kR1 = ct_l4_out(kR2), kR0 = ct_l4_out(kR1), kEnd = ct_l4_out(kR0);

    for (int i = tid; i < 32 * 32; i += kS_Threads) {
        uint4 t = __ldg(prm.tw_r1 + i);
        sts128(tw_base + i * 16, t.x, t.y, t.z, t.w);
    }
    if (tid < kS_Warps) mbar_init(bar_base + tid * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t buf = data_base + warp * kS_BlockBytes;
    const uint32_t bar = bar_base + warp * 8;
    const uint64_t stride = (uint64_t) gridDim.x * kS_Warps;
    uint64_t blk = (uint64_t) warp * gridDim.x + blockIdx.x;   // consecutive blocks on different SMs
    uint32_t parity = 0;
    if (j == 0 && blk < prm.blocks) {
        mbar_expect_tx(bar, kS_BlockBytes);
        tma_load_3d(buf, &map_lo, bar, 0, 0, (int) blk);
        tma_load_3d(buf + kS_BlockBytes / 2, &map_hi, bar, 0, 0, (int) blk);
    }
    const uint32_t stg = stage_base + warp * (kS_BlockBytes / 2);
    const uint32_t r1_row = buf + j * 128;
    const uint32_t st_row = stg + j * 128;
    const uint32_t r1_xor = (j & 7) << 4;
    const uint32_t r2_col = buf + (j >> 4) * (kS_BlockBytes / 2) + (j & 1) * 8;
    const uint32_t r2_chunk = ((j & 15) >> 1) << 4;
    const uint32_t tw_addr = tw_base + j * 16;

    for (; blk < prm.blocks; blk += stride) {
        uint32_t v[64];
        mbar_wait(bar, parity);
        parity ^= 1;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            uint32_t addr = r2_col + i * 128 + (r2_chunk ^ ((i & 7) << 4));
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                         : "=r"(v[2 * i]), "=r"(v[2 * i + 1])
                         : "r"(addr));
        }
        // stages logn-1 .. 6: the first one sees canonical inputs
        if (LOGN > 10) small_ct_col_stage<LOGN, 4, false, (L4 ? kC4 : 0)>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 9) small_ct_col_stage<LOGN, 3, (LOGN > 10), (L4 ? kC3 : 0)>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 8) small_ct_col_stage<LOGN, 2, (LOGN > 9), (L4 ? kC2 : 0)>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 7) small_ct_col_stage<LOGN, 1, (LOGN > 8), (L4 ? kC1 : 0)>(v, uni, q, two_q, zero, four_q);
        if (LOGN > 6) small_ct_col_stage<LOGN, 0, (LOGN > 7), (L4 ? kC0 : 0)>(v, uni, q, two_q, zero, four_q);
#pragma unroll
        for (int i = 0; i < 32; i++) {
            uint32_t addr = r2_col + i * 128 + (r2_chunk ^ ((i & 7) << 4));
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v[2 * i]), "r"(v[2 * i + 1])
                         : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 16; c++) {
            uint4 t = lds128(r1_row + (c >> 3) * (kS_BlockBytes / 2) + (((c & 7) << 4) ^ r1_xor));
            v[4 * c + 0] = t.x;
            v[4 * c + 1] = t.y;
            v[4 * c + 2] = t.z;
            v[4 * c + 3] = t.w;
        }
        // ---- the block buffer is free: prefetch the warp's next block; the staging slot is free
        // once the previous block's right half has been read by its store
        fence_proxy_async();
        if (j == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        const uint64_t next = blk + stride;
        if (j == 0 && next < prm.blocks) {
            mbar_expect_tx(bar, kS_BlockBytes);
            tma_load_3d(buf, &map_lo, bar, 0, 0, (int) next);
            tma_load_3d(buf + kS_BlockBytes / 2, &map_hi, bar, 0, 0, (int) next);
        }
        small_ct_row_stage<5, (LOGN > 6), (L4 ? kR5 : 0)>(v, tw_addr, q, two_q, zero, four_q);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (h == 0) {
                small_ct_row_half<4, 0, (L4 ? kR4 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<3, 0, (L4 ? kR3 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<2, 0, (L4 ? kR2 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<1, 0, (L4 ? kR1 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<0, 0, (L4 ? kR0 : 0)>(v, tw_addr, q, two_q, zero, four_q);
            } else {
                small_ct_row_half<4, 1, (L4 ? kR4 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<3, 1, (L4 ? kR3 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<2, 1, (L4 ? kR2 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<1, 1, (L4 ? kR1 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                small_ct_row_half<0, 1, (L4 ? kR0 : 0)>(v, tw_addr, q, two_q, zero, four_q);
                // the left half's store has read the slot while these stages ran
                if (j == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    uint32_t r = v[32 * h + 4 * c + e];
                    if (L4) {
                        o[e] = canon_l4(kEnd, r, q, two_q, four_q);
                    } else {
                        r = min(r - two_q, r);
                        o[e] = min(r - q, r);
                    }
                }
                sts128(st_row + ((c << 4) ^ r1_xor), o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (j == 0) {
                tma_store_3d(h == 0 ? &out_lo : &out_hi, stg, 0, 0, (int) blk);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (j == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// --------------------------------------------------------------------- host side
int encode_tile_map(CUtensorMap *map, const int32_t *base, uint32_t rows, size_t blocks,
                    size_t tile_stride_bytes = 0);  // kernels_fused.cu

template <int LOGN>
static int small_set_attr() {
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    NTTB200_CUDA(cudaFuncSetAttribute(fused_gs_small_kernel<LOGN, false, false>, attr, kS_SmemBytes));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_gs_small_kernel<LOGN, true, false>, attr, kS_SmemBytes));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_gs_small_kernel<LOGN, false, true>, attr, kS_SmemBytes));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_gs_small_kernel<LOGN, false, false, true>, attr, kS_SmemBytes));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_gs_small_kernel<LOGN, true, false, true>, attr, kS_SmemBytes));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_gs_small_kernel<LOGN, false, true, true>, attr, kS_SmemBytes));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_ct_small_kernel<LOGN>, attr, kS_SmemBytesCt));
    NTTB200_CUDA(cudaFuncSetAttribute(fused_ct_small_kernel<LOGN, true>, attr, kS_SmemBytesCt));
    return NTTB200_OK;
}

int small_prepare(nttb200_plan *p) {
    if (p->logn < 6 || p->logn > 11) return NTTB200_ERR_UNSUPPORTED;
    const uint32_t n = p->n;
    const int tpp = (int) (n >> 6);  // lanes per polynomial
    std::vector<uint2> host(n);
    NTTB200_CUDA(cudaMemcpy(host.data(), p->d_tw, sizeof(uint2) * n, cudaMemcpyDeviceToHost));
    // stage s block b of lane j: table[(n >> (s+1)) + (j % tpp)*(32 >> s) + b]
    std::vector<uint4> r1(32 * 32);
    for (int s = 0; s < 6; s++) {
        const int blocks = 32 >> s;
        const int slot0 = 32 - (blocks >= 2 ? blocks : 1);
        for (int j = 0; j < 32; j++) {
            size_t base = (size_t) (n >> (s + 1)) + (size_t) (j % tpp) * blocks;
            for (int b = 0; b < blocks; b += 2) {
                uint2 t0 = host[base + b];
                uint2 t1 = blocks >= 2 ? host[base + b + 1] : make_uint2(0, 0);
                r1[(size_t) (slot0 + b / 2) * 32 + j] = make_uint4(t0.x, t0.y, t1.x, t1.y);
            }
        }
    }
    NTTB200_CUDA(cudaMalloc(&p->d_tw_r1, sizeof(uint4) * r1.size()));
    NTTB200_CUDA(cudaMemcpy(p->d_tw_r1, r1.data(), sizeof(uint4) * r1.size(),
                            cudaMemcpyHostToDevice));
    for (int i = 0; i < 64; i++) {
        p->uni_gs.w[i] = i < tpp ? host[i].x : 0;
        p->uni_gs.wp[i] = i < tpp ? host[i].y : 0;
    }
    switch (p->logn) {
        case 6: return small_set_attr<6>();
        case 7: return small_set_attr<7>();
        case 8: return small_set_attr<8>();
        case 9: return small_set_attr<9>();
        case 10: return small_set_attr<10>();
        default: return small_set_attr<11>();
    }
}

// kind: 0 = GS, 1 = GS with ans_order store, 2 = GS of d_in (*) d_b scaled by N^-1, 3 = CT
template <int LOGN>
static void small_launch_t(int kind, int grid, cudaStream_t st, const CUtensorMap &lo,
                           const CUtensorMap &hi, const CUtensorMap &blo, const CUtensorMap &bhi,
                           const UniformTw &uni, const SmallParams &prm, bool l4) {
    switch (l4 && kind < 3 ? kind + 4 : kind) {
        case 4:
            fused_gs_small_kernel<LOGN, false, false, true><<<grid, kS_Threads, kS_SmemBytes, st>>>(
                lo, hi, lo, hi, uni, prm);
            break;
        case 5:
            fused_gs_small_kernel<LOGN, true, false, true><<<grid, kS_Threads, kS_SmemBytes, st>>>(
                lo, hi, lo, hi, uni, prm);
            break;
        case 6:
            fused_gs_small_kernel<LOGN, false, true, true><<<grid, kS_Threads, kS_SmemBytes, st>>>(
                lo, hi, blo, bhi, uni, prm);
            break;
        case 0:
            fused_gs_small_kernel<LOGN, false, false><<<grid, kS_Threads, kS_SmemBytes, st>>>(
                lo, hi, lo, hi, uni, prm);
            break;
        case 1:
            fused_gs_small_kernel<LOGN, true, false><<<grid, kS_Threads, kS_SmemBytes, st>>>(
                lo, hi, lo, hi, uni, prm);
            break;
        case 2:
            fused_gs_small_kernel<LOGN, false, true><<<grid, kS_Threads, kS_SmemBytes, st>>>(
                lo, hi, blo, bhi, uni, prm);
            break;
        default: {
            // 4q-lazy forward butterflies: N = 512 / 1024 / 2048 0.72 / 0.64 / 0.59 -> 0.73 / 0.68 / 0.62
            if (l4 && LOGN >= 9) {
                fused_ct_small_kernel<LOGN, true><<<grid, kS_Threads, kS_SmemBytesCt, st>>>(lo, hi, blo, bhi,
                                                                                        uni, prm);
            } else {
                fused_ct_small_kernel<LOGN><<<grid, kS_Threads, kS_SmemBytesCt, st>>>(lo, hi, blo, bhi, uni,
                                                                                  prm);
            }
            break;
        }
    }
}

// Handles the largest multiple of 2048 coefficients; *done_polys tells the caller how
// many polynomials were transformed (the ragged tail goes through the generic pass).
// d_b: second operand (kind 2) -- for kind 3 the output tensor maps are built on d_out.
int launch_small(nttb200_plan *p, int kind, const int32_t *d_in, const int32_t *d_b, int32_t *d_out,
                 size_t batch, cudaStream_t st, size_t *done_polys) {
    *done_polys = 0;
    if (p->logn < 6 || p->logn > 11 || !p->d_tw_r1) return NTTB200_ERR_UNSUPPORTED;
    if (kind == 1 && p->logn != 11) return NTTB200_ERR_UNSUPPORTED;
    if (kind == 2 && !(p->q & 1u)) return NTTB200_ERR_UNSUPPORTED;
    const size_t per_block = (size_t) 2048 >> p->logn;
    const size_t blocks = batch / per_block;
    if (blocks == 0 || blocks > 0x7fffffffull || ((uintptr_t) d_in & 15u) ||
        ((uintptr_t) d_out & 15u) || (d_b && ((uintptr_t) d_b & 15u))) {
        return NTTB200_ERR_UNSUPPORTED;
    }
    CUtensorMap lo, hi, blo, bhi;
    if (encode_tile_map(&lo, d_in, 32, blocks) != NTTB200_OK ||
        encode_tile_map(&hi, d_in + 32, 32, blocks) != NTTB200_OK) {
        return NTTB200_ERR_UNSUPPORTED;
    }
    blo = lo;
    bhi = hi;
    const int32_t *second = kind == 2 ? d_b : (kind == 3 ? d_out : nullptr);
    if (second && (encode_tile_map(&blo, second, 32, blocks) != NTTB200_OK ||
                   encode_tile_map(&bhi, second + 32, 32, blocks) != NTTB200_OK)) {
        return NTTB200_ERR_UNSUPPORTED;
    }
    SmallParams prm;
    prm.out = reinterpret_cast<uint32_t *>(d_out);
    prm.tw_r1 = p->d_tw_r1;
    prm.blocks = blocks;
    prm.q = p->q;
    prm.zero = 0;
    prm.qinv = prm.scale = prm.scale_shoup = 0;
    if (kind == 2) {
        prm.qinv = inv_mod_2_32(p->q);
        uint64_t sc = ((uint64_t) p->n_inv << 32) % p->q;
        prm.scale = (uint32_t) sc;
        prm.scale_shoup = (uint32_t) ((sc << 32) / p->q);
    }
    prm.four_q = 4u * p->q;
    // measured at 2^26 coefficients: N = 2048/1024/512 gain 4/4/2 %, N = 256 loses 6 % (two
    // strided stages cannot amortise the extra canonicalisation step), N <= 128 is HBM-bound
    const bool l4 = use_l4(p) && p->logn >= 9;
    int grid = (int) (blocks < (uint64_t) p->sm_count ? blocks : (uint64_t) p->sm_count);
    switch (p->logn) {
        case 6: small_launch_t<6>(kind, grid, st, lo, hi, blo, bhi, p->uni_gs, prm, l4); break;
        case 7: small_launch_t<7>(kind, grid, st, lo, hi, blo, bhi, p->uni_gs, prm, l4); break;
        case 8: small_launch_t<8>(kind, grid, st, lo, hi, blo, bhi, p->uni_gs, prm, l4); break;
        case 9: small_launch_t<9>(kind, grid, st, lo, hi, blo, bhi, p->uni_gs, prm, l4); break;
        case 10: small_launch_t<10>(kind, grid, st, lo, hi, blo, bhi, p->uni_gs, prm, l4); break;
        default: small_launch_t<11>(kind, grid, st, lo, hi, blo, bhi, p->uni_gs, prm, l4); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    p->last_path = kind == 3 ? "fused_ct_small_warp_tma"
                             : (kind == 2 ? "fused_gs_small_warp_tma_dual" : "fused_gs_small_warp_tma");
    *done_polys = blocks * per_block;
    return NTTB200_OK;
}

int launch_small_gs(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    bool permute_out, cudaStream_t st, size_t *done_polys) {
    return launch_small(p, permute_out ? 1 : 0, d_in, nullptr, d_out, batch, st, done_polys);
}

}  // namespace nttb200
