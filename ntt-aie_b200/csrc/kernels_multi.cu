// kernels_multi.cu -- register-radix passes for transforms longer than one 4096 tile.
//
// A length-2^logn golden transform (reference src/test.cpp:34-60) is run as
//   pass 1  "tile pass": stages 0..11 on every contiguous 4096-coefficient tile, the
//           same 64 x 64 TMA-staged team kernel as kernels_fused.cu, except that a
//           tile's twiddles depend on its position c inside the polynomial
//           (table[(N >> (s+1)) + c*(2048 >> s) + ...]) so they are read from a
//           per-position table in global memory laid out [c][slot][65] uint4:
//           coalesced LDG.128 for the thread-private round-1 pairs, one broadcast
//           LDG.128 for the round-2 pairs (L1-resident across a batch);
//   pass 2+ "column pass": the remaining stages s0..s0+K-1 (K <= 6) pair
//           coefficients 2^s apart.  Each thread owns all 2^K partners of VC
//           adjacent columns in registers: 2^K coalesced 32..128-bit loads, K
//           stages, 2^K stores; the twiddle of a butterfly depends only on the row
//           block and the tile's high index bits -- a broadcast load.
// Each pass costs one HBM read + one HBM write (8 bytes per coefficient).  This is
// the GPU analogue of the reference's tile-local stages followed by cross-tile
// stages (src/aie2.py:178-295, src/aie_core.cc:161-187).
#include <stdlib.h>

#include <vector>

#include "fused_common.cuh"
#include "plan.h"

namespace nttb200 {

constexpr int kM_Teams = 8;
constexpr int kM_Threads = kF_Team * kM_Teams;
constexpr int kM_TwRow = 65;                    // 64 round-1 threads + 1 round-2 entry
constexpr int kM_TwTile = 32 * kM_TwRow;        // uint4s of twiddles per tile position
constexpr int kM_SmemBytes = kM_Teams * kF_PolyBytes + 64 + 1024;

__device__ __forceinline__ uint4 ldg128(const uint4 *p) { return __ldg(p); }

// One stage on the thread's 64 registers (pairs i, i + 2^S); the two (w, w') pairs of
// blocks b, b+1 come as one uint4 from tw[slot * 65] (slot = 0,16,24,28,30,31 + b/2).
template <int S, bool REDUCE>
__device__ __forceinline__ void gs_stage_g(uint32_t (&v)[64], const uint4 *tw, uint32_t q,
                                           uint32_t two_q, uint32_t zero) {
    constexpr int kBlocks = 32 >> S;
    constexpr int kSlot0 = 32 - (kBlocks >= 2 ? kBlocks : 1);
    constexpr int kStride = 1 << S;
#pragma unroll
    for (int b = 0; b < kBlocks; b += 2) {
        uint4 t = ldg128(tw + (kSlot0 + b / 2) * kM_TwRow);
#pragma unroll
        for (int e = 0; e < kStride; e++) {
            int i0 = b * 2 * kStride + e;
            gs_bfly<REDUCE>(v[i0], v[i0 + kStride], t.x, t.y, q, two_q, zero);
        }
        if (kBlocks >= 2) {
#pragma unroll
            for (int e = 0; e < kStride; e++) {
                int i0 = (b + 1) * 2 * kStride + e;
                gs_bfly<REDUCE>(v[i0], v[i0 + kStride], t.z, t.w, q, two_q, zero);
            }
        }
    }
}

struct TileParams {
    uint32_t *out;
    const uint4 *tw_tile;  // [chunks][32][65]
    uint32_t batch;
    uint32_t chunks;       // tiles per polynomial
    uint32_t q;
    uint32_t zero;
};

__global__ void __launch_bounds__(kM_Threads, 1)
tile_gs_kernel(const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_hi,
               const TileParams prm) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t data_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = data_base + kM_Teams * kF_PolyBytes;
    const int tid = threadIdx.x;
    const int team = tid >> 6;
    const int j = tid & 63;
    const uint32_t q = prm.q, two_q = 2u * prm.q, zero = prm.zero;

    if (tid < kM_Teams) mbar_init(bar_base + tid * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // Work: tiles in position-major order u = c * batch + poly, so that one CTA stays
    // on (at most two) tile positions and their twiddles stay in L1.  CTA b owns the
    // contiguous range [b*T/G, (b+1)*T/G); its teams stride through it.
    const uint64_t total = (uint64_t) prm.batch * prm.chunks;
    const uint64_t u_begin = total * blockIdx.x / gridDim.x;
    const uint64_t u_end = total * (blockIdx.x + 1) / gridDim.x;
    const uint32_t buf = data_base + team * kF_PolyBytes;
    const uint32_t bar = bar_base + team * 8;
    uint32_t parity = 0;
    uint64_t u = u_begin + team;

    auto tile_of = [&](uint64_t uu, uint32_t &c) -> uint32_t {
        c = (uint32_t) (uu / prm.batch);
        uint32_t poly = (uint32_t) (uu - (uint64_t) c * prm.batch);
        return poly * prm.chunks + c;  // tile index in memory
    };

    uint32_t c_cur = 0, tile_cur = 0;
    if (u < u_end) {
        tile_cur = tile_of(u, c_cur);
        if (j == 0) {
            mbar_expect_tx(bar, kF_PolyBytes);
            tma_load_3d(buf, &map_lo, bar, 0, 0, (int) tile_cur);
            tma_load_3d(buf + kF_PolyBytes / 2, &map_hi, bar, 0, 0, (int) tile_cur);
        }
    }
    const uint32_t r1_row = buf + j * 128;
    const uint32_t r1_xor = (j & 7) << 4;
    const uint32_t r2_col = buf + (j >> 5) * (kF_PolyBytes / 2) + (j & 3) * 4;
    const uint32_t r2_chunk = ((j & 31) >> 2) << 4;

    for (; u < u_end; u += kM_Teams) {
        uint32_t v[64];
        const uint4 *tw = prm.tw_tile + (size_t) c_cur * kM_TwTile;
        mbar_wait(bar, parity);
        parity ^= 1;
#pragma unroll
        for (int c = 0; c < 16; c++) {
            uint4 t = lds128(r1_row + (c >> 3) * (kF_PolyBytes / 2) + (((c & 7) << 4) ^ r1_xor));
            v[4 * c + 0] = t.x;
            v[4 * c + 1] = t.y;
            v[4 * c + 2] = t.z;
            v[4 * c + 3] = t.w;
        }
        const uint4 *tw1 = tw + j;
        gs_stage_g<0, false>(v, tw1, q, two_q, zero);
        gs_stage_g<1, true>(v, tw1, q, two_q, zero);
        gs_stage_g<2, true>(v, tw1, q, two_q, zero);
        gs_stage_g<3, true>(v, tw1, q, two_q, zero);
        gs_stage_g<4, true>(v, tw1, q, two_q, zero);
        gs_stage_g<5, true>(v, tw1, q, two_q, zero);
#pragma unroll
        for (int c = 0; c < 16; c++) {
            sts128(r1_row + (c >> 3) * (kF_PolyBytes / 2) + (((c & 7) << 4) ^ r1_xor), v[4 * c],
                   v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        team_sync(team);
#pragma unroll
        for (int i = 0; i < 64; i++) {
            v[i] = lds32(r2_col + i * 128 + (r2_chunk ^ ((i & 7) << 4)));
        }
        fence_proxy_async();
        team_sync(team);

        const uint32_t tile_store = tile_cur;
        const uint64_t next = u + kM_Teams;
        if (next < u_end) {
            tile_cur = tile_of(next, c_cur);
            if (j == 0) {
                mbar_expect_tx(bar, kF_PolyBytes);
                tma_load_3d(buf, &map_lo, bar, 0, 0, (int) tile_cur);
                tma_load_3d(buf + kF_PolyBytes / 2, &map_hi, bar, 0, 0, (int) tile_cur);
            }
        }
        const uint4 *tw2 = tw + 64;
        gs_stage_g<0, true>(v, tw2, q, two_q, zero);
        gs_stage_g<1, true>(v, tw2, q, two_q, zero);
        gs_stage_g<2, true>(v, tw2, q, two_q, zero);
        gs_stage_g<3, true>(v, tw2, q, two_q, zero);
        gs_stage_g<4, true>(v, tw2, q, two_q, zero);
        gs_stage_g<5, true>(v, tw2, q, two_q, zero);

        uint32_t *dst = prm.out + (size_t) tile_store * 4096 + j;
#pragma unroll
        for (int i = 0; i < 64; i++) dst[i * 64] = min(v[i] - q, v[i]);
    }
}

// ------------------------------------------------------------------ column pass
struct ColParams {
    uint32_t logn;
    uint32_t s0;      // first stage of the pass (stride 2^s0)
    uint32_t q;
    uint32_t zero;
    uint64_t threads; // total logical threads = batch * N / (2^K * VC)
};

template <int VC> struct VecT;
template <> struct VecT<1> { using type = uint32_t; };
template <> struct VecT<2> { using type = uint2; };
template <> struct VecT<4> { using type = uint4; };

template <int K, int VC>
__global__ void __launch_bounds__(256)
column_gs_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                 const uint2 *__restrict__ tw, const ColParams p) {
    constexpr int R = 1 << K;
    using V = typename VecT<VC>::type;
    const uint32_t q = p.q, two_q = 2u * p.q, zero = p.zero;
    const uint32_t n = 1u << p.logn;
    const uint32_t cg_bits = p.s0 - (VC == 4 ? 2 : VC == 2 ? 1 : 0);   // column groups per row
    const uint32_t high_bits = p.logn - p.s0 - K;
    for (uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; t < p.threads;
         t += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t cg = (uint32_t) (t & ((1ull << cg_bits) - 1));
        const uint64_t rest = t >> cg_bits;
        const uint32_t high = (uint32_t) (rest & ((1ull << high_bits) - 1));
        const uint64_t poly = rest >> high_bits;
        const size_t base = poly * n + ((size_t) high << (p.s0 + K)) + (size_t) cg * VC;
        uint32_t v[R][VC];
#pragma unroll
        for (int r = 0; r < R; r++) {
            V x = *reinterpret_cast<const V *>(in + base + ((size_t) r << p.s0));
            const uint32_t *xs = reinterpret_cast<const uint32_t *>(&x);
#pragma unroll
            for (int c = 0; c < VC; c++) v[r][c] = xs[c];
        }
#pragma unroll
        for (int m = 0; m < K; m++) {
            const uint32_t s = p.s0 + m;
            const uint2 *tws = tw + (n >> (s + 1)) + ((size_t) high << (K - m - 1));
#pragma unroll
            for (int b = 0; b < (R >> (m + 1)); b++) {
                const uint2 w = __ldg(tws + b);
#pragma unroll
                for (int e = 0; e < (1 << m); e++) {
                    const int r0 = (b << (m + 1)) + e;
#pragma unroll
                    for (int c = 0; c < VC; c++) {
                        if (m == 0) {
                            gs_bfly<false>(v[r0][c], v[r0 + (1 << m)][c], w.x, w.y, q, two_q, zero);
                        } else {
                            gs_bfly<true>(v[r0][c], v[r0 + (1 << m)][c], w.x, w.y, q, two_q, zero);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            V x;
            uint32_t *xs = reinterpret_cast<uint32_t *>(&x);
#pragma unroll
            for (int c = 0; c < VC; c++) xs[c] = min(v[r][c] - q, v[r][c]);
            *reinterpret_cast<V *>(out + base + ((size_t) r << p.s0)) = x;
        }
    }
}

// --------------------------------------------------------------------- host side
int tile_maps(CUtensorMap *lo, CUtensorMap *hi, const int32_t *base, size_t tiles);  // kernels_fused.cu

int multi_prepare(nttb200_plan *p) {
    if (p->logn < 13 || p->logn > 24) return NTTB200_ERR_UNSUPPORTED;
    const uint32_t chunks = p->n >> 12;
    std::vector<uint2> host(p->n);
    NTTB200_CUDA(cudaMemcpy(host.data(), p->d_tw, sizeof(uint2) * p->n, cudaMemcpyDeviceToHost));
    std::vector<uint4> t((size_t) chunks * kM_TwTile);
    for (uint32_t c = 0; c < chunks; c++) {
        for (int s = 0; s < 6; s++) {
            const int blocks = 32 >> s;
            const int slot0 = 32 - (blocks >= 2 ? blocks : 1);
            for (int j = 0; j <= 64; j++) {
                // j < 64: stage s of round 1, thread j; j == 64: stage 6+s of round 2
                size_t base = j < 64
                                  ? (size_t) (p->n >> (s + 1)) + (size_t) c * (2048 >> s) + (size_t) j * blocks
                                  : (size_t) (p->n >> (s + 7)) + (size_t) c * blocks;
                for (int b = 0; b < blocks; b += 2) {
                    uint2 t0 = host[base + b];
                    uint2 t1 = blocks >= 2 ? host[base + b + 1] : make_uint2(0, 0);
                    t[(size_t) c * kM_TwTile + (size_t) (slot0 + b / 2) * kM_TwRow + j] =
                        make_uint4(t0.x, t0.y, t1.x, t1.y);
                }
            }
        }
    }
    NTTB200_CUDA(cudaMalloc(&p->d_tw_tile, sizeof(uint4) * t.size()));
    NTTB200_CUDA(cudaMemcpy(p->d_tw_tile, t.data(), sizeof(uint4) * t.size(), cudaMemcpyHostToDevice));
    NTTB200_CUDA(cudaFuncSetAttribute(tile_gs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kM_SmemBytes));
    return NTTB200_OK;
}

void multi_release(nttb200_plan *p) {
    if (p->d_tw_tile) cudaFree(p->d_tw_tile);
    p->d_tw_tile = nullptr;
}

template <int K, int VC>
static int launch_column(nttb200_plan *p, const int32_t *in, int32_t *out, size_t batch, int s0,
                         cudaStream_t st) {
    ColParams cp;
    cp.logn = p->logn;
    cp.s0 = (uint32_t) s0;
    cp.q = p->q;
    cp.zero = 0;
    cp.threads = ((uint64_t) batch << p->logn) >> (K + (VC == 4 ? 2 : VC == 2 ? 1 : 0));
    uint64_t blocks = (cp.threads + 255) / 256;
    uint64_t cap = (uint64_t) p->sm_count * 64;
    int grid = (int) (blocks < cap ? blocks : cap);
    column_gs_kernel<K, VC><<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(in),
                                                  reinterpret_cast<uint32_t *>(out), p->d_tw, cp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    return NTTB200_OK;
}

// stages [s0, s0+k) as one column pass; needs s0 >= 2 (128-bit columns) and aligned buffers
int launch_column_pass(nttb200_plan *p, const int32_t *in, int32_t *out, size_t batch, int s0, int k,
                       cudaStream_t st) {
    if (s0 < 2 || k < 1 || k > 6 || s0 + k > (int) p->logn) return NTTB200_ERR_UNSUPPORTED;
    if (((uintptr_t) in & 15u) || ((uintptr_t) out & 15u)) return NTTB200_ERR_UNSUPPORTED;
    switch (k) {
        case 1: return launch_column<1, 4>(p, in, out, batch, s0, st);
        case 2: return launch_column<2, 4>(p, in, out, batch, s0, st);
        case 3: return launch_column<3, 4>(p, in, out, batch, s0, st);
        case 4: return launch_column<4, 4>(p, in, out, batch, s0, st);
        case 5: return launch_column<5, 2>(p, in, out, batch, s0, st);
        default: return launch_column<6, 1>(p, in, out, batch, s0, st);
    }
}

static int launch_multi_gs_once(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                                cudaStream_t st) {
    const uint32_t chunks = p->n >> 12;
    const uint64_t tiles = (uint64_t) batch * chunks;
    CUtensorMap map_lo, map_hi;
    if (tile_maps(&map_lo, &map_hi, d_in, (size_t) tiles) != NTTB200_OK) {
        return NTTB200_ERR_UNSUPPORTED;
    }
    TileParams tp;
    tp.out = reinterpret_cast<uint32_t *>(d_out);
    tp.tw_tile = p->d_tw_tile;
    tp.batch = (uint32_t) batch;
    tp.chunks = chunks;
    tp.q = p->q;
    tp.zero = 0;
    uint64_t ctas = (tiles + kM_Teams - 1) / kM_Teams;
    int grid = (int) (ctas < (uint64_t) p->sm_count ? ctas : (uint64_t) p->sm_count);
    tile_gs_kernel<<<grid, kM_Threads, kM_SmemBytes, st>>>(map_lo, map_hi, tp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    // remaining stages 12..logn-1 in column passes of <= 6 stages, evenly split
    int rest = (int) p->logn - 12;
    int passes = (rest + 5) / 6;
    int s0 = 12;
    for (int k = 0; k < passes; k++) {
        int take = (rest + (passes - k) - 1) / (passes - k);
        int rc = launch_column_pass(p, d_out, d_out, batch, s0, take, st);
        if (rc != NTTB200_OK) return rc;
        s0 += take;
        rest -= take;
    }
    return NTTB200_OK;
}

int launch_multi_gs(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    cudaStream_t st) {
    if (!p->d_tw_tile) return NTTB200_ERR_UNSUPPORTED;
    if (batch == 0) return NTTB200_OK;
    const uint64_t tiles = (uint64_t) batch * (p->n >> 12);
    if (tiles > 0x7fffffffull || batch > 0xffffffffull || ((uintptr_t) d_in & 15u) ||
        ((uintptr_t) d_out & 15u)) {
        return NTTB200_ERR_UNSUPPORTED;
    }
    // L2 blocking: run the passes over sub-batches small enough that what the tile
    // pass writes is still in the 126 MB L2 when the column pass reads it, so the
    // intermediate never costs HBM bandwidth.
    static const long l2_mb = []() {
        const char *e = getenv("NTTB200_L2_CHUNK_MB");
        return e ? atol(e) : 0L;  // measured: separate sub-batch launches lose more than L2 hits win
    }();
    size_t sub = batch;
    if (l2_mb > 0) {
        sub = ((size_t) l2_mb << 20) / ((size_t) p->n * 4);
        if (sub < 1) sub = 1;
    }
    for (size_t b0 = 0; b0 < batch; b0 += sub) {
        size_t nb = batch - b0 < sub ? batch - b0 : sub;
        int rc = launch_multi_gs_once(p, d_in + b0 * p->n, d_out + b0 * p->n, nb, st);
        if (rc != NTTB200_OK) return rc;
    }
    p->last_path = "tile_tma + column_passes";
    return NTTB200_OK;
}

}  // namespace nttb200
