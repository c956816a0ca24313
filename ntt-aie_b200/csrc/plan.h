// plan.h -- internal plan object and kernel launch interfaces (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/nttb200.h"

namespace nttb200 {

constexpr int kHostStreams = 4;  // chunk pipeline depth of nttb200_gs_host

// Uniform twiddles of the strided (second) round of the fused kernels: stage s of
// that round uses 2^(R2-1-k) table entries that depend only on the register index,
// never on the thread or the polynomial, so they travel as kernel parameters
// (constant bank operands of IMAD) -- 63 (w, w') pairs at most.
struct UniformTw {
    uint32_t w[64];
    uint32_t wp[64];
};

// Cross-tile twiddles table[1..15] of the one-pass polynomial kernels (stage 12+m, block b
// uses entry (G >> (m+1)) + b): kernel parameters = constant-bank operands.
struct CrossTw {
    uint32_t w[16];
    uint32_t wp[16];
};

}  // namespace nttb200

struct nttb200_plan {
    int device = 0;
    uint32_t logn = 0;
    uint32_t n = 0;
    uint32_t q = 0;
    uint32_t flags = 0;
    uint64_t mu = 0;         // floor(2^62 / q) for barrett_mul
    uint32_t n_inv = 0;      // N^-1 mod q (0 if it does not exist)
    uint32_t n_inv_shoup = 0;
    uint2 *d_tw = nullptr;   // [N] (w, floor(w*2^32/q)), golden index rule table[h+i]
    // fused-kernel twiddle staging (built lazily per kernel family)
    uint4 *d_tw_r1 = nullptr;        // round-1 per-thread twiddles, kernel-private order
    uint4 *d_tw_tile = nullptr;      // [N/4096][32][65] tile-pass twiddles (logn 12..26)
    nttb200::UniformTw uni_gs{};     // round-2 uniform twiddles, GS network
    nttb200::CrossTw cross_tw{};     // table[1..15] (logn >= 13)
    // persistent kernels (kernels_tilecol.cu): a waiter that gives up sets this word (pinned,
    // mapped host memory, allocated on first use); the next launch on the plan reports it
    uint32_t *tc_err_host = nullptr;
    uint32_t *tc_err_dev = nullptr;
    std::mutex tc_mu;                // guards the lazy allocation above (not host_mu: gs_host holds that one)
    int sm_count = 148;
    // written by every launch, possibly from several host threads driving different streams
    std::atomic<const char *> last_path{"none"};

    // nttb200_gs_host resources (lazily created, guarded by host_mu)
    std::mutex host_mu;
    cudaStream_t hstream[nttb200::kHostStreams] = {};
    int32_t *d_stage[nttb200::kHostStreams] = {};
    size_t stage_polys = 0;  // capacity of each staging buffer in polynomials
    bool host_ready = false;
};

// N = 4096, L residue channels with their own moduli and tables
struct nttb200_rns_plan {
    int device = 0;
    uint32_t limbs = 0;
    int sm_count = 148;
    std::vector<nttb200_plan *> sub;   // one validated single-modulus plan per channel
    uint4 *d_tw_tile = nullptr;        // [limbs][32][65]
    std::vector<uint4> pos;            // (q, q^-1 mod 2^32, N^-1*2^32 mod q, its Shoup companion)
};

namespace nttb200 {

extern std::atomic<uint64_t> g_launches;

// error plumbing (api.cu)
int cuda_fail(cudaError_t e, const char *what);
#define NTTB200_CUDA(call)                                        \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return ::nttb200::cuda_fail(e__, #call); \
    } while (0)

// 4q-lazy kernels (fused_common.cuh, gs_bfly_l4 / ct_bfly_l4) serve moduli whose 8q fits a word.
// NTTB200_NO_L4=1 is the A/B switch for measurements.
bool l4_enabled();
inline bool use_l4(const nttb200_plan *p) { return p->q < (1u << 29) && l4_enabled(); }

// Stream-ordered scratch (transform scratch of the multi-kernel products, the counters of the
// persistent kernel) from a library-private memory pool of the CURRENT device that keeps its
// pages across synchronisations (release threshold = max): with the default pool every
// cudaDeviceSynchronize hands the pages back to the driver and the next call pays milliseconds
// to map them again (measured: 8 ms on the first product after a synchronize at N = 2^13).
// Freed with cudaFreeAsync; trimmed when the last plan of the process is destroyed.
int scratch_alloc_async(void **p, size_t bytes, cudaStream_t st);

// device-side table construction and input canonicalisation (tables.cu)
int build_shoup_table(nttb200_plan *p, const int32_t *d_table);
int build_generated_table(nttb200_plan *p, uint32_t kind, uint32_t base, uint32_t gen_logn,
                          uint32_t block_mult);
int build_tile_table(nttb200_plan *p);
int launch_bitrev_permute(nttb200_plan *p, const int32_t *in, int32_t *out, size_t batch,
                          cudaStream_t st);
int launch_transpose(nttb200_plan *p, const int32_t *in, int32_t *out, uint64_t rows, uint64_t cols,
                     cudaStream_t st);
int launch_reduce(nttb200_plan *p, const int32_t *in, int32_t *out, size_t count, cudaStream_t st);

// N = 2^13..2^15 in one pass, private twiddles in tensor memory (kernels_poly.cu)
int polyt_prepare();
int launch_polyt_gs(nttb200_plan *p, const int32_t *d_in, const int32_t *d_b, int32_t *d_out,
                    size_t batch, cudaStream_t st);
int launch_polyt_ct(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    cudaStream_t st);
int launch_polyc_gs(nttb200_plan *p, const int32_t *d_in, const int32_t *d_b, int32_t *d_out,
                    size_t batch, cudaStream_t st);

// N = 2^13..2^16 as one persistent kernel of tile items and column items (kernels_tilecol.cu)
int tilecol_prepare();
int launch_tilecol_gs(nttb200_plan *p, const int32_t *d_in, const int32_t *d_b, int32_t *d_out,
                      size_t batch, cudaStream_t st);

int launch_tilecol_ct(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch, cudaStream_t st);
void tilecol_release(nttb200_plan *p);
int fail_msg(int status, const char *msg);   // sets nttb200_last_error, returns status

// one-kernel negacyclic product, N = 4096 (kernels_polymul.cu)
int polymul_prepare();
int launch_polymul4096(nttb200_plan *fwd, nttb200_plan *inv, const int32_t *d_a, const int32_t *d_b,
                       int32_t *d_c, size_t batch, cudaStream_t st);
int launch_polymul4096_strided(nttb200_plan *fwd, nttb200_plan *inv, const int32_t *d_a,
                               const int32_t *d_b, int32_t *d_c, size_t batch, uint32_t tile_mul,
                               uint32_t tile_off, cudaStream_t st);

// generic stage-pass kernels (kernels_generic.cu): stages [sb, se) of the GS
// (ascending stride) or CT (descending stride) network; permute_out applies the
// ans_order block permutation on the last pass's store.
int generic_prepare();
int launch_generic(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch, int sb,
                   int se, bool ct, bool permute_out, cudaStream_t st);
int launch_pointwise(nttb200_plan *p, const int32_t *a, const int32_t *b, int32_t *c, size_t count,
                     cudaStream_t st);
int launch_scale(nttb200_plan *p, const int32_t *a, int32_t *c, size_t count, uint32_t s,
                 uint32_t s_shoup, cudaStream_t st);

// fused register-radix kernels (kernels_fused.cu).  Return NTTB200_ERR_UNSUPPORTED
// when the (logn, options) combination has no fused kernel so the caller can fall
// back to the generic passes (still CUDA -- never a CPU path).
int fused_prepare(nttb200_plan *p);
void fused_release(nttb200_plan *p);
int launch_fused_gs_bitrev(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                           bool in_br, bool out_br, cudaStream_t st);
int launch_fused_gs(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    bool permute_out, cudaStream_t st);

// warp-per-block kernel for N = 2^9..2^11 (kernels_small.cu)
int small_prepare(nttb200_plan *p);
int launch_small(nttb200_plan *p, int kind, const int32_t *d_in, const int32_t *d_b, int32_t *d_out,
                 size_t batch, cudaStream_t st, size_t *done_polys);
int launch_small_gs(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    bool permute_out, cudaStream_t st, size_t *done_polys);

// tile pass + column passes for logn 12..26 (kernels_multi.cu)
int multi_prepare(nttb200_plan *p);
void multi_release(nttb200_plan *p);
int launch_multi_gs(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    cudaStream_t st);
int launch_column_pass(nttb200_plan *p, const int32_t *in, int32_t *out, size_t batch, int s0, int k,
                       cudaStream_t st);
int launch_gs_range_scatter(nttb200_plan *p, int32_t *d_buf, int sb, int se, void *const *peers,
                            int world, int rank, cudaStream_t st);
int rns_launch(int sm_count, int kind, const uint4 *d_tw_tile, const uint4 *h_pos, uint32_t limbs,
               const int32_t *d_a, const int32_t *d_b, int32_t *d_out, size_t batch,
               cudaStream_t st);
// q^-1 mod 2^32 for odd q (Montgomery products): Newton iteration, 3 correct bits doubling per step
inline uint32_t inv_mod_2_32(uint32_t q) {
    uint32_t x = q;
    for (int i = 0; i < 5; i++) x *= 2u - q * x;
    return x;
}
constexpr int kRnsTwTile = 32 * 65;  // uint4s per channel in d_tw_tile (kernels_multi.cu)
int launch_multi_ct(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                    cudaStream_t st);
int launch_multi_ct_mul(nttb200_plan *p, const int32_t *d_in, const int32_t *d_mul, int32_t *d_out,
                        size_t batch, cudaStream_t st);
int launch_fused_gs_scaled(nttb200_plan *p, const int32_t *d_in, int32_t *d_out, size_t batch,
                           cudaStream_t st);
int launch_multi_gs_dual(nttb200_plan *p, const int32_t *d_a, const int32_t *d_b, int32_t *d_out,
                         size_t batch, cudaStream_t st);

}  // namespace nttb200
