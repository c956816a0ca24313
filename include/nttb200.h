/*
 * nttb200.h -- C ABI of the B200-native NTT engine (libnttb200.so).
 *
 * This is the drop-in boundary for the ONE hot path of hal-lab-u-tokyo/ntt-aie:
 * the butterfly-stage network with 32-bit modular multiplication over a
 * precomputed twiddle table.  Plain pointers and sizes only; no C++ types, no
 * torch types, no exceptions cross this boundary.  Every entry point names the
 * reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - element type is int32_t exactly as in the reference (src/test.cpp:34,
 *     src/aie_core.cc:161,189); values are canonical residues in [0, q);
 *   - polynomials are batch-major contiguous: poly b occupies [b*N, (b+1)*N);
 *   - the twiddle table is an INPUT of length N indexed table[h + i]
 *     (src/test.cpp:45): h = number of butterfly blocks of the stage, i = block
 *     index; table[0] is never read.  Any canonical residues are accepted -- the
 *     reference's own natural-power table (src/test.cpp:27-32) reproduces the
 *     reference bit for bit, bit-reversed psi tables make it a real (I)NTT;
 *   - return value 0 = success, non-zero = nttb200_status (the reference's host
 *     returns 0/1 from main, src/test.cpp:162-166,240-247);
 *   - device entry points are asynchronous on the given CUDA stream
 *     (cudaStream_t passed as void*, NULL = default stream); in == out (in place)
 *     is allowed; host entry points are synchronous;
 *   - there is NO CPU fallback: every compute entry point fails with
 *     NTTB200_ERR_CUDA / NTTB200_ERR_NO_DEVICE when no B200 is usable.
 */
#ifndef NTTB200_H
#define NTTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NTTB200_API __attribute__((visibility("default")))
#else
#define NTTB200_API
#endif

typedef enum nttb200_status {
    NTTB200_OK = 0,
    NTTB200_ERR_INVALID_ARG = 1, /* null pointer, logn out of range, bad flags      */
    NTTB200_ERR_MODULUS = 2,     /* q outside [2, 2^31).  q <= 2^30 is the golden's  */
                                 /* own domain (2q-1 <= INT32_MAX, src/test.cpp:     */
                                 /* 48,50) and takes the fast kernels; 2^30 < q <    */
                                 /* 2^31 is served by the canonical stage-pass       */
                                 /* kernels (wider moduli, outside the reference)    */
    NTTB200_ERR_TABLE = 3,       /* table entry outside [0, q)                      */
    NTTB200_ERR_CUDA = 4,        /* a CUDA call failed; see nttb200_last_error()    */
    NTTB200_ERR_NO_DEVICE = 5,   /* no CUDA device / wrong device ordinal           */
    NTTB200_ERR_ALLOC = 6,       /* host or device allocation failed                */
    NTTB200_ERR_UNSUPPORTED = 7  /* valid request this build cannot serve           */
} nttb200_status;

/* plan flags */
#define NTTB200_ORDER_GOLDEN 0u     /* output in the CPU golden's order (src/test.cpp:34-60)    */
#define NTTB200_ORDER_AIE_DEVICE 1u /* output permuted in 16 blocks of N/16 by ans_order, as the */
                                    /* AIE device leaves it (src/test.cpp:69-71,212-219); N>=16  */
#define NTTB200_FORCE_GENERIC 2u    /* debugging: always take the generic stage-pass kernels     */

#define NTTB200_REDUCE_INPUT 4u     /* nttb200_gs_batch / ct_batch accept ANY int32 words and reduce */
                                    /* them mod q first, as the golden's `%` does on first touch     */
                                    /* (src/test.cpp:46-50; e.g. a[i] = i with n > p); costs one     */
                                    /* extra pass.  Without it device inputs must be in [0, q).      */
                                    /* nttb200_gs_host always reduces (the pass hides behind PCIe).  */

#define NTTB200_INPUT_BITREV 8u     /* layout adapters (new): the input is stored in    */
#define NTTB200_OUTPUT_BITREV 16u   /* bit-reversed order / the output is delivered in bit-reversed  */
                                    /* order (index i <-> bitrev_logn(i) inside every polynomial).   */
                                    /* Fused into the load / store of the N = 4096 golden kernel     */
                                    /* (no extra pass); one permutation pass elsewhere.  Not         */
                                    /* combinable with NTTB200_ORDER_AIE_DEVICE.                     */

#define NTTB200_MAX_LOGN 27

/* Opaque plan: owns the device copies of the twiddle table (+ Shoup companions
 * floor(w*2^32/q)) for one (N, q, table).  Replaces the compile-time constants
 * the reference bakes into its device image (N, p, Barrett w/u:
 * src/aie2.py:14-19,178) and the root-table buffer object bo_root
 * (src/test.cpp:119-120,137-144,150). */
typedef struct nttb200_plan nttb200_plan;

/* ---- host-side table generation ------------------------------------------ */

/* Replaces make_roots + modPow (src/test.cpp:15-32) exactly as main() uses them
 * (src/test.cpp:137-139): roots[0] = 1, w = g^((p-1)/n) with INTEGER division,
 * roots[i] = roots[i-1]*w mod p.  64-bit intermediates, so it stays correct for
 * every p <= 2^30 (the reference's int32/uint32 products are only valid for
 * p < 46341). */
NTTB200_API int nttb200_make_roots(int32_t n, int32_t *roots, int32_t p, int32_t g);

/* table[k] = base^bitrev_logn(k) mod p (table[0] = 1): with base = psi (a
 * primitive 2n-th root of unity) this is the forward table of nttb200_ct_batch,
 * with base = psi^-1 it turns the golden network (nttb200_gs_batch) into the
 * un-scaled inverse negacyclic NTT.  New operator, no reference counterpart. */
NTTB200_API int nttb200_make_bitrev_table(int32_t n, int32_t *table, int32_t p, int32_t base);

/* b^e mod m with 64-bit intermediates (helper for psi, n^-1 ...). */
NTTB200_API int32_t nttb200_powmod(int32_t b, int64_t e, int32_t m);

/* ---- plan ------------------------------------------------------------------ */

/* Replaces device image load + bo_root fill/sync (src/test.cpp:110-112,137-151).
 * table_host: N = 2^logn int32 words on the host, index rule above.  1 <= logn <=
 * NTTB200_MAX_LOGN, 2 <= q < 2^31 (see NTTB200_ERR_MODULUS). */
NTTB200_API int nttb200_plan_create(nttb200_plan **plan, int device, uint32_t logn, uint32_t q,
                                    const int32_t *table_host, uint32_t flags);
NTTB200_API int nttb200_plan_destroy(nttb200_plan *plan);

/* Same plan, but the table is GENERATED ON THE DEVICE -- the host ships nothing
 * (the reference builds its table on the host and syncs it to the device,
 * src/test.cpp:27-32,137-151; at N = 2^26 that is a 256 MiB upload per plan):
 *     table[h + i] = gen(h * block_mult + i),   h = 1, 2, .., N/2,  i < h,  table[0] = 1
 *     NTTB200_GEN_POWERS: gen(e) = base^e mod q
 *     NTTB200_GEN_BITREV: gen(e) = base^bitrev(e) mod q, bit reversal over gen_logn bits
 * With gen_logn = logn and block_mult = 1 these are make_roots (base = g^((q-1)/N)) and
 * nttb200_make_bitrev_table.  gen_logn > logn with block_mult = G + r gives rank r's local
 * table of a transform of length 2^gen_logn split over G = 2^(gen_logn-logn) devices
 * (the derived table T_r of the four-step split), block_mult = 1 its cross-device table.
 * Needs (N/2) * (block_mult + 1) <= 2^gen_logn and base < q. */
#define NTTB200_GEN_POWERS 0u
#define NTTB200_GEN_BITREV 1u
NTTB200_API int nttb200_plan_create_generated(nttb200_plan **plan, int device, uint32_t logn,
                                              uint32_t q, uint32_t kind, uint32_t base,
                                              uint32_t gen_logn, uint32_t block_mult,
                                              uint32_t flags);

/* Copies the plan's table (N words, the w values; table[0] reported as 1) to the host:
 * what the reference would have had in bo_root (src/test.cpp:119-120). */
NTTB200_API int nttb200_plan_table(const nttb200_plan *plan, int32_t *table_host);

/* ---- the hot path ---------------------------------------------------------- */

/* Replaces the golden `ntt(a, n, roots_rev, p, stage)` (src/test.cpp:34-60) and
 * the device kernels that implement it (ntt_stage0_to_Nminus5 + ntt_1stage +
 * swap_buff + write_back, src/aie_core.cc:133-361, scheduled by src/aie2.py:
 * 161-317): Gentleman-Sande network, stride 1 -> N/2, twiddle table[h+i].
 * stage_limit mirrors the golden's early exit (src/test.cpp:55-58): stages
 * 0..stage_limit are applied; any value outside [0, logn-2] means full depth
 * (the reference passes n-1, src/test.cpp:67).  d_in/d_out: device pointers,
 * batch*N words. */
NTTB200_API int nttb200_gs_batch(nttb200_plan *plan, const int32_t *d_in, int32_t *d_out,
                                 size_t batch, int stage_limit, void *cuda_stream);

/* Forward partner (new operator): Cooley-Tukey network, stride N/2 -> 1, twiddle
 * table[m+i], butterfly V = a[j+t]*S; a[j] = U+V; a[j+t] = U-V.  Same table
 * index rule, same stage_limit meaning. */
NTTB200_API int nttb200_ct_batch(nttb200_plan *plan, const int32_t *d_in, int32_t *d_out,
                                 size_t batch, int stage_limit, void *cuda_stream);

/* A contiguous range of stages [stage_begin, stage_end) of the GS network on
 * batch polynomials: the building block of the multi-GPU split (the reference
 * splits one transform the same way: tile-local stages, then cross-tile stages,
 * src/aie2.py:178-295).  Stage s has stride 2^s. */
NTTB200_API int nttb200_gs_stage_range(nttb200_plan *plan, const int32_t *d_in, int32_t *d_out,
                                       size_t batch, int stage_begin, int stage_end,
                                       void *cuda_stream);

/* The exchange step of the multi-GPU split fused into the last pass: stages
 * [stage_begin, stage_end = logn) of ONE length-N vector d_buf (this rank's shard or
 * its post-transpose rows), whose results are stored straight into the peers that own
 * them after the transpose -- element idx goes to peer idx >> (logn - log2 world) at
 * offset rank*2^(logn-log2 world) + (idx mod 2^(logn-log2 world)).  peer_bufs[k] is a
 * device pointer to rank k's receive buffer mapped into this process (CUDA IPC /
 * symmetric memory; peer_bufs[rank] is the local one).  The successor of the
 * reference's cross-tile butterflies that write into a neighbour tile's memory
 * (src/aie2.py:184-187,226-229).  The caller synchronises the ranks before the
 * receive buffers are read.  world: power of two <= 16. */
NTTB200_API int nttb200_gs_stage_range_scatter(nttb200_plan *plan, int32_t *d_buf, int stage_begin,
                                               int stage_end, void *const *peer_bufs, int world,
                                               int rank, void *cuda_stream);

/* Host-buffer form of the hot path: what the reference host does around one
 * launch -- sync inputs to the device, run, sync the output back
 * (src/test.cpp:148-151,159-168,181-190).  h_in/h_out are HOST pointers
 * (pinned or pageable), batch*N words.  Copies and kernels are pipelined in
 * chunks over internal streams; returns after h_out is complete. */
NTTB200_API int nttb200_gs_host(nttb200_plan *plan, const int32_t *h_in, int32_t *h_out,
                                size_t batch, int stage_limit);

/* Page-locked host buffers for nttb200_gs_host: the successor of the reference's
 * host-only buffer objects `xrt::bo(device, size, XRT_BO_FLAGS_HOST_ONLY, ...)` +
 * `.map<int32_t*>()` (src/test.cpp:115-134).  write_combined != 0 asks for
 * write-combined memory (fast for the device to read, slow for the CPU to read back:
 * use it for inputs only). */
/* out[b][i] = in[b][bitrev_logn(i)] for batch polynomials (in place allowed): the
 * standalone bit-reversal adapter. */
NTTB200_API int nttb200_bitrev_permute(nttb200_plan *plan, const int32_t *d_in, int32_t *d_out,
                                       size_t batch, void *cuda_stream);

/* Layout adapter: batch-major [batch][N] (this library's layout, the reference's flat buffer
 * object per transform) <-> batch-minor [N][batch] (coefficient i of polynomial b at
 * i*batch + b).  to_batch_minor != 0: d_in is [batch][N], d_out becomes [N][batch]; 0: the
 * reverse.  One transposing pass, out of place (d_in != d_out). */
NTTB200_API int nttb200_transpose(nttb200_plan *plan, const int32_t *d_in, int32_t *d_out,
                                  size_t batch, int to_batch_minor, void *cuda_stream);

/* out[i] = in[i] mod q in [0, q) for ANY int32 words (device pointers): the reduction the
 * golden applies with `%` when it first touches an input (src/test.cpp:46-50). */
NTTB200_API int nttb200_reduce(nttb200_plan *plan, const int32_t *d_in, int32_t *d_out,
                               size_t count, void *cuda_stream);

NTTB200_API int nttb200_host_alloc(void **ptr, size_t bytes, int write_combined);
NTTB200_API int nttb200_host_free(void *ptr);

/* ---- operators that make it a polynomial multiplier (new) ------------------- */

/* c[i] = a[i]*b[i] mod q over count words. */
NTTB200_API int nttb200_pointwise(nttb200_plan *plan, const int32_t *d_a, const int32_t *d_b,
                                  int32_t *d_c, size_t count, void *cuda_stream);

/* c[i] = a[i]*scalar mod q over count words (scalar in [0,q)), e.g. N^-1. */
NTTB200_API int nttb200_scale(nttb200_plan *plan, const int32_t *d_a, int32_t *d_c, size_t count,
                              int32_t scalar, void *cuda_stream);

/* c = a (*) b mod (x^N + 1, q) for batch products: CT(fwd) on a and b,
 * pointwise product, GS(inv), scaled by N^-1.  fwd must hold psi^bitrev, inv
 * must hold psi^-bitrev (nttb200_make_bitrev_table); both plans share N and q.
 * d_c may alias d_a or d_b.  d_a/d_b are not modified unless aliased. */
NTTB200_API int nttb200_polymul_negacyclic(nttb200_plan *fwd, nttb200_plan *inv,
                                           const int32_t *d_a, const int32_t *d_b, int32_t *d_c,
                                           size_t batch, void *cuda_stream);

/* ---- RNS batches (new) ------------------------------------------- */

/* Polynomials in residue-number-system form: L channels ("limbs") modulo L different
 * primes, N = 4096, coefficients laid out [batch][L][4096]; channel l is transformed
 * modulo q[l] with its own table tables_host[l] (4096 words, the golden's table[h+i]
 * rule).  One launch serves all channels: the kernels pick twiddles AND modulus by
 * channel.  1 <= L <= 32. */
typedef struct nttb200_rns_plan nttb200_rns_plan;
NTTB200_API int nttb200_rns_plan_create(nttb200_rns_plan **plan, int device, uint32_t limbs,
                                        const uint32_t *q, const int32_t *const *tables_host,
                                        uint32_t flags);
NTTB200_API int nttb200_rns_plan_destroy(nttb200_rns_plan *plan);
/* golden GS network / CT network on every channel of every polynomial */
NTTB200_API int nttb200_rns_gs_batch(nttb200_rns_plan *plan, const int32_t *d_in, int32_t *d_out,
                                     size_t batch, void *cuda_stream);
NTTB200_API int nttb200_rns_ct_batch(nttb200_rns_plan *plan, const int32_t *d_in, int32_t *d_out,
                                     size_t batch, void *cuda_stream);
/* c = a (*) b mod (x^4096 + 1, q[l]) on every channel; fwd holds psi_l^bitrev tables,
 * inv psi_l^-bitrev tables of the same primes */
NTTB200_API int nttb200_rns_polymul_negacyclic(nttb200_rns_plan *fwd, nttb200_rns_plan *inv,
                                               const int32_t *d_a, const int32_t *d_b, int32_t *d_c,
                                               size_t batch, void *cuda_stream);

/* ---- introspection ---------------------------------------------------------- */
NTTB200_API const char *nttb200_strerror(int status);
/* text of the last CUDA error seen by the calling thread ("" if none) */
NTTB200_API const char *nttb200_last_error(void);
/* how many kernels this library has launched since load (all plans, all threads) */
NTTB200_API uint64_t nttb200_kernel_launches(void);
/* name of the kernel path the last nttb200_gs_batch/ct_batch on this plan took */
NTTB200_API const char *nttb200_plan_last_path(const nttb200_plan *plan);
NTTB200_API uint32_t nttb200_plan_logn(const nttb200_plan *plan);
NTTB200_API uint32_t nttb200_plan_modulus(const nttb200_plan *plan);
NTTB200_API const char *nttb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NTTB200_H */
