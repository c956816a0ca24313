// tables.cu -- twiddle tables built ON THE DEVICE.
//
// The reference fills its root table on the host (make_roots, src/test.cpp:27-32) and
// ships it to the accelerator as a buffer object (bo_root, src/test.cpp:137-151).  At
// N = 2^26 that is a 256 MiB table plus 512 MiB of derived layouts per plan; here the
// host ships at most the caller's int32 table (4N bytes) -- or nothing at all for the
// generated families -- and every derived layout is produced by a kernel:
//   shoup_table_kernel     int32 table            -> (w, floor(w*2^32/q)) pairs
//   generate_table_kernel  (kind, base, q, N)     -> the same pairs without any host table
//   tile_table_kernel      pairs                  -> [N/4096][32][65] uint4 tile-pass layout
//   reduce_kernel          arbitrary int32 words  -> canonical residues (the golden's `%`
//                                                    on first touch, src/test.cpp:46-50)
#include <stdlib.h>

#include "plan.h"

namespace nttb200 {

__device__ __forceinline__ uint32_t mulmod_u64(uint32_t a, uint32_t b, uint32_t q) {
    return (uint32_t) (((uint64_t) a * b) % q);
}
__device__ __forceinline__ uint2 shoup_pair(uint32_t w, uint32_t q) {
    return make_uint2(w, (uint32_t) (((uint64_t) w << 32) / q));
}

__global__ void shoup_table_kernel(const int32_t *__restrict__ table, uint2 *__restrict__ tw,
                                   uint32_t n, uint32_t q) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        tw[i] = i == 0 ? make_uint2(0, 0) : shoup_pair((uint32_t) table[i], q);
    }
}

// table[h + i] = gen(h * block_mult + i) for h = 1, 2, 4, .., n/2 and i < h, where
//   kind 0 (powers): gen(e) = base^e            -- make_roots with base = g^((q-1)/N)
//   kind 1 (bitrev): gen(e) = base^bitrev(e)    -- bit reversal over gen_logn bits
// base^x comes from two small power tables: base^x = hi[x >> 13] * lo[x & 8191].
struct GenParams {
    uint32_t n, q, kind, gen_logn, block_mult;
};
constexpr int kGenLoBits = 13;

__global__ void generate_table_kernel(uint2 *__restrict__ tw, const uint32_t *__restrict__ pow_lo,
                                      const uint32_t *__restrict__ pow_hi, const GenParams g) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < g.n; k += gridDim.x * blockDim.x) {
        if (k == 0) {
            tw[0] = make_uint2(0, 0);  // never read (src/test.cpp:45: h >= 1)
            continue;
        }
        const uint32_t h = 1u << (31 - __clz(k));
        uint64_t e = (uint64_t) h * g.block_mult + (k - h);
        if (g.kind == 1) e = __brevll(e) >> (64 - g.gen_logn);
        const uint32_t w = mulmod_u64(pow_hi[e >> kGenLoBits], pow_lo[e & ((1u << kGenLoBits) - 1)], g.q);
        tw[k] = shoup_pair(w, g.q);
    }
}

// [c][slot][j] uint4 of kernels_multi.cu: j < 64 = thread j's pairs of round-1 stage s,
// j == 64 = the tile's round-2 pairs (stage 6+s); slot -> (s, block pair) as in gs_stage_t.
__global__ void tile_table_kernel(const uint2 *__restrict__ tw, uint4 *__restrict__ out,
                                  uint32_t n, uint32_t chunks) {
    const uint64_t total = (uint64_t) chunks * 32 * 65;
    for (uint64_t o = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; o < total;
         o += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t) (o / (32 * 65));
        const uint32_t rem = (uint32_t) (o - (uint64_t) c * (32 * 65));
        const uint32_t slot = rem / 65, j = rem - slot * 65;
        // slots 0-15: stage 0, 16-23: 1, 24-27: 2, 28-29: 3, 30: 4, 31: 5
        const int s = slot < 16 ? 0 : slot < 24 ? 1 : slot < 28 ? 2 : slot < 30 ? 3 : slot < 31 ? 4 : 5;
        const uint32_t blocks = 32u >> s;
        const uint32_t slot0 = 32u - (blocks >= 2 ? blocks : 1u);
        const uint32_t b = (slot - slot0) * 2;
        const size_t base = j < 64 ? (size_t) (n >> (s + 1)) + (size_t) c * (2048u >> s) + (size_t) j * blocks
                                   : (size_t) (n >> (s + 7)) + (size_t) c * blocks;
        const uint2 t0 = tw[base + b];
        const uint2 t1 = blocks >= 2 ? tw[base + b + 1] : make_uint2(0, 0);
        out[o] = make_uint4(t0.x, t0.y, t1.x, t1.y);
    }
}

// x mod q, non-negative, for any int32 x (the golden reduces with `%` at first touch;
// for x >= 0 this is the same residue, negative x is outside the golden's domain and is
// mapped to the mathematical residue)
__global__ void reduce_kernel(const int32_t *in, int32_t *out, size_t count, uint32_t q, int vec) {
    const size_t count4 = vec ? count / 4 : 0;  // unaligned buffers: scalar loop only
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    auto red = [q](int32_t x) -> int32_t {
        int32_t r = x % (int32_t) q;  // q <= 2^30 fits int32
        return r < 0 ? r + (int32_t) q : r;
    };
    for (size_t v = i; v < count4; v += stride) {
        int4 x = reinterpret_cast<const int4 *>(in)[v];
        x.x = red(x.x);
        x.y = red(x.y);
        x.z = red(x.z);
        x.w = red(x.w);
        reinterpret_cast<int4 *>(out)[v] = x;
    }
    for (size_t v = count4 * 4 + i; v < count; v += stride) out[v] = red(in[v]);
}

// out[b][i] = in[b][bitrev_logn(i)]; in place (in == out) as swaps of the pairs i < bitrev(i).
// The standalone form of the bit-reversal adapter for the kernels that do not
// fuse it into their load/store: one extra pass.
__global__ void bitrev_permute_kernel(const int32_t *in, int32_t *out, uint32_t logn, uint64_t total) {
    const uint32_t mask = (1u << logn) - 1u;
    const bool in_place = in == out;
    for (uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t) t & mask;
        const uint32_t r = __brev(i) >> (32 - logn);
        const uint64_t base = t - i;
        if (in_place) {
            if (i < r) {
                const int32_t x = out[base + i], y = out[base + r];
                out[base + i] = y;
                out[base + r] = x;
            }
        } else {
            out[t] = in[base + r];
        }
    }
}

// out[c][r] = in[r][c] for a rows x cols int32 matrix (32 x 32 tiles through shared memory,
// both sides coalesced): batch-major [batch][N] <-> batch-minor [N][batch].
__global__ void transpose_kernel(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                 uint64_t rows, uint64_t cols) {
    __shared__ int32_t tile[32][33];
    const uint64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    for (uint64_t t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
        const uint64_t tr = t / tiles_c, tc = t - tr * tiles_c;
        for (int y = threadIdx.y; y < 32; y += blockDim.y) {
            const uint64_t r = tr * 32 + y, c = tc * 32 + threadIdx.x;
            if (r < rows && c < cols) tile[y][threadIdx.x] = in[r * cols + c];
        }
        __syncthreads();
        for (int y = threadIdx.y; y < 32; y += blockDim.y) {
            const uint64_t c = tc * 32 + y, r = tr * 32 + threadIdx.x;
            if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][y];
        }
        __syncthreads();
    }
}

static int grid_1d(uint64_t items, int sm_count) {
    uint64_t blocks = (items + 255) / 256;
    uint64_t cap = (uint64_t) sm_count * 16;
    return (int) (blocks < cap ? (blocks ? blocks : 1) : cap);
}

int build_shoup_table(nttb200_plan *p, const int32_t *d_table) {
    shoup_table_kernel<<<grid_1d(p->n, p->sm_count), 256>>>(d_table, p->d_tw, p->n, p->q);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    return NTTB200_OK;
}

static uint64_t powmod_host(uint64_t b, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    b %= m;
    while (e) {
        if (e & 1) r = r * b % m;  // m <= 2^30: products fit 64 bits
        b = b * b % m;
        e >>= 1;
    }
    return r;
}

int build_generated_table(nttb200_plan *p, uint32_t kind, uint32_t base, uint32_t gen_logn,
                          uint32_t block_mult) {
    const uint32_t lo_n = 1u << kGenLoBits;
    const uint32_t hi_n = gen_logn > (uint32_t) kGenLoBits ? 1u << (gen_logn - kGenLoBits) : 1u;
    uint32_t *h_pow = nullptr, *d_pow = nullptr;
    h_pow = (uint32_t *) malloc(sizeof(uint32_t) * ((size_t) lo_n + hi_n));
    if (!h_pow) return NTTB200_ERR_ALLOC;
    const uint64_t q = p->q;
    uint64_t cur = 1 % q;
    for (uint32_t i = 0; i < lo_n; i++) {
        h_pow[i] = (uint32_t) cur;
        cur = cur * (base % q) % q;
    }
    const uint64_t step = powmod_host(base, lo_n, q);
    cur = 1 % q;
    for (uint32_t i = 0; i < hi_n; i++) {
        h_pow[lo_n + i] = (uint32_t) cur;
        cur = cur * step % q;
    }
    cudaError_t e = cudaMalloc(&d_pow, sizeof(uint32_t) * ((size_t) lo_n + hi_n));
    if (e == cudaSuccess) {
        e = cudaMemcpy(d_pow, h_pow, sizeof(uint32_t) * ((size_t) lo_n + hi_n), cudaMemcpyHostToDevice);
    }
    free(h_pow);
    if (e != cudaSuccess) {
        if (d_pow) cudaFree(d_pow);
        return cuda_fail(e, "generated table: power tables");
    }
    GenParams g{p->n, p->q, kind, gen_logn, block_mult};
    generate_table_kernel<<<grid_1d(p->n, p->sm_count), 256>>>(p->d_tw, d_pow, d_pow + lo_n, g);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_pow);
    if (e != cudaSuccess) return cuda_fail(e, "generate_table_kernel");
    return NTTB200_OK;
}

int build_tile_table(nttb200_plan *p) {
    const uint32_t chunks = p->n >> 12;
    tile_table_kernel<<<grid_1d((uint64_t) chunks * 32 * 65, p->sm_count), 256>>>(p->d_tw, p->d_tw_tile,
                                                                                  p->n, chunks);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    return NTTB200_OK;
}

int launch_bitrev_permute(nttb200_plan *p, const int32_t *in, int32_t *out, size_t batch,
                          cudaStream_t st) {
    if (batch == 0) return NTTB200_OK;
    const uint64_t total = (uint64_t) batch << p->logn;
    bitrev_permute_kernel<<<grid_1d(total, p->sm_count), 256, 0, st>>>(in, out, p->logn, total);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    return NTTB200_OK;
}

int launch_transpose(nttb200_plan *p, const int32_t *in, int32_t *out, uint64_t rows, uint64_t cols,
                     cudaStream_t st) {
    if (rows == 0 || cols == 0) return NTTB200_OK;
    const uint64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    const uint64_t cap = (uint64_t) p->sm_count * 16;
    transpose_kernel<<<(int) (tiles < cap ? tiles : cap), dim3(32, 8), 0, st>>>(in, out, rows, cols);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    return NTTB200_OK;
}

int launch_reduce(nttb200_plan *p, const int32_t *in, int32_t *out, size_t count, cudaStream_t st) {
    if (count == 0) return NTTB200_OK;
    const int vec = !((uintptr_t) in & 15u) && !((uintptr_t) out & 15u);
    reduce_kernel<<<grid_1d(vec ? count / 4 + 1 : count, p->sm_count), 256, 0, st>>>(in, out, count,
                                                                                     p->q, vec);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    NTTB200_CUDA(cudaGetLastError());
    return NTTB200_OK;
}

}  // namespace nttb200
