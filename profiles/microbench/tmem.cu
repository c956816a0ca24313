// tmem.cu -- can Tensor Memory serve as a per-thread table / parking space for the
// NTT kernels on sm_100a?  (1) addressing check: tcgen05.st then tcgen05.ld of a
// lane/column pattern from every warp of a 512-thread CTA; (2) tcgen05.ld throughput
// against LDS.128 for the same bytes, 16 warps per SM, all SMs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem tmem.cu && ./tmem
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                   "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// MODE 0: addressing check.  MODE 1: ld.x8 + wait per step.  MODE 2: ld.x16 + wait.
// MODE 3: ld.x4 + wait.  MODE 4: two LDS.128 per step (same bytes as MODE 1).
// MODE 5: ld.x8 software-pipelined (next load in flight during the "use" of the current one).
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, long long *clk, int *bad) {
    __shared__ uint32_t tmem_base_s;
    __shared__ uint4 lds_tab[32 * 64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 32 * 64; i += 512) lds_tab[i] = make_uint4(i, i + 1, i + 2, i + 3);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base_s;
    // this warp's lane quadrant; warps sharing a quadrant (warp, warp+4, ..) use disjoint columns
    const uint32_t quad = (uint32_t) (warp & 3);
    const uint32_t col0 = (uint32_t) (warp >> 2) * 128u;   // 4 warps per quadrant x 128 columns
    const uint32_t taddr = base + ((quad * 32u) << 16) + col0;

    // fill this warp's 128 columns: value = (global lane << 16) | column
    for (int c = 0; c < 128; c += 8) {
        uint32_t r[8];
#pragma unroll
        for (int e = 0; e < 8; e++) r[e] = ((quad * 32u + lane) << 16) | (col0 + c + e);
        tmem_st8(taddr + c, r);
    }
    tmem_wait_st();

    if (MODE == 0) {
        int errs = 0;
        for (int c = 0; c < 128; c += 8) {
            uint32_t r[8];
            tmem_ld8(taddr + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 8; e++) errs += r[e] != (((quad * 32u + lane) << 16) | (col0 + c + e));
        }
        // a warp sharing the quadrant reads the neighbour warp's columns (cross-warp visibility
        // after a block barrier)
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t ncol0 = (uint32_t) (((warp >> 2) + 1) & 3) * 128u;
        for (int c = 0; c < 128; c += 8) {
            uint32_t r[8];
            tmem_ld8(base + ((quad * 32u) << 16) + ncol0 + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 8; e++) errs += r[e] != (((quad * 32u + lane) << 16) | (ncol0 + c + e));
        }
        if (errs) atomicAdd(bad, errs);
    } else {
        uint32_t acc = 0;
        __syncthreads();
        long long t0 = clock64();
        if (MODE == 5) {
            uint32_t cur[8], nxt[8];
            tmem_ld8(taddr, cur);
            tmem_wait_ld();
#pragma unroll 1
            for (int it = 0; it < ITERS; it += 2) {
                tmem_ld8(taddr + ((it + 1) & 15) * 8, nxt);
#pragma unroll
                for (int e = 0; e < 8; e++) acc = acc * 3 + cur[e];
                tmem_wait_ld();
                tmem_ld8(taddr + ((it + 2) & 15) * 8, cur);
#pragma unroll
                for (int e = 0; e < 8; e++) acc = acc * 3 + nxt[e];
                tmem_wait_ld();
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < ITERS; it++) {
                if (MODE == 1) {
                    uint32_t r[8];
                    tmem_ld8(taddr + (it & 15) * 8, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 8; e++) acc ^= r[e];
                } else if (MODE == 2) {
                    uint32_t r[16];
                    tmem_ld16(taddr + (it & 7) * 16, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; e++) acc ^= r[e];
                } else if (MODE == 3) {
                    uint32_t r[4];
                    tmem_ld4(taddr + (it & 31) * 4, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 4; e++) acc ^= r[e];
                } else if (MODE == 4) {
                    uint4 a = lds_tab[((it * 2) & 31) * 64 + (tid & 63)];
                    uint4 b = lds_tab[((it * 2 + 1) & 31) * 64 + (tid & 63)];
                    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
                }
            }
        }
        long long t1 = clock64();
        if (acc == 0x12345678u) out[0] = acc;
        if (tid == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
    }
}

template <int MODE>
void run(int sms, const char *name, int bytes_per_step) {
    uint32_t *out;
    long long *clk, h = 0;
    int *bad, hb = 0;
    cudaMalloc(&out, 64);
    cudaMalloc(&clk, 8);
    cudaMalloc(&bad, 4);
    cudaMemset(bad, 0, 4);
    k<MODE><<<sms, 512>>>(out, clk, bad);
    k<MODE><<<sms, 512>>>(out, clk, bad);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    if (MODE == 0) {
        printf("%-40s %s, mismatches=%d\n", name, cudaGetErrorString(e), hb);
    } else {
        // 16 warps per SM, each thread bytes_per_step per step
        double cyc_per_step = (double) h / ITERS;
        double b_per_clk_sm = 512.0 * bytes_per_step / cyc_per_step;
        printf("%-40s %s  %.1f clk per step (16 warps)  %.0f B/clk/SM  %.2f clk per warp-instr per SMSP\n",
               name, cudaGetErrorString(e), cyc_per_step, b_per_clk_sm, cyc_per_step / 4.0);
    }
    cudaFree(out);
    cudaFree(clk);
    cudaFree(bad);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    int sms = p.multiProcessorCount;
    run<0>(sms, "addressing: st/ld own + neighbour columns", 0);
    run<1>(sms, "tcgen05.ld 32x32b.x8 + wait", 32);
    run<2>(sms, "tcgen05.ld 32x32b.x16 + wait", 64);
    run<3>(sms, "tcgen05.ld 32x32b.x4 + wait", 16);
    run<4>(sms, "2 x LDS.128 (same bytes as x8)", 32);
    run<5>(sms, "tcgen05.ld x8 software-pipelined", 32);
    return 0;
}
