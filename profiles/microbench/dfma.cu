// dfma.cu -- can the FP64 pipe take the quotient estimate of the Shoup butterfly off the
// FMA-heavy pipe?  h = floor(d * w / q) (or one less) from ONE DFMA.RM:
//   dd   = bitcast(hi = 0x43300000, lo = d)            = 2^52 + d          (no instruction)
//   winv = floor(w/q * 2^52) / 2^52                     (multiple of 2^-52, exact double)
//   C    = 2^52 - 2^52*winv                             (integer, exact double)
//   fma.rm(dd, winv, C) = 2^52 + floor(d * winv)        -> low word = h in {floor(dw/q)-1, floor(dw/q)}
// then y = d*w - h*q in [0, 2q) exactly as with the IMAD.HI estimate.
// Part 1 checks that claim on random and extreme operands; part 2 times butterfly streams.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define ITERS 2048
#define NCH 8

__device__ __forceinline__ uint32_t dfma_quot(uint32_t d, double winv, double c) {
    double dd = __hiloint2double(0x43300000, (int) d);
    return (uint32_t) __double2loint(__fma_rd(dd, winv, c));
}

__global__ void check(uint32_t q, const uint32_t *w, const double *winv, const double *c, int nw,
                      unsigned long long *bad, unsigned long long *minus1) {
    uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t st = lane * 2654435761u + 12345u;
    for (int it = 0; it < 4096; it++) {
        st = st * 1664525u + 1013904223u;
        uint32_t d = (it & 63) == 0 ? 2 * q - 1 : (it & 63) == 1 ? 0 : (uint32_t) (((uint64_t) st * (2ull * q)) >> 32);
        int k = (st >> 7) % nw;
        uint32_t h = dfma_quot(d, winv[k], c[k]);
        uint64_t prod = (uint64_t) d * w[k];
        uint64_t fl = prod / q;
        uint32_t r = (uint32_t) prod - h * q;
        if (!(h == fl || h + 1 == fl) || r >= 2 * q || r % q != prod % q) atomicAdd(bad, 1ull);
        if (h + 1 == fl) atomicAdd(minus1, 1ull);
    }
}

template <int V>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t a, uint32_t b, uint32_t q,
                                          uint32_t zero, double winv, double cc, long long *clk) {
    uint32_t x[NCH], y[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { x[i] = threadIdx.x * 7 + i + a; y[i] = x[i] * 3 + b; }
    const uint32_t two_q = 2 * q;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            uint32_t s = x[i] + y[i] + zero;
            uint32_t d = x[i] - y[i] + two_q;
            s = min(s - two_q, s);
            uint32_t h;
            bool use_d = V == 1 || (V == 2 && (i & 1)) || (V == 3 && (i & 3) == 3) || (V == 4 && (i & 3) != 0);
            if (use_d) h = dfma_quot(d, winv, cc);
            else h = __umulhi(d, b);
            x[i] = s;
            y[i] = d * a - h * q;
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) r ^= x[i] ^ y[i];
    if (r == 0x12345678u) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int V>
void run(int sms, int warps, const char *name) {
    uint32_t *out; long long *clk, h;
    cudaMalloc(&out, 64); cudaMemset(out, 0, 64); cudaMalloc(&clk, 8);
    const uint32_t q = 469762049u, w = 3;
    double winv = (double) ((((unsigned __int128) w) << 52) / q) * (1.0 / 4503599627370496.0);
    double cc = 4503599627370496.0 - winv * 4503599627370496.0;
    k<V><<<sms, warps * 32>>>(out, w, 0x9E3779B9u, q, 0, winv, cc, clk);
    k<V><<<sms, warps * 32>>>(out, w, 0x9E3779B9u, q, 0, winv, cc, clk);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double per = (double) h / ((double) ITERS * NCH * warps / 4.0);
    printf("%-52s warps/SM=%2d  %6.2f SMSP-clk per warp-butterfly\n", name, warps, per);
    cudaFree(out); cudaFree(clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    // ---- part 1: exactness
    for (uint32_t q : {469762049u, 3329u, 1073741789u, 1073479681u, 7681u, 3u}) {
        const int nw = 4096;
        uint32_t *hw = new uint32_t[nw]; double *hi = new double[nw], *hc = new double[nw];
        uint32_t st = 99;
        for (int i = 0; i < nw; i++) {
            st = st * 1664525u + 1013904223u;
            uint32_t w = i == 0 ? 0 : i == 1 ? q - 1 : i == 2 ? 1 : st % q;
            hw[i] = w;
            unsigned __int128 f = (((unsigned __int128) w) << 52) / q;
            hi[i] = (double) (uint64_t) f * (1.0 / 4503599627370496.0);
            hc[i] = 4503599627370496.0 - (double) (uint64_t) f;
        }
        uint32_t *dw; double *di, *dc; unsigned long long *bad, hb[2];
        cudaMalloc(&dw, nw * 4); cudaMalloc(&di, nw * 8); cudaMalloc(&dc, nw * 8); cudaMalloc(&bad, 16);
        cudaMemset(bad, 0, 16);
        cudaMemcpy(dw, hw, nw * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(di, hi, nw * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dc, hc, nw * 8, cudaMemcpyHostToDevice);
        check<<<1024, 256>>>(q, dw, di, dc, nw, bad, bad + 1);
        cudaDeviceSynchronize();
        cudaMemcpy(hb, bad, 16, cudaMemcpyDeviceToHost);
        printf("q=%10u: %llu bad of %llu, quotient one low in %llu\n", q, hb[0], 1024ull * 256 * 4096, hb[1]);
    }
    // ---- part 2: throughput
    for (int w : {16, 32}) {
        run<0>(p.multiProcessorCount, w, "butterfly, IMAD.HI quotient (baseline)");
        run<1>(p.multiProcessorCount, w, "butterfly, DFMA.RM quotient");
        run<2>(p.multiProcessorCount, w, "butterfly, 1/2 DFMA + 1/2 IMAD.HI");
        run<3>(p.multiProcessorCount, w, "butterfly, 1/4 DFMA + 3/4 IMAD.HI");
        run<4>(p.multiProcessorCount, w, "butterfly, 3/4 DFMA + 1/4 IMAD.HI");
    }
    return 0;
}
