// bfly.cu -- which pipe limits the Shoup butterfly on sm_100a?  Variants of the
// butterfly body, 8 independent chains per thread, to separate FMA-heavy-pipe time
// from issue / register-file effects.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define ITERS 2048
#define NCH 8

template <int V>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t a, uint32_t b, uint32_t q,
                                          uint32_t zero, long long *clk) {
    uint32_t x[NCH], y[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { x[i] = threadIdx.x * 7 + i + a; y[i] = x[i] * 3 + b; }
    const uint32_t two_q = 2 * q;
    // per-thread copies the compiler cannot prove uniform (threadIdx.x >> 10 is 0)
    const uint32_t qv = q + out[1 + (threadIdx.x & 1)], two_qv = 2 * qv, zv = zero + out[1 + (threadIdx.x & 3)];
    const uint32_t wv = a + out[2 + (threadIdx.x & 1)], wpv = b + out[4 + (threadIdx.x & 3)];
    const uint32_t neg_q = 0u - q;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (V == 0) {                       // 3 multiplies only
                uint32_t h = __umulhi(x[i], b);
                x[i] = x[i] * a - h * q;
            } else if (V == 1) {                // + sub
                uint32_t d = x[i] - y[i] + two_q;
                uint32_t h = __umulhi(d, b);
                y[i] = d * a - h * q;
            } else if (V == 2) {                // full butterfly, 3-input add
                uint32_t s = x[i] + y[i] + zero;
                uint32_t d = x[i] - y[i] + two_q;
                s = min(s - two_q, s);
                uint32_t h = __umulhi(d, b);
                x[i] = s;
                y[i] = d * a - h * q;
            } else if (V == 6) {                // full butterfly, q / 2q / zero in VECTOR registers
                uint32_t s = x[i] + y[i] + zv;
                uint32_t d = x[i] - y[i] + two_qv;
                s = min(s - two_qv, s);
                uint32_t h = __umulhi(d, b);
                x[i] = s;
                y[i] = d * a - h * qv;
            } else if (V == 7) {                // per-thread (vector register) twiddles, y = d*w - h*q
                uint32_t s = x[i] + y[i] + zero;
                uint32_t d = x[i] - y[i] + two_q;
                s = min(s - two_q, s);
                uint32_t h = __umulhi(d, wpv);
                x[i] = s;
                y[i] = d * wv - h * q;
            } else if (V == 8) {                // per-thread twiddles, y = h*(-q) + d*w
                uint32_t s = x[i] + y[i] + zero;
                uint32_t d = x[i] - y[i] + two_q;
                s = min(s - two_q, s);
                uint32_t h = __umulhi(d, wpv);
                uint32_t t = d * wv;
                x[i] = s;
                y[i] = h * neg_q + t;
            } else if (V == 9) {                // uniform twiddles, y = h*(-q) + d*w
                uint32_t s = x[i] + y[i] + zero;
                uint32_t d = x[i] - y[i] + two_q;
                s = min(s - two_q, s);
                uint32_t h = __umulhi(d, b);
                uint32_t t = d * a;
                x[i] = s;
                y[i] = h * neg_q + t;
            } else if (V == 3) {                // mulhi + one IMAD (h*q fused away)
                uint32_t h = __umulhi(x[i], b);
                x[i] = y[i] - h * q;
            } else if (V == 4) {                // two IMAD lo only
                uint32_t h = x[i] * b;
                x[i] = x[i] * a - h * q + y[i];
            } else if (V == 5) {                // full butterfly with per-thread (register) twiddles
                uint32_t s = x[i] + y[i] + zero;
                uint32_t d = x[i] - y[i] + two_q;
                s = min(s - two_q, s);
                uint32_t h = __umulhi(d, y[(i + 1) % NCH] | 1);
                x[i] = s;
                y[i] = d * (x[(i + 3) % NCH] | 1) - h * q;
            }
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
void run(int sms, int warps, const char *name, int fma_ops) {
    uint32_t *out; long long *clk, h;
    cudaMalloc(&out, 64); cudaMemset(out, 0, 64); cudaMalloc(&clk, 8);
    k<V><<<sms, warps * 32>>>(out, 3, 0x9E3779B9u, 469762049u, 0, clk);
    k<V><<<sms, warps * 32>>>(out, 3, 0x9E3779B9u, 469762049u, 0, clk);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double per = (double) h / ((double) ITERS * NCH * warps / 4.0);  // SMSP clk per warp-body
    printf("%-44s warps/SM=%2d  %6.2f SMSP-clk per warp-body (%d FMA-pipe ops)\n", name, warps, per, fma_ops);
    cudaFree(out); cudaFree(clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int w : {16, 32}) {
        run<0>(p.multiProcessorCount, w, "mulhi + 2 IMAD", 3);
        run<1>(p.multiProcessorCount, w, "sub + mulhi + 2 IMAD", 3);
        run<2>(p.multiProcessorCount, w, "full butterfly (uniform twiddle)", 3);
        run<6>(p.multiProcessorCount, w, "full butterfly, constants in vector regs", 3);
        run<7>(p.multiProcessorCount, w, "vector twiddles, d*w - h*q", 3);
        run<8>(p.multiProcessorCount, w, "vector twiddles, h*(-q) + d*w", 3);
        run<9>(p.multiProcessorCount, w, "uniform twiddles, h*(-q) + d*w", 3);
        run<3>(p.multiProcessorCount, w, "mulhi + 1 IMAD", 2);
        run<4>(p.multiProcessorCount, w, "2 IMAD + add", 2);
    }
    return 0;
}
