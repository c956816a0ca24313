// pipes.cu -- integer pipe microbenchmark for sm_100a (B200).
// Measures warp-instructions per clock per SM for the instructions the NTT butterfly
// is made of, and the rate of the complete lazy Harvey/Shoup GS butterfly, so the
// kernel design can be checked against the integer-issue bound (SURVEY 7 "hard parts").
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define ILP 8

enum Op { MULHI, MADLO, ADD, MIN, LOP, MADWIDE, SHFL, ADDMIN, BFLY, BFLY_NORED, BFLY_ALU, ADD3, DFMA, OPS };
static const char *names[OPS] = {"mul.hi.u32 (IMAD.HI)", "mad.lo.u32 (IMAD)", "add.u32 (IADD3)",
                                 "min.u32 (VIMNMX)", "xor (LOP3)", "mad.wide.u32 (IMAD.WIDE)",
                                 "shfl.bfly", "add+min pair", "GS butterfly lazy (7 op)",
                                 "GS butterfly no-reduce (5 op)", "GS butterfly, adds forced to ALU",
                                 "3-input add (IADD3)", "fma.rn.f64 (DFMA)"};
static const int ops_per_iter[OPS] = {ILP, ILP, ILP, ILP, ILP, ILP, ILP, 2 * ILP, ILP / 2, ILP / 2, ILP / 2, ILP, ILP / 2};

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t a, uint32_t b, uint32_t q,
                                          long long *clk) {
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 7 + i + a;
    uint64_t acc = 0;
    const uint32_t zero = a - 3;  // opaque runtime zero
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (OP == BFLY_ALU) {
#pragma unroll
            for (int i = 0; i < ILP; i += 2) {
                uint32_t u = x[i], v = x[i + 1];
                uint32_t t = u + v + zero;           // 3-input add: cannot become IMAD.IADD
                uint32_t s = min(t - 2 * q, t);      // VIADDMNMX
                uint32_t d = u - v + 2 * q;
                uint32_t h = __umulhi(d, b);
                x[i] = s;
                x[i + 1] = d * a - h * q;
            }
        } else if (OP == DFMA) {
#pragma unroll
            for (int i = 0; i < ILP; i += 2) {
                double dd = __hiloint2double(x[i + 1], x[i]);
                dd = fma(dd, 1.0000001, 0.5);
                x[i] = __double2loint(dd);
                x[i + 1] = __double2hiint(dd);
            }
        } else if (OP == BFLY || OP == BFLY_NORED) {
#pragma unroll
            for (int i = 0; i < ILP; i += 2) {
                uint32_t u = x[i], v = x[i + 1];
                uint32_t s = u + v;
                if (OP == BFLY) s = min(s, s - 2 * q);
                uint32_t d = u - v + 2 * q;
                uint32_t h = __umulhi(d, b);
                x[i] = s;
                x[i + 1] = d * a - h * q;
            }
        } else {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == MULHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                if (OP == MADLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == ADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                if (OP == MIN) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a + it));
                if (OP == LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(a + it));
                if (OP == MADWIDE) {
                    uint64_t w;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(a));
                    x[i] = (uint32_t) (w >> 32) ^ (uint32_t) w;
                }
                if (OP == ADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == SHFL) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 15));
                if (OP == ADDMIN) {
                    uint32_t t;
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i]), "r"(a));
                    asm volatile("min.u32 %0, %1, %2;" : "+r"(x[i]) : "r"(t), "r"(b));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t r = (uint32_t) acc;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= x[i];
    if (r == 0x12345678u) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int OP>
void run(int sms, int warps_per_sm) {
    uint32_t *out;
    long long *clk, hclk;
    cudaMalloc(&out, 4);
    cudaMalloc(&clk, 8);
    int threads = warps_per_sm * 32;
    int bpsm = 1;
    if (threads > 1024) { bpsm = threads / 1024; threads = 1024; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<sms * bpsm, threads>>>(out, 3, 0x9E3779B9u, 469762049u, clk);
    cudaEventRecord(e0);
    k<OP><<<sms * bpsm, threads>>>(out, 3, 0x9E3779B9u, 469762049u, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&hclk, clk, 8, cudaMemcpyDeviceToHost);
    double warp_instr = (double) ITERS * ops_per_iter[OP] * warps_per_sm;
    printf("%-32s warps/SM=%2d  clk=%9lld  %6.3f warp-ops/clk/SM  (%.1f us, %.0f MHz eff)\n",
           names[OP], warps_per_sm, hclk, warp_instr / (double) hclk, ms * 1e3,
           (double) hclk / (ms * 1e3));
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    int sms = prop.multiProcessorCount;
    for (int w : {8, 16, 32}) {
        run<MULHI>(sms, w);
        run<MADLO>(sms, w);
        run<ADD>(sms, w);
        run<MIN>(sms, w);
        run<LOP>(sms, w);
        run<MADWIDE>(sms, w);
        run<SHFL>(sms, w);
        run<ADDMIN>(sms, w);
        run<BFLY>(sms, w);
        run<BFLY_NORED>(sms, w);
        run<BFLY_ALU>(sms, w);
        run<ADD3>(sms, w);
        run<DFMA>(sms, w);
    }
    return 0;
}
