// rowstore.cu -- can a thread store its own 256-byte row with 16 STG.128 (thread stride 256 B,
// each warp instruction touches 32 different 128-byte lines, half a sector per thread) fast
// enough to replace the shared-memory transpose + TMA store of the forward (CT) kernels?
// Compared with the coalesced column store of the GS kernels (STG.32, 128 B per warp
// instruction) while the same kernel streams its input with coalesced 128-bit loads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>   // 0: coalesced STG.32 columns, 1: row STG.128, 2: row STG.128 with .cs hint
__global__ void __launch_bounds__(512, 1) k(const uint4 *in, uint32_t *out, size_t polys) {
    const int team = threadIdx.x >> 6, j = threadIdx.x & 63;
    for (size_t p = (size_t) team * gridDim.x + blockIdx.x; p < polys; p += (size_t) gridDim.x * 8) {
        uint32_t v[64];
        const uint4 *src = in + p * 1024;
#pragma unroll
        for (int c = 0; c < 16; c++) {          // coalesced: 16 B per thread, 1 KiB per team instruction
            uint4 x = __ldg(src + c * 64 + j);
            v[4 * c] = x.x + j; v[4 * c + 1] = x.y ^ c; v[4 * c + 2] = x.z + 1; v[4 * c + 3] = x.w;
        }
        uint32_t *dst = out + p * 4096;
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 64; i++) dst[i * 64 + j] = v[i];
        } else {
            uint4 *row = reinterpret_cast<uint4 *>(dst + j * 64);
#pragma unroll
            for (int c = 0; c < 16; c++) {
                uint4 x = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                if (MODE == 2) __stcs(row + c, x); else row[c] = x;
            }
        }
    }
}

template <int MODE>
void run(const char *name, const uint4 *in, uint32_t *out, size_t polys, int sms) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) k<MODE><<<sms, 512>>>(in, out, polys);
    cudaEventRecord(a);
    for (int i = 0; i < 10; i++) k<MODE><<<sms, 512>>>(in, out, polys);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    printf("%-40s %.4f ms per %zu x 16 KiB in + out = %.0f GB/s\n", name, ms, polys, polys * 32768.0 / ms / 1e6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const size_t polys = 65536;
    uint4 *in; uint32_t *out;
    cudaMalloc(&in, polys * 16384); cudaMalloc(&out, polys * 16384);
    cudaMemset(in, 1, polys * 16384);
    run<0>("coalesced STG.32 columns", in, out, polys, p.multiProcessorCount);
    run<1>("row STG.128 (stride 256 B)", in, out, polys, p.multiProcessorCount);
    run<2>("row STG.128 .cs", in, out, polys, p.multiProcessorCount);
    run<0>("coalesced STG.32 columns", in, out, polys, p.multiProcessorCount);
    return 0;
}
