// Microbenchmark: throughput of the FP32 -> BF16 pair conversion the tcgen05 epilogues run on every accumulator value
// (cvt.rn[.relu].bf16x2.f32 = SASS F2FP.BF16.F32.PACK_AB) against integer emulations, per SM, on sm_100a.
// 128 x 128 accumulators per tile and layer = 8192 pair conversions: at R pairs / clk / SM the epilogue of one layer needs
// 8192 / R cycles of that pipe (the layer's MMAs need ~520).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cvt_rate cvt_rate.cu && ./cvt_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t *out, int iters, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            uint32_t r;
            if (MODE == 0) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i + 1]), "f"(a[i]));
            if (MODE == 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i + 1]), "f"(a[i]));
            if (MODE == 2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i + 1]), "f"(a[i]));
            if (MODE == 3) {  // integer round-to-nearest-even + ReLU + pack
                uint32_t u0 = __float_as_uint(fmaxf(a[i], 0.f)), u1 = __float_as_uint(fmaxf(a[i + 1], 0.f));
                u0 += 0x7FFFu + ((u0 >> 16) & 1u);
                u1 += 0x7FFFu + ((u1 >> 16) & 1u);
                r = __byte_perm(u0, u1, 0x7632);
            }
            if (MODE == 4) {  // truncation + ReLU + pack (not the product's rounding; a lower bound for any ALU emulation)
                r = __byte_perm(__float_as_uint(fmaxf(a[i], 0.f)), __float_as_uint(fmaxf(a[i + 1], 0.f)), 0x7632);
            }
            acc ^= r;
            a[i] += 1.0f;  // keep the inputs changing (one FADD per pair in every mode)
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int threads) {
    uint32_t *out;
    const int blocks = 148, iters = 8192;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 16, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double pairs = (double)blocks * threads * iters * 8;
    printf("%-34s %2d warps/SM  %8.3f ms  %6.1f pair conversions / clk / SM (at 1.965 GHz)\n", name, threads / 32, ms,
           pairs / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    for (int t : {128, 256, 512}) {
        run<0>("cvt.rn.relu.bf16x2.f32", t);
        run<1>("cvt.rn.bf16x2.f32", t);
        run<2>("cvt.rn.f16x2.f32", t);
        run<3>("int RN-even + relu + prmt", t);
        run<4>("truncate + relu + prmt", t);
    }
    return 0;
}
