// Microbenchmark: issue rate of scalar FFMA (register / constant operand) vs packed fma.rn.f32x2 on sm_100a.
// Decides how the residual-MLP inner product should be written.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float W[64];

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    float x0 = seed * 0.5f, x1 = seed * 0.25f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // scalar FFMA, all-register operands: 16 independent chains x 4
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x0, x1);
        } else if (MODE == 1) {  // scalar FFMA with a constant-bank weight operand
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], W[r * 16 + i], x1);
        } else {  // packed: 8 independent f32x2 chains x 4 (same FMA count as MODE 0 in half the instructions)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long d, A, B, C;
                    asm("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(x0), "f"(x0));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(C) : "f"(x1), "f"(x1));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(A), "l"(B), "l"(C));
                    asm("mov.b64 {%0,%1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
                }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name) {
    float *out;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 16, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * iters * 64;
    printf("%-28s %8.3f ms  %7.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    float w[64]; for (int i = 0; i < 64; ++i) w[i] = 1.0f + i * 1e-6f;
    cudaMemcpyToSymbol(W, w, sizeof w);
    run<0>("FFMA reg,reg,reg");
    run<1>("FFMA reg,const,reg");
    run<2>("FFMA2 (fma.rn.f32x2)");
    return 0;
}
