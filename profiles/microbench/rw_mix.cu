// Microbenchmark: what DRAM bandwidth does the step kernel's TRAFFIC SHAPE allow on this B200, arithmetic aside?
// The E2E env step moves per env: 108 B in (92 B state block, 68 of which are rewritten in place, + 16 B action) and
// 170 B out (68 B state in place, 96 B observation row, 4 B reward, 1+1 B done / flags): reads : writes = 39 : 61, while
// the roofline denominator (MEASURED_PEAKS.json, torch copy_) is a 50 : 50 copy.  Modes, all fully coalesced float4 planes,
// plain grid-stride loads / stores, no arithmetic, working set >> L2, back-to-back launches like the bench:
//   copy   : 1 plane in, 1 plane out (the denominator's shape, on this harness)
//   step   : 7 planes in (4 of them updated in place), 11 planes out (4 in place + 7 streamed)  = the step's mix
//   read   : 7 planes in, 1 float per warp out
//   write  : 11 planes out
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rw_mix rw_mix.cu && ./rw_mix [envs=1048576] [reps=200]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kIn = 7, kOut = 11, kInPlace = 4;

template <int MODE>
__global__ void __launch_bounds__(256) k(float4 *const *in, float4 *const *out, long long n, float *sink) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (MODE == 0) {
            out[kInPlace][i] = in[kInPlace][i];
        } else if (MODE == 1) {
            float4 v[kIn];
#pragma unroll
            for (int p = 0; p < kIn; ++p) v[p] = in[p][i];
#pragma unroll
            for (int p = 0; p < kOut; ++p) out[p][i] = v[p % kIn];
        } else if (MODE == 2) {
#pragma unroll
            for (int p = 0; p < kIn; ++p) { const float4 v = in[p][i]; acc += v.x + v.y + v.z + v.w; }
        } else {
            const float4 v = make_float4((float)i, 1.f, 2.f, 3.f);
#pragma unroll
            for (int p = 0; p < kOut; ++p) out[p][i] = v;
        }
    }
    if (MODE == 2 && acc == 12345.678f) sink[0] = acc;
}

template <int MODE>
void run(const char *name, float4 **d_in, float4 **d_out, long long n, int reps, double bytes_per_env, float *sink, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int r = 0; r < 10; ++r) k<MODE><<<grid, 256>>>(d_in, d_out, n, sink);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) k<MODE><<<grid, 256>>>(d_in, d_out, n, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / reps;
    printf("%-6s ctas/SM %2d  %9.2f us/launch  %8.1f GB/s  (%.0f B/env)\n", name, ctas_per_sm, us, bytes_per_env * n / us * 1e-3, bytes_per_env);
}

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 1048576;
    const int reps = argc > 2 ? atoi(argv[2]) : 200;
    float4 *h_in[kIn], *h_out[kOut];
    for (int p = 0; p < kOut; ++p) { cudaMalloc(&h_out[p], n * 16); cudaMemset(h_out[p], 0, n * 16); }
    for (int p = 0; p < kIn; ++p) {
        if (p < kInPlace) h_in[p] = h_out[p];  // the state planes are rewritten in place
        else { cudaMalloc(&h_in[p], n * 16); cudaMemset(h_in[p], 0, n * 16); }
    }
    float4 **d_in, **d_out; float *sink;
    cudaMalloc(&d_in, sizeof h_in); cudaMalloc(&d_out, sizeof h_out); cudaMalloc(&sink, 4);
    cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
    cudaMemcpy(d_out, h_out, sizeof h_out, cudaMemcpyHostToDevice);
    printf("envs %lld, %d launches each; planes of %.1f MB\n", n, reps, n * 16 / 1048576.0);
    for (int c : {2, 4, 8}) {
        run<0>("copy", d_in, d_out, n, reps, 32.0, sink, c);
        run<1>("step", d_in, d_out, n, reps, 16.0 * (kIn + kOut), sink, c);
        run<2>("read", d_in, d_out, n, reps, 16.0 * kIn, sink, c);
        run<3>("write", d_in, d_out, n, reps, 16.0 * kOut, sink, c);
    }
    // the copy at the size the driver measured its peak on (1 Gi bf16 = 2 GiB in, 2 GiB out)
    {
        float4 *a, *b;
        const long long big = 134217728;  // float4 elements = 2 GiB
        if (cudaMalloc(&a, big * 16) == cudaSuccess && cudaMalloc(&b, big * 16) == cudaSuccess) {
            float4 *hi[kIn] = {a, a, a, a, a, a, a}, *ho[kOut] = {b, b, b, b, b, b, b, b, b, b, b};
            cudaMemcpy(d_in, hi, sizeof hi, cudaMemcpyHostToDevice);
            cudaMemcpy(d_out, ho, sizeof ho, cudaMemcpyHostToDevice);
            run<0>("copy2G", d_in, d_out, big, 10, 32.0, sink, 8);
        }
    }
    return 0;
}
