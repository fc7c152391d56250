// Microbenchmark: TMEM read-out (tcgen05.ld, SASS LDTM) throughput per SM on sm_100a -- the resource that bounds
// policy_kernel / rollout_kernel / ppo_grad_kernel (every hidden layer leaves a 128 x 128 FP32 tile in TMEM).
// One CTA per SM, 512 TMEM columns, `warps` warps (warp w reads lane quadrant w % 4, column window (w / 4) * 128),
// each issuing `iters` x [4 x tcgen05.ld.32x32b.x32 + wait::ld] (= one 128-column accumulator row block per trip).
//   mode 0: .b32                 (FP32 accumulators, what the kernels do today)
//   mode 1: .pack::16b .x32      (two 16-bit columns per register: 64 columns per instruction)
//   mode 2: .16x256b.x8 .b32     (other datapath shape, same bytes)
//   mode 3: .b32 .x64            (fewer, longer instructions)
// Reported: bytes of TMEM columns covered per clock per SM.  Nothing is computed on the values (they are xor-ed into
// a sink so that the loads cannot be dropped).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ldtm_rate ldtm_rate.cu && ./ldtm_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
    "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),       \
    "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),      \
    "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define L32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}"

template <int MODE>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[32]) {
    if (MODE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : R32(v) : "r"(taddr) : "memory");
    if (MODE == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32 ", [%32];" : R32(v) : "r"(taddr) : "memory");
    if (MODE == 2) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " L32 ", [%32];" : R32(v) : "r"(taddr) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned *sink, long long *cycles) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t t = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 3) * 128);
    uint32_t v[32], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 1) {  // 64 columns per instruction: two cover the 128-column window
#pragma unroll
            for (int c = 0; c < 2; ++c) { ld<1>(t + c * 64, v); acc ^= v[0] ^ v[31]; }
        } else if (MODE == 2) {  // 16 lanes x 256 bit x 8: a warp covers its 32 lanes x 32 columns in two halves
#pragma unroll
            for (int c = 0; c < 4; ++c) { ld<2>(t + (uint32_t)((c & 1) * 64) + ((uint32_t)((c >> 1) * 16) << 16), v); acc ^= v[0] ^ v[31]; }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) { ld<0>(t + c * 32, v); acc ^= v[0] ^ v[31]; }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

template <int MODE>
void run(const char *name, int warps, unsigned *sink, long long *cyc_dev) {
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(10, sink, cyc_dev);
    k<MODE><<<148, warps * 32>>>(iters, sink, cyc_dev);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-22s warps %2d  FAILED: %s\n", name, warps, cudaGetErrorString(e)); return; }
    long long c[148];
    cudaMemcpy(c, cyc_dev, sizeof c, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)c[i] / 148;
    // TMEM cells covered: every warp reads 32 lanes x 128 columns x 4 B = 16 KB per trip
    const double bytes = (double)warps * 16384.0 * iters;
    printf("%-22s warps %2d  %10.0f cycles  %7.1f B/clk/SM of TMEM cells  (%.1f cycles per 128x128 FP32 tile)\n", name, warps, avg,
           bytes / avg, avg / iters / (warps / 4.0));
}

int main() {
    unsigned *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 148 * 8);
    for (int w : {4, 8, 16}) {
        run<0>("32x32b.x32 .b32", w, sink, cyc);
        run<1>("32x32b.x32 .pack::16b", w, sink, cyc);
        run<2>("16x256b.x8 .b32", w, sink, cyc);
    }
    return 0;
}
