// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (".ts": D[tmem] = A[tmem] . B[smem]) on sm_100a.
//  1. Layout check: A = 128 x 128 BF16 written with tcgen05.st.32x32b (thread = row, register c = columns 2c | 2c+1 of the
//     row, low half = even column), B = 128 x 128 BF16 in shared memory in the canonical no-swizzle K-major UMMA layout
//     (the layout policy_kernel's weights use).  D is read back and compared with a CPU product; the probe reports
//     which packing (even column in the low or in the high half) the tensor core assumes.
//  2. Rate: cycles per 128 x 128 x 128 layer (8 MMAs of K = 16 + commit + mbarrier wait), A from shared memory (SS) vs A
//     from tensor memory (TS), 1 / 2 / 4 independent accumulators issued back to back.
// This decides whether the policy / rollout kernels can keep their activations in TMEM between layers.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_ts_probe umma_ts_probe.cu && ./umma_ts_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int kSlab = 128 * 16;  // one K-chunk of 8 BF16 for 128 rows

// a_packed: [128 rows][64] u32 (two BF16 per word); b_umma: 32 KB already in UMMA layout; d_out: [128][128] f32
__global__ void __launch_bounds__(128, 1) probe(const uint32_t *a_packed, const unsigned char *b_umma, float *d_out, long long *cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 64);
    unsigned char *s_b = smem + 128;             // 32 KB
    unsigned char *s_a = smem + 128 + 32768;     // 32 KB (SS timing only)
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 32768 / 16; i += 128) {
        reinterpret_cast<uint4 *>(s_b)[i] = reinterpret_cast<const uint4 *>(b_umma)[i];
        reinterpret_cast<uint4 *>(s_a)[i] = reinterpret_cast<const uint4 *>(b_umma)[i];  // any data: timing only
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_tmem = tmem;            // columns 0..63: A (BF16 pairs)
    const uint32_t d_tmem = tmem + 64;       // columns 64..191: D
    // ---- A -> TMEM: thread = row, 64 words
    {
        uint32_t v[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = a_packed[tid * 64 + h * 32 + c];
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                ::"r"(a_tmem + lane_base + (uint32_t)h * 32u), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
                "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
                "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
                "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const uint32_t idesc = idesc_bf16(128, 128);
    uint32_t phase = 0;
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < 8; ++j)  // K-step j: 16 BF16 = 8 TMEM columns of A, two 2 KB slabs of B
            mma_ts(d_tmem, a_tmem + (uint32_t)j * 8u, umma_desc(smem_u32(s_b) + (uint32_t)j * 2u * kSlab, kSlab, 128), idesc, j > 0);
        commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(d_tmem + lane_base + (uint32_t)c * 32u) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 32; ++k) d_out[tid * 128 + c * 32 + k] = __uint_as_float(v[k]);
    }
    // ---- rate: `groups` independent accumulators (columns 64 + 128 g ... for g < 3; the 4th reuses A's columns + 448..)
    for (int mode = 0; mode < 2; ++mode) {          // 0 = SS, 1 = TS
        for (int groups = 1; groups <= 3; ++groups) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            const long long t0 = clock64();
            const int reps = 200;
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int r = 0; r < reps; ++r) {
                    for (int g = 0; g < groups; ++g)
                        for (int j = 0; j < 8; ++j) {
                            const uint64_t b = umma_desc(smem_u32(s_b) + (uint32_t)j * 2u * kSlab, kSlab, 128);
                            if (mode == 0) mma_ss(d_tmem + (uint32_t)g * 128u, umma_desc(smem_u32(s_a) + (uint32_t)j * 2u * kSlab, kSlab, 128), b, idesc, j > 0);
                            else mma_ts(d_tmem + (uint32_t)g * 128u, a_tmem + (uint32_t)j * 8u, b, idesc, j > 0);
                        }
                    commit(bar);
                    mbar_wait(bar, phase); phase ^= 1u;
                }
            } else {
                for (int r = 0; r < reps; ++r) { mbar_wait(bar, phase); phase ^= 1u; }
            }
            __syncthreads();
            const long long t1 = clock64();
            if (tid == 0) cyc[mode * 3 + groups - 1] = (t1 - t0) / reps;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    const int M = 128, N = 128, K = 128;
    std::vector<uint16_t> A(M * K), B(N * K);
    srand(1);
    for (auto &x : A) x = f2bf((float)rand() / RAND_MAX - 0.5f);
    for (auto &x : B) x = f2bf((float)rand() / RAND_MAX - 0.5f);
    std::vector<uint32_t> a_packed(M * 64);
    for (int m = 0; m < M; ++m)
        for (int c = 0; c < 64; ++c) a_packed[m * 64 + c] = (uint32_t)A[m * K + 2 * c] | ((uint32_t)A[m * K + 2 * c + 1] << 16);
    // B[n][k] -> UMMA K-major no-swizzle: chunk kc = k / 8 is a slab of N rows x 16 B
    std::vector<unsigned char> b_umma(32768);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) memcpy(&b_umma[(k / 8) * (N * 16) + n * 16 + (k % 8) * 2], &B[n * K + k], 2);
    uint32_t *d_a; unsigned char *d_b; float *d_d; long long *d_c;
    cudaMalloc(&d_a, a_packed.size() * 4); cudaMalloc(&d_b, 32768); cudaMalloc(&d_d, M * N * 4); cudaMalloc(&d_c, 64);
    cudaMemcpy(d_a, a_packed.data(), a_packed.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b_umma.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemset(d_d, 0, M * N * 4);
    const int smem = 128 + 65536;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(d_a, d_b, d_d, d_c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe FAILED: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D(M * N);
    long long cyc[6];
    cudaMemcpy(D.data(), d_d, M * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc, d_c, sizeof cyc, cudaMemcpyDeviceToHost);
    double err_lo = 0, err_hi = 0, mag = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r0 = 0, r1 = 0;
            for (int k = 0; k < K; ++k) {
                r0 += (double)bf2f(A[m * K + k]) * bf2f(B[n * K + k]);        // even column in the low half (as packed)
                r1 += (double)bf2f(A[m * K + (k ^ 1)]) * bf2f(B[n * K + k]);  // the halves swapped
            }
            err_lo = fmax(err_lo, fabs(D[m * N + n] - r0));
            err_hi = fmax(err_hi, fabs(D[m * N + n] - r1));
            mag = fmax(mag, fabs(r0));
        }
    printf("TS layout: max |D - A.B^T| = %.3e with element 2c in the LOW half, %.3e with the halves swapped (|D| up to %.2f)  -> %s\n",
           err_lo, err_hi, mag, err_lo < 1e-3 ? "LOW-half packing confirmed" : (err_hi < 1e-3 ? "HIGH-half packing" : "NEITHER: layout differs"));
    for (int mode = 0; mode < 2; ++mode)
        for (int g = 1; g <= 3; ++g)
            printf("%s  %d accumulator(s) per commit: %6lld cycles per round = %6.1f per 128x128x128 layer\n", mode ? "TS" : "SS", g, cyc[mode * 3 + g - 1],
                   (double)cyc[mode * 3 + g - 1] / g);
    return 0;
}
