// Stage timing of the two policy kernels: clock64 stamps (QS_EXP_TS_TIMING) of CTA 0, first 16 tiles: where do the
// cycles of one layer-stage of a chain go?  policy_kernel_ts (activations in TMEM, 2 chains of 256 threads), threads 0
// (column half 0) and 128 (column half 1) of each chain:
//   0 epilogue done -> 1 chain barrier passed -> 2 MMAs issued + committed (thread 0) -> 3 mbarrier: MMAs complete
//   -> 4 accumulator in registers (tcgen05.ld + wait) -> 5 converted + tcgen05.st issued -> 6 wait::st
// then policy_kernel (operands in shared memory, 4 groups of 128 threads), thread 0 of each group.
// nvcc -O3 -std=c++17 -DQS_EXP_TS_TIMING -gencode arch=compute_100a,code=sm_100a -I../../include -o policy_stages policy_stages.cu
#include <cstdio>
#include <vector>
#include "../../optimal_quad_control_rl_b200/csrc/quadsim_policy.cuh"

int main() {
    const long long n = 1 << 20;
    const int in_dim = 24, k1 = 32, n_hidden = 3;
    qs::PolicyParams P{};
    float *obs, *act; unsigned char *w; unsigned long long *epoch;
    cudaMalloc(&obs, n * in_dim * 4); cudaMemset(obs, 0, n * in_dim * 4);
    cudaMalloc(&act, n * 16);
    const uint32_t wb = qs::policy_weight_bytes(k1, n_hidden);
    cudaMalloc(&w, wb); cudaMemset(w, 0, wb);
    cudaMalloc(&epoch, 16); cudaMemset(epoch, 0, 16);
    P.obs = obs; P.actions = act; P.weights = w; P.epoch = epoch; P.n = n; P.in_dim = in_dim; P.k1 = k1; P.n_hidden = n_hidden;
    P.hidden = 120; P.out_dim = 4; P.deterministic = 1; P.weight_bytes = wb; P.tmem_cols = 512;
    const size_t smem = qs::policy_ts_smem_bytes(k1, n_hidden);
    cudaFuncSetAttribute(qs::policy_kernel_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int r = 0; r < 3; ++r) qs::policy_kernel_ts<<<148, qs::kTsChains * qs::kTsThreads, smem>>>(P);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    qs::policy_kernel_ts<<<148, qs::kTsChains * qs::kTsThreads, smem>>>(P);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel %.1f us (with the stamps)\n", ms * 1e3);
    std::vector<long long> d(8192);
    cudaMemcpyFromSymbol(d.data(), qs::qs_ts_dbg, 8192 * 8);
    auto at = [&](int chain, int half, int it, int layer, int pt) { return d[(((chain * 2 + half) * 16 + it) * 8 + layer) * 8 + pt]; };
    for (int chain = 0; chain < 2; ++chain)
        for (int half = 0; half < 2; ++half) {
            printf("chain %d half %d: mean cycles over tiles 2..13\n  layer:   bar(0>1) issue(1>2) mma(2>3)  ld(3>4) cvt+st(4>5) stwait(5>6) | stage\n", chain, half);
            for (int layer = 0; layer <= n_hidden; ++layer) {
                double s[7] = {0};
                for (int it = 2; it < 14; ++it) {
                    for (int pt = 0; pt < 6; ++pt) s[pt] += (double)(at(chain, half, it, layer, pt + 1) - at(chain, half, it, layer, pt)) / 12;
                    const long long next0 = layer < n_hidden ? at(chain, half, it, layer + 1, 0) : at(chain, half, it + 1, 0, 0);
                    s[6] += (double)(next0 - at(chain, half, it, layer, 0)) / 12;
                }
                if (layer == n_hidden) printf("  %d      %8.0f %8.0f %8.0f  (last layer: action epilogue)             | %6.0f\n", layer, s[0], s[1], s[2], s[6]);
                else printf("  %d      %8.0f %8.0f %8.0f %8.0f %8.0f %8.0f    | %6.0f\n", layer, s[0], s[1], s[2], s[3], s[4], s[5], s[6]);
            }
        }
    // ---- the shared-memory-operand kernel (policy_kernel, 4 tile groups of 128 threads), thread 0 of every group
    //   0 epilogue done -> 1 group barrier passed -> 2 MMAs issued + committed -> 3 mbarrier: complete -> 4 first 64 columns in
    //   registers -> 5 all 128 columns converted and stored to the A slabs
    cudaMemset(epoch, 0, 16);
    {
        const int groups = 4;
        const size_t smem_ss = qs::policy_smem_bytes(k1, n_hidden, groups);
        cudaFuncSetAttribute(qs::policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ss);
        for (int r = 0; r < 3; ++r) qs::policy_kernel<<<148, groups * qs::kPolRows, smem_ss>>>(P);
        cudaEventRecord(e0);
        qs::policy_kernel<<<148, groups * qs::kPolRows, smem_ss>>>(P);
        cudaEventRecord(e1);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("policy_kernel (SS, 4 groups) %.1f us (with the stamps)\n", ms * 1e3);
        cudaMemcpyFromSymbol(d.data(), qs::qs_ts_dbg, 8192 * 8);
        auto ss = [&](int group, int it, int layer, int pt) { return d[((group * 16 + it) * 8 + layer) * 8 + pt]; };
        for (int g = 0; g < groups; ++g) {
            printf("group %d: mean cycles over tiles 2..11\n  layer:   bar(0>1) issue(1>2) mma(2>3) ld64(3>4) cvt+sts+ld64+cvt+sts(4>5) | stage\n", g);
            for (int layer = 0; layer <= n_hidden; ++layer) {
                double s[6] = {0};
                for (int it = 2; it < 12; ++it) {
                    for (int pt = 0; pt < 5; ++pt) s[pt] += (double)(ss(g, it, layer, pt + 1) - ss(g, it, layer, pt)) / 10;
                    const long long next0 = layer < n_hidden ? ss(g, it, layer + 1, 0) : ss(g, it + 1, 0, 0);
                    s[5] += (double)(next0 - ss(g, it, layer, 0)) / 10;
                }
                if (layer == n_hidden) {  // last layer: 3>4 = accumulator read (+ next tile's loads issued), 4>5 = action epilogue, 5>6 = next tile's A operand stored
                    double t6 = 0;
                    for (int it = 2; it < 12; ++it) t6 += (double)(ss(g, it, layer, 6) - ss(g, it, layer, 5)) / 10;
                    printf("  %d      %8.0f %8.0f %8.0f %8.0f %8.0f  next-A %6.0f | %6.0f\n", layer, s[0], s[1], s[2], s[3], s[4], t6, s[5]);
                } else
                printf("  %d      %8.0f %8.0f %8.0f %8.0f %8.0f    | %6.0f\n", layer, s[0], s[1], s[2], s[3], s[4], s[5]);
            }
        }
    }
    return 0;
}
