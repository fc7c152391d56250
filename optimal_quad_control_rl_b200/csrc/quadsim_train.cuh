// quadsim_train.cuh -- sm_100a: the PPO update on the tensor cores (BASELINE config 5).
//
// What it replaces: SB3's `PPO.train()` behind `model.learn` (`3D quad race.ipynb:820`) for the reference's configuration
// (`:784-795`): two MLPs obs -> 120 -> 120 -> 120 -> {4 | 1} with ReLU (policy mean / value), diagonal Gaussian with a free
// log_std, clipped-surrogate loss + vf_coef * value MSE - ent_coef * entropy, minibatch advantage normalisation, global
// gradient-norm clipping, Adam (eps 1e-5).  In torch this is ~60 kernel launches per minibatch (cuBLAS GEMMs whose K is
// the minibatch, elementwise passes, reductions); here it is three:
//
//   ppo_grad_kernel   forward + loss + backward of BOTH networks for one minibatch.  Persistent CTAs, one per SM; even
//                     CTAs own the policy network, odd CTAs the value network.  A CTA walks 128-sample tiles of the
//                     minibatch (rows gathered by index straight from the rollout buffers); per tile
//                       forward   H1 = relu(X W1^T), H2, H3, OUT           4 GEMMs   tcgen05.mma kind::f16, BF16 operands,
//                       loss      dOUT from (OUT, action, old log-prob, advantage, return)      FP32 accumulators in TMEM
//                       backward  dH3 = dOUT W4 ; dZ3 = dH3 * [H3 > 0] ; dH2 = dZ3 W3 ; ...      3 GEMMs
//                       weights   dW4^T += H3^T dOUT ; dW3 += dZ3^T H2 ; dW2 += dZ2^T H1 ; dW1 += dZ1^T X   4 GEMMs
//                     The activations a layer writes (thread = sample, 16-byte chunks of 8 features: the canonical
//                     no-swizzle UMMA layout) are read THREE ways without ever being copied or transposed: as the K-major
//                     A operand of the next forward GEMM, as the MN-major B operand (N = features, K = samples) of the
//                     weight-gradient GEMM, and -- after the in-place mask -- as K-major A (dZ W) and MN-major A (dZ^T H)
//                     of the backward GEMMs.  The forward weights serve the backward pass the same way (MN-major B).
//                     Weight gradients never leave the tensor core between tiles: the four dW accumulators (304 TMEM
//                     columns) accumulate over ALL tiles of the CTA and are read out once per launch into a per-CTA
//                     partial buffer.  Biases ride along as the column that multiplies a constant-1 input
//                     (quadsim_policy.cuh), so bias gradients are column 127 (in_dim for layer 1) of the dW tiles.
//   ppo_reduce_kernel sums the per-CTA partials, applies 1 / sum(weights), accumulates the global squared gradient norm.
//   ppo_adam_kernel   gradient-norm clip + Adam on the float32 master parameters, and re-emits the BF16 UMMA-layout
//                     weight blobs the next minibatch (and the actor, quadsim_policy.cuh) read.
//
// Numerics: BF16 operands (activations, weights, dZ), FP32 accumulation, FP32 loss math and optimizer -- the
// mixed-precision recipe of the torch path's `amp=True`, tested against torch autograd in float32
// (tests/test_gpu_train.py: relative gradient error <= 1e-2, Adam step <= 1e-6).
#pragma once
#include "quadsim_policy.cuh"

namespace qs {

constexpr int kTrW1 = 0;                                  // offsets (floats) into a network's padded parameter block
constexpr int kTrMaxK1 = 64;
__host__ __device__ constexpr int tr_w1_floats(int k1) { return kPolHidden * k1; }           // [out 128][in k1]
__host__ __device__ constexpr int tr_wh_floats() { return kPolHidden * kPolHidden; }        // [out 128][in 128]
__host__ __device__ constexpr int tr_wo_floats() { return kPolHidden * kPolOut; }           // TRANSPOSED [in 128][out 16]
__host__ __device__ constexpr int tr_net_floats(int k1) { return tr_w1_floats(k1) + 2 * tr_wh_floats() + tr_wo_floats(); }
constexpr int kTrStats = 8;  // per-CTA loss statistics: pg_sum, v_sum, clipped, kl_sum, dlogstd[4]

struct TrainParams {
    // minibatch: `rows` sample indices into the flat rollout buffers
    const long long *idx;        // (rows) or NULL = samples 0..rows-1
    long long rows;
    const float *obs;            // (total, in_dim) f32
    const float *act;            // (total, 4) f32  un-clipped sampled actions
    const float *old_logp;       // (total)
    const float *adv;            // (total) raw advantages
    const float *ret;            // (total)
    const float *weight;         // (total) 0/1 sample weights, or NULL
    const double *mb;            // minibatch statistics from ppo_mbstats_kernel: [0] sum w, [1] sum w*adv, [2] sum w*adv^2
    const unsigned char *w_pi;   // BF16 UMMA-layout weight blobs (policy_weight_bytes)
    const unsigned char *w_vf;
    const float *log_std;        // (4)
    float *partial;              // (gridDim.x, tr_net_floats + kTrStats) per-CTA partial gradients + statistics
    int in_dim, k1;
    int normalize_adv;
    float clip_range, vf_coef, obs_limit, act_limit;
    uint32_t weight_bytes;
};

// instruction descriptor with operand major-ness: bit 15 = A is MN-major, bit 16 = B is MN-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_ex(int m, int n, bool a_mn, bool b_mn) {
    return umma_idesc_bf16(m, n) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// dZ = dH * [H > 0] for kSlabs x 8 columns of this thread's row: H (BF16, post-ReLU, >= 0) is read from the thread's own
// 16-byte chunks and overwritten in place with dZ (BF16).  `last_zero`: force column 127 (the constant-1 unit) to 0.
template <int kSlabs>
__device__ __forceinline__ void mask_pack_store(const uint32_t *v, unsigned char *slab0_row, bool last_zero) {
#pragma unroll
    for (int q = 0; q < kSlabs; ++q) {
        const uint4 h = *reinterpret_cast<const uint4 *>(slab0_row + q * kSlab);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lo = (hw[k] & 0x0000FFFFu) ? __uint_as_float(v[q * 8 + 2 * k]) : 0.0f;
            const float hi = (hw[k] & 0xFFFF0000u) ? __uint_as_float(v[q * 8 + 2 * k + 1]) : 0.0f;
            w[k] = pack_bf16(lo, hi);
        }
        if (last_zero && q == kSlabs - 1) w[3] &= 0x0000FFFFu;
        *reinterpret_cast<uint4 *>(slab0_row + q * kSlab) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// The same masking with the result kept in registers (w: kSlabs x 4 words): the stores follow later, once the
// weight-gradient GEMM that still reads H through the async proxy has completed (mask_store_packed).
template <int kSlabs>
__device__ __forceinline__ void mask_pack_regs(const uint32_t *v, const unsigned char *slab0_row, bool last_zero, uint32_t *w) {
#pragma unroll
    for (int q = 0; q < kSlabs; ++q) {
        const uint4 h = *reinterpret_cast<const uint4 *>(slab0_row + q * kSlab);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lo = (hw[k] & 0x0000FFFFu) ? __uint_as_float(v[q * 8 + 2 * k]) : 0.0f;
            const float hi = (hw[k] & 0xFFFF0000u) ? __uint_as_float(v[q * 8 + 2 * k + 1]) : 0.0f;
            w[q * 4 + k] = pack_bf16(lo, hi);
        }
        if (last_zero && q == kSlabs - 1) w[q * 4 + 3] &= 0x0000FFFFu;
    }
}
template <int kSlabs>
__device__ __forceinline__ void mask_store_packed(const uint32_t *w, unsigned char *slab0_row) {
#pragma unroll
    for (int q = 0; q < kSlabs; ++q)
        *reinterpret_cast<uint4 *>(slab0_row + q * kSlab) = make_uint4(w[q * 4], w[q * 4 + 1], w[q * 4 + 2], w[q * 4 + 3]);
}

// dynamic shared memory: [barriers 128 B][weights][X: k1/8 slabs][H1][H2][H3][dOUT: 2 slabs][reduction scratch 64 floats]
__host__ __device__ constexpr size_t train_smem_bytes(int k1) {
    return 128 + policy_weight_bytes(k1, 3) + (size_t)(k1 / 8) * kSlab + 3 * (size_t)(kPolHidden / 8) * kSlab + 2 * kSlab + 256;
}

// TMEM columns: [0,128) working accumulator (Z / dH / OUT), then the weight-gradient accumulators
constexpr uint32_t kTmAcc = 0, kTmDW2 = 128, kTmDW3 = 256, kTmDW1 = 384, kTmDW4 = 384 + kTrMaxK1;  // dW1: k1 (<= 64) cols
constexpr uint32_t kTmCols = 512;

__global__ void __launch_bounds__(kPolRows, 1) ppo_grad_kernel(const __grid_constant__ TrainParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *bar_mma = bar_w + 1;
    uint64_t *bar_dw = bar_w + 2;  // completion of a stage's weight-gradient GEMM (committed after, and apart from, its dH GEMM)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int net = blockIdx.x & 1;                       // 0 = policy network, 1 = value network
    constexpr int kHBytes = (kPolHidden / 8) * kSlab;     // 32 KB
    unsigned char *s_w = smem_raw + 128;
    unsigned char *s_x = s_w + P.weight_bytes;
    unsigned char *s_h1 = s_x + (P.k1 / 8) * kSlab, *s_h2 = s_h1 + kHBytes, *s_h3 = s_h2 + kHBytes;
    unsigned char *s_do = s_h3 + kHBytes;
    float *s_red = reinterpret_cast<float *>(s_do + 2 * kSlab);

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_dw, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, P.weight_bytes);
        bulk_load(s_w, net == 0 ? P.w_pi : P.w_vf, P.weight_bytes, bar_w);
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    mbar_wait(bar_w, 0);

    const uint32_t x_smem = smem_u32(s_x), h1_smem = smem_u32(s_h1), h2_smem = smem_u32(s_h2), h3_smem = smem_u32(s_h3);
    const uint32_t do_smem = smem_u32(s_do);
    const uint32_t w1_smem = smem_u32(s_w), w2_smem = w1_smem + policy_w1_bytes(P.k1), w3_smem = w2_smem + policy_wh_bytes();
    const uint32_t w4_smem = w3_smem + policy_wh_bytes();
    const int k1 = P.k1;

    // minibatch statistics (advantage normalisation over the valid samples, like SB3's per-minibatch normalisation)
    const double wsum = fmax(P.mb[0], 1.0);
    const float adv_mean = P.normalize_adv ? (float)(P.mb[1] / wsum) : 0.0f;
    float adv_rstd = 1.0f;
    if (P.normalize_adv) {
        const double var = fmax(P.mb[2] - P.mb[1] * P.mb[1] / wsum, 0.0) / fmax(wsum - 1.0, 1.0);
        adv_rstd = 1.0f / ((float)sqrt(var) + 1e-8f);
    }
    float std_inv[4], log_std[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { log_std[k] = P.log_std[k]; std_inv[k] = __expf(-log_std[k]); }

    float st_pg = 0.f, st_v = 0.f, st_clip = 0.f, st_kl = 0.f, st_dls[4] = {0.f, 0.f, 0.f, 0.f};
    const long long n_tiles = (P.rows + kPolRows - 1) / kPolRows;
    const long long ctas_per_net = (gridDim.x + 1 - net) / 2;  // even CTAs: ceil(g/2), odd: floor(g/2)
    uint32_t phase = 0;
    bool first_tile = true;

    auto stage_sync_issue = [&](auto &&issue) {
        // generic-proxy smem writes -> async proxy; TMEM reads of the previous epilogue before the MMAs that overwrite it
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue();
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1u;
        tc_fence_after();
    };
    // A backward stage issues TWO groups of MMAs: the dH GEMM the epilogue waits for (bar_mma) and the weight-gradient GEMM,
    // which nobody reads before the end of the launch -- it only has to be finished before the epilogue overwrites its
    // operand H with dZ (bar_dw).  The epilogue's TMEM read-out and masking run while that GEMM is still in the tensor pipe.
    uint32_t phase_dw = 0;
    auto stage_sync_issue2 = [&](auto &&issue_dh, auto &&issue_dw) {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_dh();
            umma_commit(bar_mma);
            issue_dw();
            umma_commit(bar_dw);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1u;
        tc_fence_after();
    };
    // K-major A [128 rows x K] (activations) times K-major B [N x K] (forward weights): D[128 x N]
    auto mma_fwd = [&](uint32_t d, uint32_t a, uint32_t w, int k, int n_rows) {
        const uint32_t idesc = umma_idesc_bf16(kPolRows, n_rows), w_slab = (uint32_t)n_rows * 16u;
        for (int j = 0; j < k / 16; ++j)
            umma_bf16(d, umma_desc(a + (uint32_t)j * 2u * kSlab, kSlab, 128), umma_desc(w + (uint32_t)j * 2u * w_slab, w_slab, 128), idesc, j > 0);
    };
    // backward through a layer: dH[128 x 128] = dZ[128 x K] (K-major A) . W[K(out) x 128(in)]  (forward weights as MN-major B)
    auto mma_bwd = [&](uint32_t d, uint32_t a, uint32_t w, int k, uint32_t w_slab) {
        const uint32_t idesc = umma_idesc_bf16_ex(kPolRows, kPolHidden, false, true);
        for (int j = 0; j < k / 16; ++j)
            umma_bf16(d, umma_desc(a + (uint32_t)j * 2u * kSlab, kSlab, 128), umma_desc(w + (uint32_t)j * 256u, 128, w_slab), idesc, j > 0);
    };
    // weight gradient: dW[M x N] (+)= G^T[M x 128 samples] (MN-major A) . H[128 samples x N] (MN-major B), over the tile
    auto mma_dw = [&](uint32_t d, uint32_t g, uint32_t h, int n_cols, bool acc) {
        const uint32_t idesc = umma_idesc_bf16_ex(kPolRows, n_cols, true, true);
        for (int j = 0; j < kPolRows / 16; ++j)
            umma_bf16(d, umma_desc(g + (uint32_t)j * 256u, 128, kSlab), umma_desc(h + (uint32_t)j * 256u, 128, kSlab), idesc, acc || j > 0);
    };

    // The sample rows are gathered by index from the rollout buffers: random 100-byte reads whose DRAM latency (2 - 3 k
    // cycles) used to stall the top of every tile and its loss stage (ncu: long_scoreboard 5.0 per issue).  The NEXT
    // tile's row and loss inputs are requested while this tile computes and wait in registers (128 threads per SM:
    // registers are free).
    const bool vec_obs = (P.in_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(P.obs) & 15u) == 0;
    float xr[kTrMaxK1];                                   // this thread's observation row of the tile to come
    float4 n_act = make_float4(0.f, 0.f, 0.f, 0.f);
    float n_lp = 0.f, n_adv = 0.f, n_ret = 0.f, n_w = 0.f;
    long long n_s = 0;
    auto fetch_sample = [&](long long t) {
        const long long r = t * kPolRows + tid;
        const bool on = t < n_tiles && r < P.rows;
        n_s = on ? (P.idx ? P.idx[r] : r) : 0;
        const float *row = P.obs + n_s * P.in_dim;
        if (vec_obs) {
#pragma unroll
            for (int q = 0; q < kTrMaxK1 / 4; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (on && 4 * q < P.in_dim) v = *reinterpret_cast<const float4 *>(row + 4 * q);
                xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kTrMaxK1; ++k) xr[k] = (on && k < P.in_dim) ? row[k] : 0.0f;
        }
        n_w = on ? (P.weight ? P.weight[n_s] : 1.0f) : 0.0f;
        if (on) {
            if (net == 0) { n_act = *reinterpret_cast<const float4 *>(P.act + n_s * 4); n_lp = P.old_logp[n_s]; n_adv = P.adv[n_s]; }
            else n_ret = P.ret[n_s];
        }
    };
    fetch_sample(blockIdx.x >> 1);

    for (long long tile = blockIdx.x >> 1; tile < n_tiles; tile += ctas_per_net) {
        const long long r = tile * kPolRows + tid;
        const bool active = r < P.rows;
        // ---- X: this thread's observation row (sanitised like the torch path), BF16, constant 1 in column in_dim
        const float4 a4 = n_act;
        const float s_lp = n_lp, s_adv = n_adv, s_ret = n_ret, w = n_w;
        {
#pragma unroll
            for (int c = 0; c < kTrMaxK1 / 8; ++c) {
                if (c < k1 / 8) {
                    float x[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k = c * 8 + q;
                        float v = xr[k];
                        v = (v == v) ? fminf(fmaxf(v, -P.obs_limit), P.obs_limit) : 0.0f;
                        x[q] = (k == P.in_dim) ? 1.0f : v;
                    }
                    *reinterpret_cast<uint4 *>(s_x + c * kSlab + tid * 16) =
                        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
                }
            }
        }
        fetch_sample(tile + ctas_per_net);  // in flight for the whole tile
        // ---- forward: three hidden layers
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const uint32_t a_in = l == 0 ? x_smem : (l == 1 ? h1_smem : h2_smem);
            const uint32_t w_in = l == 0 ? w1_smem : (l == 1 ? w2_smem : w3_smem);
            stage_sync_issue([&] { mma_fwd(tmem + kTmAcc, a_in, w_in, l == 0 ? k1 : kPolHidden, kPolHidden); });
            uint32_t v0[32], v1[32];
            unsigned char *dst = (l == 0 ? s_h1 : (l == 1 ? s_h2 : s_h3)) + tid * 16;
            tmem_ld32(t_lane + kTmAcc, v0);
            tmem_ld32(t_lane + kTmAcc + 32u, v1);
            tmem_ld_wait();
            relu_pack_store<4>(v0, dst);
            relu_pack_store<4>(v1, dst + 4 * kSlab);
            tmem_ld32(t_lane + kTmAcc + 64u, v0);
            tmem_ld32(t_lane + kTmAcc + 96u, v1);
            tmem_ld_wait();
            relu_pack_store<4>(v0, dst + 8 * kSlab);
            relu_pack_store<4>(v1, dst + 12 * kSlab);
        }
        // ---- output layer + loss gradient (thread = sample)
        stage_sync_issue([&] { mma_fwd(tmem + kTmAcc, h3_smem, w4_smem, kPolHidden, kPolOut); });
        {
            uint32_t v[8];
            tmem_ld8(t_lane + kTmAcc, v);
            tmem_ld_wait();
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (active && w != 0.0f) {
                if (net == 0) {
                    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                    float logp = 0.0f, z[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float ak = (a[k] == a[k]) ? fminf(fmaxf(a[k], -P.act_limit), P.act_limit) : 0.0f;
                        z[k] = (ak - __uint_as_float(v[k])) * std_inv[k];
                        logp += -0.5f * z[k] * z[k] - log_std[k] - 0.9189385332046727f;
                    }
                    float lr = logp - s_lp;
                    lr = (lr == lr) ? lr : 0.0f;
                    const bool lr_in = lr > -20.0f && lr < 20.0f;
                    lr = fminf(fmaxf(lr, -20.0f), 20.0f);
                    const float ratio = __expf(lr);
                    const float ad = (s_adv - adv_mean) * adv_rstd;
                    const float s1 = ad * ratio, s2 = ad * fminf(fmaxf(ratio, 1.0f - P.clip_range), 1.0f + P.clip_range);
                    // d(-min(s1, s2))/d ratio: -adv where the unclipped term is the minimum (ties: both terms carry adv)
                    const float g_lp = (s1 <= s2 && lr_in) ? -ad * ratio * w : 0.0f;
                    // L = -min(s1, s2) * w, so dL/dlogp = g_lp (it carries the minus sign)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        g[k] = g_lp * z[k] * std_inv[k];           // d logp / d mean_k = z_k / std_k
                        st_dls[k] += g_lp * (z[k] * z[k] - 1.0f);  // d logp / d log_std_k = z_k^2 - 1
                    }
                    st_pg += -fminf(s1, s2) * w;
                    st_clip += (fabsf(ratio - 1.0f) > P.clip_range) ? w : 0.0f;
                    st_kl += ((ratio - 1.0f) - lr) * w;
                } else {
                    const float d = __uint_as_float(v[0]) - s_ret;
                    g[0] = 2.0f * P.vf_coef * d * w;
                    st_v += d * d * w;
                }
            }
            *reinterpret_cast<uint4 *>(s_do + tid * 16) = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), 0u, 0u);
            *reinterpret_cast<uint4 *>(s_do + kSlab + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        // ---- backward.  Each stage issues the dH GEMM of a layer together with the weight-gradient GEMM that needs the
        // same operands, then masks dH with [H > 0] in place over H (all readers of H have completed by then).
        const bool acc = !first_tile;
        // stage B3: dH3 = dOUT . W4 ; dW4^T += H3^T . dOUT
        stage_sync_issue2([&] {
            const uint32_t idesc = umma_idesc_bf16_ex(kPolRows, kPolHidden, false, true);
            umma_bf16(tmem + kTmAcc, umma_desc(do_smem, kSlab, 128), umma_desc(w4_smem, 128, kPolOut * 16), idesc, false);
        }, [&] { mma_dw(tmem + kTmDW4, h3_smem, do_smem, kPolOut, acc); });
        auto mask_epilogue = [&](unsigned char *h) {
            uint32_t v0[32], v1[32], w[64];
            unsigned char *dst = h + tid * 16;
            tmem_ld32(t_lane + kTmAcc, v0);
            tmem_ld32(t_lane + kTmAcc + 32u, v1);
            tmem_ld_wait();
            mask_pack_regs<4>(v0, dst, false, w);
            mask_pack_regs<4>(v1, dst + 4 * kSlab, false, w + 16);
            tmem_ld32(t_lane + kTmAcc + 64u, v0);
            tmem_ld32(t_lane + kTmAcc + 96u, v1);
            tmem_ld_wait();
            mask_pack_regs<4>(v0, dst + 8 * kSlab, false, w + 32);
            mask_pack_regs<4>(v1, dst + 12 * kSlab, true, w + 48);   // the constant-1 unit takes no gradient
            mbar_wait(bar_dw, phase_dw);                              // the weight-gradient GEMM has finished reading H
            phase_dw ^= 1u;
            mask_store_packed<16>(w, dst);
        };
        mask_epilogue(s_h3);                                   // H3 <- dZ3
        // stage B2: dH2 = dZ3 . W3 ; dW3 += dZ3^T . H2
        stage_sync_issue2([&] { mma_bwd(tmem + kTmAcc, h3_smem, w3_smem, kPolHidden, kSlab); },
                          [&] { mma_dw(tmem + kTmDW3, h3_smem, h2_smem, kPolHidden, acc); });
        mask_epilogue(s_h2);                                   // H2 <- dZ2
        // stage B1: dH1 = dZ2 . W2 ; dW2 += dZ2^T . H1
        stage_sync_issue2([&] { mma_bwd(tmem + kTmAcc, h2_smem, w2_smem, kPolHidden, kSlab); },
                          [&] { mma_dw(tmem + kTmDW2, h2_smem, h1_smem, kPolHidden, acc); });
        mask_epilogue(s_h1);                                   // H1 <- dZ1
        // stage B0: dW1 += dZ1^T . X  (no gradient flows to the observations).  Waited for here because the next tile
        // overwrites X and H1 while this GEMM would still be reading them.
        stage_sync_issue([&] { mma_dw(tmem + kTmDW1, h1_smem, x_smem, k1, acc); });
        first_tile = false;
    }

    // ---- read-out: the four weight-gradient accumulators -> this CTA's partial buffer (thread = accumulator row)
    float *part = P.partial + (size_t)blockIdx.x * (size_t)(tr_net_floats(k1) + kTrStats);
    if (!first_tile) {
        auto dump = [&](uint32_t col0, int n_cols, float *dst) {  // dst: [128 rows][n_cols]
            float *row = dst + (size_t)tid * n_cols;
            for (int c = 0; c < n_cols; c += 16) {
                uint32_t v[16];
                tmem_ld16(t_lane + col0 + (uint32_t)c, v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; q += 4)
                    *reinterpret_cast<float4 *>(row + c + q) = make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]),
                                                                           __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
            }
        };
        dump(kTmDW1, k1, part);
        dump(kTmDW2, kPolHidden, part + tr_w1_floats(k1));
        dump(kTmDW3, kPolHidden, part + tr_w1_floats(k1) + tr_wh_floats());
        dump(kTmDW4, kPolOut, part + tr_w1_floats(k1) + 2 * tr_wh_floats());
    } else {  // a CTA without a tile contributes zeros
        for (int i = tid; i < tr_net_floats(k1); i += kPolRows) part[i] = 0.0f;
    }
    // loss statistics: block reduction of the per-thread sums
    {
        float vals[kTrStats] = {st_pg, st_v, st_clip, st_kl, st_dls[0], st_dls[1], st_dls[2], st_dls[3]};
#pragma unroll
        for (int k = 0; k < kTrStats; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
            if ((tid & 31) == 0) s_red[warp * kTrStats + k] = vals[k];
        }
        __syncthreads();
        if (tid < kTrStats) part[tr_net_floats(k1) + tid] = s_red[tid] + s_red[kTrStats + tid] + s_red[2 * kTrStats + tid] + s_red[3 * kTrStats + tid];
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmCols) : "memory");
}

// sum w, sum w*adv, sum w*adv^2 over the minibatch (double atomics; rows / 256 CTAs)
__global__ void ppo_mbstats_kernel(const long long *idx, long long rows, const float *adv, const float *weight, double *mb) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double w = 0.0, a = 0.0;
    if (r < rows) {
        const long long s = idx ? idx[r] : r;
        w = weight ? (double)weight[s] : 1.0;
        a = w != 0.0 ? (double)adv[s] : 0.0;  // a masked sample may carry anything
    }
    double v0 = w, v1 = w * a, v2 = w * a * a;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    __shared__ double sh[3][8];
    const int wp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][wp] = v0; sh[1][wp] = v1; sh[2][wp] = v2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[threadIdx.x][k];
        atomicAdd(mb + threadIdx.x, t);
    }
}

struct AdamParams {
    const float *partial;        // (n_ctas, net_floats + kTrStats)
    int n_ctas, k1;
    const double *mb;            // [0] = sum of sample weights of the minibatch
    float *grad;                 // (2 * net_floats + 4): policy net, value net, log_std
    double *norm2;               // [0] accumulated squared gradient norm (zeroed by the host before ppo_reduce_kernel)
    float *stats_out;            // (8) pg_loss, v_loss, clip_frac, approx_kl (minibatch means), grad_norm, 0, 0, 0 -- accumulated
    float *param, *m, *v;        // (2 * net_floats + 4) float32 master parameters and Adam moments
    unsigned char *w_pi, *w_vf;  // BF16 blobs to re-emit
    float lr, beta1, beta2, eps, max_grad_norm, ent_coef;
    float bc1, bc2;              // 1 - beta1^t, 1 - beta2^t
};

// grad[e] = (sum over the CTAs of the element's network) / sum(w); squared norm accumulated.  One thread per element.
__global__ void ppo_reduce_kernel(const __grid_constant__ AdamParams A) {
    const int nf = tr_net_floats(A.k1), stride = nf + kTrStats;
    const int total = 2 * nf + 4 + 4;  // + log_std gradient (4) + the four loss statistics
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const float inv = 1.0f / (float)fmax(A.mb[0], 1.0);
    float g = 0.0f;
    bool is_grad = false;
    if (e < 2 * nf) {  // network parameters: CTAs of parity `net`
        const int net = e / nf, off = e - net * nf;
        for (int c = net; c < A.n_ctas; c += 2) g += A.partial[(size_t)c * stride + off];
        g *= inv;
        A.grad[e] = g;
        is_grad = true;
    } else if (e < 2 * nf + 4) {  // log_std: statistics slots 4..7 of the policy CTAs, minus the entropy bonus
        const int k = e - 2 * nf;
        for (int c = 0; c < A.n_ctas; c += 2) g += A.partial[(size_t)c * stride + nf + 4 + k];
        g = g * inv - A.ent_coef;
        A.grad[e] = g;
        is_grad = true;
    } else if (e < total) {  // loss statistics (means over the minibatch), accumulated over the minibatches of an update
        const int k = e - 2 * nf - 4;
        for (int c = 0; c < A.n_ctas; ++c) g += A.partial[(size_t)c * stride + nf + k];
        atomicAdd(A.stats_out + k, g * inv);
    }
    float g2 = is_grad ? g * g : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = g2;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
        if (t != 0.0f) atomicAdd(A.norm2, (double)t);
    }
}

// element (n = out, k = in) of a [rows x K] weight matrix in the BF16 UMMA K-major blob (pack_policy_weights)
__device__ __forceinline__ void blob_put(unsigned char *blob, size_t layer_off, int rows, int n, int k, float val) {
    const __nv_bfloat16 h = __float2bfloat16_rn(val);
    *reinterpret_cast<__nv_bfloat16 *>(blob + layer_off + (size_t)(k / 8) * rows * 16 + (size_t)n * 16 + (size_t)(k % 8) * 2) = h;
}

// torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (no weight decay, no amsgrad) on every element; the BF16 blobs follow
__global__ void ppo_adam_kernel(const __grid_constant__ AdamParams A) {
    const int nf = tr_net_floats(A.k1);
    const int total = 2 * nf + 4;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const float norm = (float)sqrt(*A.norm2);
    const float coef = fminf(A.max_grad_norm / (norm + 1e-6f), 1.0f);
    if (e == 0) atomicAdd(A.stats_out + 4, norm);
    const float g = A.grad[e] * coef;
    const float m = A.beta1 * A.m[e] + (1.0f - A.beta1) * g;
    const float v = A.beta2 * A.v[e] + (1.0f - A.beta2) * g * g;
    A.m[e] = m; A.v[e] = v;
    const float denom = sqrtf(v) / sqrtf(A.bc2) + A.eps;
    const float p = A.param[e] - (A.lr / A.bc1) * (m / denom);
    A.param[e] = p;
    if (e < 2 * nf) {
        const int net = e / nf;
        int off = e - net * nf;
        unsigned char *blob = net == 0 ? A.w_pi : A.w_vf;
        const int w1 = tr_w1_floats(A.k1), wh = tr_wh_floats();
        if (off < w1) {
            blob_put(blob, 0, kPolHidden, off / A.k1, off % A.k1, p);
        } else if (off < w1 + 2 * wh) {
            const int l = (off - w1) / wh, o2 = (off - w1) - l * wh;
            blob_put(blob, policy_w1_bytes(A.k1) + (size_t)l * policy_wh_bytes(), kPolHidden, o2 / kPolHidden, o2 % kPolHidden, p);
        } else {  // output layer is stored transposed [in][out]
            const int o2 = off - w1 - 2 * wh;
            blob_put(blob, policy_w1_bytes(A.k1) + 2 * (size_t)policy_wh_bytes(), kPolOut, o2 % kPolOut, o2 / kPolOut, p);
        }
    }
}

}  // namespace qs
