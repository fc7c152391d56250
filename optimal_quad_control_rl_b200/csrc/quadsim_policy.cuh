// quadsim_policy.cuh -- sm_100a device code of the on-device policy forward.
//
// What it replaces: the reference evaluates its trained controller -- an MLP obs(20+4*ga) -> 120 -> 120 -> 120 -> 4
// with ReLU, Gaussian exploration noise and a clip to [-1,1] -- either through SB3 (`model.predict(env.states)`,
// `3D quad race.ipynb:803`) or through its generated C (`c_code/neural_network.c:397-430` nn_linear / nn_relu /
// nn_forward; noise + clip `c_code/nn_controller.c:158-176`).  Here the same function runs next to the simulator:
// observations never leave the GPU and actions are produced where the step kernel reads them.
//
// Shape of the computation: per 128-env tile, a chain of small GEMMs  [128 x K] . [K x 128]  (K = 32, 128, 128) and
// [128 x 128] . [128 x 16] -- 64 kFLOP per env, two orders of magnitude more arithmetic than the env step, so this
// (and only this) part of the system belongs on the 5th-generation tensor cores:
//   * operands in shared memory in the canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices),
//     BF16; weights are converted and laid out once on the host and arrive as ONE TMA bulk copy per CTA;
//   * tcgen05.mma (kind::f16, M=128, N=128|8, K=16) issued by one elected thread, FP32 accumulators in TMEM;
//   * completion via tcgen05.commit -> mbarrier; epilogue tcgen05.ld 32x32b (thread = row = env), ReLU, BF16
//     re-pack straight into the A-operand slabs of the next layer (in place);
//   * biases are folded into the GEMMs: column `in_dim` of the input is the constant 1, weight row 127 of every
//     hidden layer reproduces it, and the bias sits in the weight column that multiplies it.
// Numerics: BF16 inputs / weights / activations, FP32 accumulation.  The CPU oracle (test infrastructure) restates exactly this
// rounding (qo_policy_forward_bf16); the FP32 reference (the repo's own nn_forward) differs by <= ~1e-2 absolute on
// outputs of magnitude 1, well below the policy's exploration noise (std 0.85-0.90, `nn_controller.c:7-12`).
#pragma once
#include <cuda_bf16.h>

#include "quadsim_kernels.cuh"

namespace qs {

constexpr int kPolRows = 128;     // envs per tile = UMMA M
constexpr int kPolHidden = 128;   // padded hidden width = UMMA N of hidden layers, K of the layers after
constexpr int kPolOut = 16;       // padded output width = UMMA N of the last layer (M=128 needs N % 16 == 0)
constexpr int kPolMaxLayers = 5;  // hidden layers + output layer
constexpr int kPolOnes = kPolHidden - 1;  // hidden unit that carries the constant 1 (bias folding)
constexpr int kSlab = kPolRows * 16;      // one K-chunk of 8 BF16 for 128 rows: 2 KB

struct PolicyParams {
    const float *obs;         // (n, in_dim) f32 row-major; with obs_packed: the step kernel's packed BF16 blocks (pack_block_bytes)
    int obs_packed;           // 1: `obs` already holds the first layer's A operand (k1 <= kPackK): TMA-loaded, nothing converted
    float *actions;           // (n, 4) f32: clip(mean + std * N(0,1), -1, 1)
    float *mean;              // (n, 4) f32 network output before noise, or NULL
    float *raw;               // (n, 4) f32 sampled action BEFORE the clip (what PPO's log-prob is taken of), or NULL
    const unsigned char *weights;  // BF16 blob in UMMA layout (pack_policy_weights in quadsim_capi.cu)
    unsigned long long *epoch;     // [0] forward launches so far (noise key), [1] CTA arrivals
    long long n, env_offset;
    unsigned long long seed;
    int in_dim, k1;           // k1 = K of layer 1 (multiple of 16, > in_dim)
    int n_hidden;             // hidden layers (1..4)
    int hidden;               // real hidden width (<= 127); units hidden..126 are zero padding, unit 127 the constant 1
    int out_dim;              // <= 4
    int deterministic;
    int activation;           // hidden activation: 0 = ReLU (`nn_relu`), 1 = tanh (`nn_tanh`, c_code/neural_network.c:413-417)
    uint32_t weight_bytes;
    uint32_t tmem_cols;       // accumulator columns to allocate: 128 per tile group, rounded up to a power of two
    float std[4];
    float obs_limit;          // > 0: observations are sanitised on the way in like the PPO learner does (NaN -> 0, clamp to +-limit)
};

__host__ __device__ constexpr uint32_t policy_w1_bytes(int k1) { return (uint32_t)(k1 / 8) * kPolHidden * 16; }
__host__ __device__ constexpr uint32_t policy_wh_bytes() { return (kPolHidden / 8) * kPolHidden * 16; }   // 32 KB
__host__ __device__ constexpr uint32_t policy_wo_bytes() { return (kPolHidden / 8) * kPolOut * 16; }      // 4 KB
__host__ __device__ constexpr uint32_t policy_weight_bytes(int k1, int n_hidden) {
    return policy_w1_bytes(k1) + (uint32_t)(n_hidden - 1) * policy_wh_bytes() + policy_wo_bytes();
}
// dynamic shared memory: [mbarriers + tmem slot : 128 B][A operand, 16 slabs, per tile group][weights]
__host__ __device__ constexpr size_t policy_smem_bytes(int k1, int n_hidden, int groups) {
    return 128 + (size_t)groups * (kPolHidden / 8) * kSlab + policy_weight_bytes(k1, n_hidden);
}

// shared-memory matrix descriptor, canonical K-major layout without swizzle: 8 rows x 16 B core matrices,
// `lbo` = byte distance between the two K-adjacent core matrices of one MMA, `sbo` = between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);  // bits 46-47: descriptor version 1 (sm_100)
}
// instruction descriptor: D=F32, A=B=BF16, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of FP32 accumulators: thread i of the warp receives row (lane base + i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {  // round-to-nearest-even, lo in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// One thread issues the K/16 MMAs of a layer: D[128 x N] (+)= A[128 x K] . W[N x K]^T
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_smem, uint32_t w_smem, int k, int n_rows,
                                            uint64_t *bar) {
    const uint32_t idesc = umma_idesc_bf16(kPolRows, n_rows);
    const uint32_t w_slab = (uint32_t)n_rows * 16u;
    for (int j = 0; j < k / 16; ++j) {
        const uint64_t a = umma_desc(a_smem + (uint32_t)j * 2u * kSlab, kSlab, 128);
        const uint64_t b = umma_desc(w_smem + (uint32_t)j * 2u * w_slab, w_slab, 128);
        umma_bf16(d_tmem, a, b, idesc, j > 0);
    }
    umma_commit(bar);  // implies tcgen05.fence::before_thread_sync
}

__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {  // max(x, 0) fused into the conversion
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// kSlabs x 8 accumulator columns of this thread's row -> ReLU -> BF16 -> the next layer's A slabs (16 B per slab and row)
template <int kSlabs>
__device__ __forceinline__ void relu_pack_store(const uint32_t *v, unsigned char *slab0_row) {
#pragma unroll
    for (int q = 0; q < kSlabs; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
            w[h] = pack_relu_bf16(__uint_as_float(v[q * 8 + 2 * h]), __uint_as_float(v[q * 8 + 2 * h + 1]));
        *reinterpret_cast<uint4 *>(slab0_row + q * kSlab) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
// tanh hidden activation (SB3's default `activation_fn`; the reference's generated C has `nn_tanh` next to `nn_relu`):
// MUFU.TANH (max rel. error ~2^-11, below the BF16 rounding that follows)
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int kSlabs>
__device__ __forceinline__ void tanh_pack_store(const uint32_t *v, unsigned char *slab0_row) {
#pragma unroll
    for (int q = 0; q < kSlabs; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
            w[h] = pack_bf16(tanh_fast(__uint_as_float(v[q * 8 + 2 * h])), tanh_fast(__uint_as_float(v[q * 8 + 2 * h + 1])));
        *reinterpret_cast<uint4 *>(slab0_row + q * kSlab) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
template <int kSlabs>
__device__ __forceinline__ void act_pack_store(const uint32_t *v, unsigned char *slab0_row, int act) {
    if (act == 0) relu_pack_store<kSlabs>(v, slab0_row);
    else tanh_pack_store<kSlabs>(v, slab0_row);
}
// Columns 96..127 of a hidden layer.  With hidden <= 120 the last slab (units 120..126 = zero padding, unit 127 = the
// constant 1 of the bias folding) never changes: it is written once per kernel (init_const_slab) and neither read
// from TMEM nor stored again -- 6 % less of the TMEM read-out that bounds this kernel.
__device__ __forceinline__ void epilogue_tail(uint32_t t_lane, unsigned char *s_a_row, bool const_last, uint32_t *v, int act) {
    if (const_last) {
        tmem_ld16(t_lane + 96u, v);
        tmem_ld8p(t_lane + 112u, v + 16);
        tmem_ld_wait();
        act_pack_store<3>(v, s_a_row + 12 * kSlab, act);
    } else {
        tmem_ld32(t_lane + 96u, *reinterpret_cast<uint32_t (*)[32]>(v));
        tmem_ld_wait();
        act_pack_store<4>(v, s_a_row + 12 * kSlab, act);
    }
}
__device__ __forceinline__ void init_const_slab(unsigned char *s_a_row) {  // {0 x 7, 1.0} in BF16
    *reinterpret_cast<uint4 *>(s_a_row + 15 * kSlab) = make_uint4(0u, 0u, 0u, 0x3F800000u);
}
__device__ __forceinline__ void group_barrier(int group) {  // the 128 threads of one tile group
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

// mean + std * N(0,1) (`c_code/nn_controller.c:158-169`): Box-Muller on ONE Philox block keyed by (seed, global env,
// launch epoch).  Shared by policy_kernel and the fused rollout_kernel (epoch = launch epoch + step) so that both
// sample the same actions, bit for bit.
__device__ __forceinline__ void add_exploration_noise(const PolicyParams &P, long long env, unsigned long long epoch,
                                                      float (&a)[4]) {
    const unsigned long long g = (unsigned long long)(env + P.env_offset);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)epoch, ((uint32_t)(epoch >> 32) << 3) | 7u),
                                  make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
    const float u0 = sub_rn(1.0f, u01(r.x)), u1 = u01(r.y), u2 = sub_rn(1.0f, u01(r.z)), u3 = u01(r.w);  // u0, u2 in (0,1]
    const float r0 = sqrtf(mul_rn(-2.0f, logf(u0))), r1 = sqrtf(mul_rn(-2.0f, logf(u2)));
    float s0, c0, s1, c1;
    sincospif(mul_rn(2.0f, u1), &s0, &c0);
    sincospif(mul_rn(2.0f, u3), &s1, &c1);
    a[0] = fmaf(P.std[0], mul_rn(r0, c0), a[0]); a[1] = fmaf(P.std[1], mul_rn(r0, s0), a[1]);
    a[2] = fmaf(P.std[2], mul_rn(r1, c1), a[2]); a[3] = fmaf(P.std[3], mul_rn(r1, s1), a[3]);
}

#ifdef QS_EXP_TS_TIMING  // experiment only (profiles/microbench/policy_stages.cu): clock64 stamps of the stages of a chain
__device__ long long qs_ts_dbg[8192];
#define QS_TS_STAMP(pt) do { if (blockIdx.x == 0 && (tid & 127) == 0 && it < 16) qs_ts_dbg[(((chain * 2 + half) * 16 + it) * 8 + layer) * 8 + (pt)] = clock64(); } while (0)
#define QS_SS_STAMP(pt) do { if (blockIdx.x == 0 && tid == 0 && it < 16) qs_ts_dbg[((group * 16 + it) * 8 + layer) * 8 + (pt)] = clock64(); } while (0)
#define QS_SS_STAMP_AT(it_, layer_, pt) do { if (blockIdx.x == 0 && tid == 0 && (it_) >= 0 && (it_) < 16) qs_ts_dbg[((group * 16 + (it_)) * 8 + (layer_)) * 8 + (pt)] = clock64(); } while (0)
#else
#define QS_TS_STAMP(pt) do { } while (0)
#define QS_SS_STAMP(pt) do { } while (0)
#define QS_SS_STAMP_AT(it_, layer_, pt) do { } while (0)
#endif
// Persistent, ONE CTA per SM made of `groups` (<= 4) independent tile groups of 128 threads.  The groups share the
// weights in shared memory; each owns an A-operand buffer (32 KB), 128 accumulator columns of TMEM, an mbarrier and
// a named block barrier, and walks its own 128-env tiles: thread r of a group owns row r of the tile (its
// observation on the way in, TMEM lane r on the way out).  A tile is a strictly serial chain  A -> MMA -> epilogue
// -> MMA ...  of ~4-5 k cycles of which the tensor pipe is busy ~1.7 k, so four chains in flight per SM are what
// keeps the tensor cores fed: while one group's epilogue runs on the CUDA cores, the others' MMAs run.
__global__ void __launch_bounds__(4 * kPolRows, 1) policy_kernel(const __grid_constant__ PolicyParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem_raw);        // weights landed
    uint64_t *bar_mma_all = bar_w + 1;                               // [group]: a layer's MMAs completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + 64);
    const int groups = blockDim.x / kPolRows;
    const int group = threadIdx.x / kPolRows, tid = threadIdx.x % kPolRows, warp = tid >> 5;
    unsigned char *s_a = smem_raw + 128 + group * ((kPolHidden / 8) * kSlab);
    unsigned char *s_w = smem_raw + 128 + groups * ((kPolHidden / 8) * kSlab);
    uint64_t *bar_mma = bar_mma_all + group;
    uint64_t *bar_obs = reinterpret_cast<uint64_t *>(smem_raw + 72) + group;  // packed observations of a tile landed
    const long long n_tiles = (P.n + kPolRows - 1) / kPolRows;

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int g = 0; g < groups; ++g) { mbar_init(bar_mma_all + g, 1); mbar_init(reinterpret_cast<uint64_t *>(smem_raw + 72) + g, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, P.weight_bytes);
        bulk_load(s_w, P.weights, P.weight_bytes, bar_w);  // launch constant: may precede the PDL wait
    }
    if (threadIdx.x < 32) {  // one warp allocates all accumulator columns (a power of two) and owns the dealloc
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem = tmem_base + (uint32_t)group * kPolHidden;          // this group's accumulator columns
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    pdl_launch_dependents();
    pdl_wait();  // observations come from the previous kernel on the stream (the env step)
    mbar_wait(bar_w, 0);

    const uint32_t a_smem = smem_u32(s_a), w_smem = smem_u32(s_w);
    const bool const_last = P.hidden <= kPolHidden - 8;
    if (const_last) init_const_slab(s_a + tid * 16);  // made visible to the MMAs by the fences of the first layer
    const unsigned long long epoch = P.deterministic ? 0ull : *reinterpret_cast<const volatile unsigned long long *>(P.epoch);
    const bool vec = (P.in_dim & 3) == 0;  // observation rows of 16-byte multiples: float4 loads
    const long long stride = (long long)gridDim.x * groups;
    uint32_t phase = 0, obs_phase = 0;
    int it = 0;
    for (long long tile = (long long)blockIdx.x * groups + group; tile < n_tiles; tile += stride, ++it) {
        const long long env = tile * kPolRows + tid;
        const bool active = env < P.n;
        // ---- A operand of layer 1.  Packed observations (the step kernel's BF16 blocks, one per 32 rows) ARE the operand:
        // sixteen 512-byte TMA copies put the tile's four K-chunk slabs in place; nobody loads, converts or stores a value.
        if (P.obs_packed) {
            const int chunks = pack_chunks(P.in_dim);  // the chunks that travel; a pure-constant chunk behind them is made here
            if (tid == 0) {
                mbar_expect_tx(bar_obs, (uint32_t)chunks * kSlab);
                const uint32_t pkb = (uint32_t)pack_block_bytes(P.in_dim);
                const unsigned char *src = reinterpret_cast<const unsigned char *>(P.obs) + tile * (long long)(4 * pkb);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    for (int c = 0; c < chunks; ++c)
                        bulk_load(s_a + c * kSlab + b * 512, src + b * pkb + c * 512, 512u, bar_obs);
            }
            for (int c = chunks; c < P.k1 / 8; ++c)  // chunks without an observation value: zeros, or (in_dim % 8 == 0) the
                *reinterpret_cast<uint4 *>(s_a + c * kSlab + tid * 16) = make_uint4(c * 8 == P.in_dim ? 0x3F80u : 0u, 0u, 0u, 0u);  // bias' constant 1
            mbar_wait(bar_obs, obs_phase);
            obs_phase ^= 1u;
        } else {
            const float *row = P.obs + env * P.in_dim;
            for (int c = 0; c < P.k1 / 8; ++c) {
                float x[8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = c * 8 + q * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (active) {
                        if (vec) {
                            if (k < P.in_dim) v = *reinterpret_cast<const float4 *>(row + k);
                        } else {
                            if (k + 0 < P.in_dim) v.x = row[k + 0];
                            if (k + 1 < P.in_dim) v.y = row[k + 1];
                            if (k + 2 < P.in_dim) v.z = row[k + 2];
                            if (k + 3 < P.in_dim) v.w = row[k + 3];
                        }
                    }
                    x[q * 4 + 0] = v.x; x[q * 4 + 1] = v.y; x[q * 4 + 2] = v.z; x[q * 4 + 3] = v.w;
                }
                if (P.obs_limit > 0.0f) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[q] = (x[q] == x[q]) ? fminf(fmaxf(x[q], -P.obs_limit), P.obs_limit) : 0.0f;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (c * 8 + q == P.in_dim) x[q] = 1.0f;  // the constant input that multiplies the folded bias
                *reinterpret_cast<uint4 *>(s_a + c * kSlab + tid * 16) =
                    make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            }
        }
        QS_SS_STAMP_AT(it - 1, P.n_hidden, 6);  // the previous tile's last stage: this tile's A operand is in shared memory
        if (!P.obs_packed) {  // the next tile's observation rows: start them towards L2 now, a whole tile of compute before they are read
            const long long nenv = env + stride * kPolRows;
            if (nenv < P.n) {
                const float *nrow = P.obs + nenv * P.in_dim;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + P.in_dim - 1));
            }
        }
        uint32_t w_off = 0;
        for (int layer = 0; layer <= P.n_hidden; ++layer) {
            const bool last = layer == P.n_hidden;
            const int k = layer == 0 ? P.k1 : kPolHidden;
            // generic-proxy writes of A -> visible to the tensor core's async proxy; TMEM reads of the previous
            // epilogue ordered before the MMAs that overwrite the accumulator
            QS_SS_STAMP(0);
            fence_proxy_async();
            tc_fence_before();
            group_barrier(group);
            QS_SS_STAMP(1);
            if (tid == 0) {
                tc_fence_after();
                issue_layer(tmem, a_smem, w_smem + w_off, k, last ? kPolOut : kPolHidden, bar_mma);
            }
            QS_SS_STAMP(2);
            w_off += layer == 0 ? policy_w1_bytes(P.k1) : policy_wh_bytes();
            mbar_wait(bar_mma, phase);
            phase ^= 1u;
            tc_fence_after();
            QS_SS_STAMP(3);
            if (!last) {
                // ---- epilogue: ReLU + BF16 in one conversion, straight into the A slabs of the next layer (the MMAs
                // that read A are done); two 32-column loads in flight
                {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(t_lane, v0);
                    tmem_ld32(t_lane + 32u, v1);
                    tmem_ld_wait();
                    QS_SS_STAMP(4);
                    act_pack_store<4>(v0, s_a + tid * 16, P.activation);
                    act_pack_store<4>(v1, s_a + 4 * kSlab + tid * 16, P.activation);
                    tmem_ld32(t_lane + 64u, v0);
                    if (const_last) { tmem_ld16(t_lane + 96u, v1); tmem_ld8p(t_lane + 112u, v1 + 16); }
                    else tmem_ld32(t_lane + 96u, v1);
                    tmem_ld_wait();
                    act_pack_store<4>(v0, s_a + 8 * kSlab + tid * 16, P.activation);
                    if (const_last) act_pack_store<3>(v1, s_a + 12 * kSlab + tid * 16, P.activation);
                    else act_pack_store<4>(v1, s_a + 12 * kSlab + tid * 16, P.activation);
                    QS_SS_STAMP(5);
                }
            } else {
                uint32_t v[8];
                tmem_ld8(t_lane, v);
                tmem_ld_wait();
                QS_SS_STAMP(4);
                float a[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
                if (active) {
                    if (P.mean) *reinterpret_cast<float4 *>(P.mean + env * 4) = make_float4(a[0], a[1], a[2], a[3]);
                    if (!P.deterministic) add_exploration_noise(P, env, epoch, a);
                    if (P.raw) *reinterpret_cast<float4 *>(P.raw + env * 4) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) a[k2] = fminf(fmaxf(a[k2], -1.0f), 1.0f);  // `nn_controller.c:171-173`
                    *reinterpret_cast<float4 *>(P.actions + env * 4) = make_float4(a[0], a[1], a[2], a[3]);
                }
                QS_SS_STAMP(5);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
    if (threadIdx.x == 0 && !P.deterministic) {  // advance the noise epoch once per launch (same protocol as the step kernel)
        __threadfence();
        if (atomicAdd(P.epoch + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
            P.epoch[1] = 0;
            P.epoch[0] = P.epoch[0] + 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------ activations in TMEM
// policy_kernel_ts: the same network with the activations kept in TENSOR MEMORY between layers.
//
// Why: in policy_kernel both MMA operands come from shared memory.  An M=128, N=128, K=16 BF16 MMA reads 4 KB of A and
// 4 KB of B in the 64 cycles it occupies the tensor pipe = the SM's whole 128 B/clk of shared-memory bandwidth, and the
// epilogues store another 32 KB per layer and tile into the same memory: the four tile groups queue on shared memory
// (profiles/r2/step_kernel_ablation.md section 4; the TMEM read-out, blamed in round 1, runs at 900 B/clk).
// Here a layer's A operand lives in TMEM (tcgen05.mma with [a_tmem]; `UTCHMMA tmem[..], gdesc[..], tmem[..]`): the
// epilogue reads the FP32 accumulator (tcgen05.ld), applies the activation, packs two BF16 per 32-bit column
// (unit 2c in the low half: measured, profiles/microbench/umma_ts_probe.cu) and writes the next layer's A with
// tcgen05.st -- shared memory only serves the weights (B): a third of the traffic, no bank-conflict or proxy-fence
// concerns, and no 32 KB A buffers.
//
// Shape: one persistent CTA per SM, 2 tile groups ("chains") of 256 threads.  A chain owns 192 TMEM columns: 64 for A
// (128 BF16 per row) and 128 for the accumulator; thread = (row, column half): warps w and w+4 of a group address the
// same TMEM lane quadrant and each takes 64 of the 128 accumulator columns, so an epilogue is 2 loads + 32
// conversions + 1 store per thread.  While one chain's epilogue runs, the other chain's MMAs do.
constexpr int kTsThreads = 256;   // per chain
constexpr int kTsChains = 2;
constexpr int kTsCols = 192;      // TMEM columns per chain: A [0, 64) | D [64, 192)

__host__ __device__ constexpr size_t policy_ts_smem_bytes(int k1, int n_hidden) { return 128 + policy_weight_bytes(k1, n_hidden); }

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void chain_barrier(int chain) { asm volatile("bar.sync %0, 256;" ::"r"(chain + 1) : "memory"); }

// 32 accumulator columns -> activation -> 16 packed BF16 words
__device__ __forceinline__ void act_pack16(const uint32_t *v, uint32_t *w, int act) {
    if (act == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) w[c] = pack_relu_bf16(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
    } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) w[c] = pack_bf16(tanh_fast(__uint_as_float(v[2 * c])), tanh_fast(__uint_as_float(v[2 * c + 1])));
    }
}

// one thread: the K/16 MMAs of a layer, D[128 x N] (+)= A[tmem, 128 x K] . W[smem, N x K]^T
__device__ __forceinline__ void issue_layer_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t w_smem, int k, int n_rows, uint64_t *bar) {
    const uint32_t idesc = umma_idesc_bf16(kPolRows, n_rows);
    const uint32_t w_slab = (uint32_t)n_rows * 16u;
    for (int j = 0; j < k / 16; ++j)
        umma_bf16_ts(d_tmem, a_tmem + (uint32_t)j * 8u, umma_desc(w_smem + (uint32_t)j * 2u * w_slab, w_slab, 128), idesc, j > 0);
    umma_commit(bar);
}

// this thread's 16 observation values [16 * half, 16 * half + 16) of row `env` (zero beyond in_dim, 1 at in_dim)
__device__ __forceinline__ void load_obs16(const PolicyParams &P, long long env, bool active, int k0, float (&x)[16]) {
    const float *row = P.obs + env * P.in_dim;
    const bool vec = (P.in_dim & 3) == 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = k0 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            if (vec) {
                if (k < P.in_dim) v = *reinterpret_cast<const float4 *>(row + k);
            } else {
                if (k + 0 < P.in_dim) v.x = row[k + 0];
                if (k + 1 < P.in_dim) v.y = row[k + 1];
                if (k + 2 < P.in_dim) v.z = row[k + 2];
                if (k + 3 < P.in_dim) v.w = row[k + 3];
            }
        }
        x[q * 4 + 0] = v.x; x[q * 4 + 1] = v.y; x[q * 4 + 2] = v.z; x[q * 4 + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
        if (k0 + q == P.in_dim) x[q] = 1.0f;  // the constant input that multiplies the folded bias
}
__device__ __forceinline__ void store_obs16(uint32_t taddr, const float (&x)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = pack_bf16(x[2 * c], x[2 * c + 1]);
    tmem_st8(taddr, w);
}

__global__ void __launch_bounds__(kTsChains * kTsThreads, 1) policy_kernel_ts(const __grid_constant__ PolicyParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem_raw);        // weights landed
    uint64_t *bar_mma_all = bar_w + 1;                               // [chain]: a layer's MMAs completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + 64);
    unsigned char *s_w = smem_raw + 128;
    const int chain = threadIdx.x / kTsThreads, tid = threadIdx.x % kTsThreads;
    const int warp = tid >> 5, quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + (tid & 31);                          // row of the tile = TMEM lane
    uint64_t *bar_mma = bar_mma_all + chain;
    const long long n_tiles = (P.n + kPolRows - 1) / kPolRows;

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int g = 0; g < kTsChains; ++g) mbar_init(bar_mma_all + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, P.weight_bytes);
        bulk_load(s_w, P.weights, P.weight_bytes, bar_w);  // launch constant: may precede the PDL wait
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_chain = tmem_base + (uint32_t)chain * kTsCols;                 // lane 0 of this chain's columns
    const uint32_t t_a = t_chain + ((uint32_t)(quad * 32) << 16);                   // this thread's lane quadrant: A columns
    const uint32_t t_d = t_a + 64u;                                                 // ... accumulator columns
    pdl_launch_dependents();
    pdl_wait();  // observations come from the previous kernel on the stream (the env step)
    mbar_wait(bar_w, 0);

    const uint32_t w_smem = smem_u32(s_w);
    const bool const_last = P.hidden <= kPolHidden - 8;
    if (const_last && half == 1) {  // units 120..126 = padding, unit 127 = the constant 1: columns 60..63 of A, written once
        const uint32_t c[4] = {0u, 0u, 0u, 0x3F800000u};
        tmem_st4(t_a + 60u, c);
    }
    const unsigned long long epoch = P.deterministic ? 0ull : *reinterpret_cast<const volatile unsigned long long *>(P.epoch);
    const long long stride = (long long)gridDim.x * kTsChains;
    const bool pre = P.k1 == 32;  // the common widths (in_dim <= 31): the next tile's row is fetched a whole tile ahead
    float nx[16];
    {
        const long long tile = (long long)blockIdx.x * kTsChains + chain;
        if (pre && tile < n_tiles) load_obs16(P, tile * kPolRows + row, tile * kPolRows + row < P.n, 16 * half, nx);
    }
    uint32_t phase = 0;
    int it = 0;
    for (long long tile = (long long)blockIdx.x * kTsChains + chain; tile < n_tiles; tile += stride, ++it) {
        const long long env = tile * kPolRows + row;
        const bool active = env < P.n;
        // ---- A operand of layer 1: the observation row, BF16, two values per TMEM column; thread = (row, 16 columns)
        if (pre) {
            store_obs16(t_a + 8u * (uint32_t)half, nx);
        } else {
            for (int k0 = 16 * half; k0 < P.k1; k0 += 32) {
                float x[16];
                load_obs16(P, env, active, k0, x);
                store_obs16(t_a + (uint32_t)(k0 >> 1), x);
            }
        }
        tmem_st_wait();
        uint32_t w_off = 0;
        for (int layer = 0; layer <= P.n_hidden; ++layer) {
            const bool last = layer == P.n_hidden;
            const int k = layer == 0 ? P.k1 : kPolHidden;
            // A (tcgen05.st, waited for) and the accumulator reads of the previous epilogue are ordered before the MMAs
            QS_TS_STAMP(0);
            tc_fence_before();
            chain_barrier(chain);
            QS_TS_STAMP(1);
            if (tid == 0) {
                tc_fence_after();
                issue_layer_ts(t_chain + 64u, t_chain, w_smem + w_off, k, last ? kPolOut : kPolHidden, bar_mma);
            }
            QS_TS_STAMP(2);
            if (layer == 0 && pre) {  // the loads fly while the chain computes
                const long long nenv = env + stride * kPolRows;
                if (tile + stride < n_tiles) load_obs16(P, nenv, nenv < P.n, 16 * half, nx);
            }
            w_off += layer == 0 ? policy_w1_bytes(P.k1) : policy_wh_bytes();
            mbar_wait(bar_mma, phase);
            phase ^= 1u;
            tc_fence_after();
            QS_TS_STAMP(3);
            if (!last) {
                // ---- epilogue: accumulator columns [64 half, 64 half + 64) -> activation -> A columns [32 half, 32 half + 32)
                uint32_t v0[32], v1[32], w[32];
                tmem_ld32(t_d + 64u * (uint32_t)half, v0);
                tmem_ld32(t_d + 64u * (uint32_t)half + 32u, v1);
                tmem_ld_wait();
                QS_TS_STAMP(4);
                act_pack16(v0, w, P.activation);
                act_pack16(v1, w + 16, P.activation);
                if (half == 1 && const_last) {  // columns 60..63 keep the constant
                    tmem_st16(t_a + 32u, w);
                    tmem_st8(t_a + 48u, w + 16);
                    tmem_st4(t_a + 56u, w + 24);
                } else {
                    tmem_st16(t_a + 32u * (uint32_t)half, w);
                    tmem_st16(t_a + 32u * (uint32_t)half + 16u, w + 16);
                }
                QS_TS_STAMP(5);
                tmem_st_wait();
                QS_TS_STAMP(6);
            } else if (half == 0) {
                uint32_t v[8];
                tmem_ld8(t_d, v);
                tmem_ld_wait();
                float a[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
                if (active) {
                    if (P.mean) *reinterpret_cast<float4 *>(P.mean + env * 4) = make_float4(a[0], a[1], a[2], a[3]);
                    if (!P.deterministic) add_exploration_noise(P, env, epoch, a);
                    if (P.raw) *reinterpret_cast<float4 *>(P.raw + env * 4) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) a[k2] = fminf(fmaxf(a[k2], -1.0f), 1.0f);  // `nn_controller.c:171-173`
                    *reinterpret_cast<float4 *>(P.actions + env * 4) = make_float4(a[0], a[1], a[2], a[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    if (threadIdx.x == 0 && !P.deterministic) {  // advance the noise epoch once per launch (same protocol as the step kernel)
        __threadfence();
        if (atomicAdd(P.epoch + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
            P.epoch[1] = 0;
            P.epoch[0] = P.epoch[0] + 1;
        }
    }
}


// Generalised advantage estimation over a device-resident rollout (SB3 `RolloutBuffer.compute_returns_and_advantage`,
// the step after `collect_rollouts` in `model.learn`, `3D quad race.ipynb:820`): one thread per env walks its column
// of the (steps, N) buffers backwards -- every access of a warp is one contiguous run.
//   delta_t = r_t + gamma * V_{t+1} * (1 - done_t) - V_t ;  A_t = delta_t + gamma * lambda * (1 - done_t) * A_{t+1}
// values has steps+1 rows (the last one bootstraps).  done_t ends the episode AFTER step t (the env has already been
// reset when obs_{t+1} was written), so V_{t+1} belongs to the next episode and is masked.
__global__ void gae_kernel(const float *rew, const float *val, const uint8_t *done, float *adv, float *ret, long long n,
                           int steps, float gamma, float lambda) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.0f, v_next = val[(long long)steps * n + i];
    for (int t = steps - 1; t >= 0; --t) {
        const long long k = (long long)t * n + i;
        const float nd = done[k] ? 0.0f : 1.0f, v = val[k];
        const float delta = fmaf(gamma * nd, v_next, rew[k]) - v;
        a = fmaf(gamma * lambda * nd, a, delta);
        adv[k] = a;
        ret[k] = a + v;
        v_next = v;
    }
}

}  // namespace qs
