// quadsim_rollout.cuh -- sm_100a: the closed control loop in ONE kernel (SURVEY.md section 8 rows f1 + f2).
//
// What it replaces: SB3's `collect_rollouts` (behind `model.learn`, `3D quad race.ipynb:820`): n_steps times
//   actions = policy(obs)  ->  obs, reward, done = env.step(actions)  ->  rollout_buffer.add(...)
// The unfused device path (qs_rollout) runs that as 2 * n_steps launches: every step the policy kernel re-reads the
// observations the step kernel just wrote, and the step kernel re-reads and re-writes the whole simulator state
// (413 B of HBM traffic per env-step, of which only the 133 B that land in the rollout buffers are wanted).
//
// Here a 128-thread tile group keeps its 128 quads IN REGISTERS for all n_steps: per step it
//   1. feeds the observation rows (BF16, UMMA K-major slabs in shared memory) through the controller MLP on the
//      tensor cores (tcgen05.mma, FP32 accumulators in TMEM, ReLU + BF16 re-pack between layers -- the same chain as
//      policy_kernel),
//   2. samples / clips the action in the last epilogue (thread = env), writes it to the rollout buffers,
//   3. advances its quad with the SAME device functions as step_kernel (euler_step, reward_and_flags,
//      draw_reset_warp, write_obs), writes reward / done,
//   4. stages the new observation rows in the (idle) upper part of its A-operand buffer: one TMA bulk store sends the
//      tile to obs_buf[t+1] while the rows are converted to BF16 for the next step's first GEMM.
// The simulator state touches HBM twice per launch (load, store) instead of twice per step; observations never come
// back from HBM.  Four groups per SM overlap one group's CUDA-core phase (epilogues, env step) with the others' MMAs.
//
// Bit-compatibility: reset draws and exploration noise are Philox blocks keyed by (seed, global env, launch epoch);
// the unfused path advances the epoch once per launch = once per step, this kernel uses (epoch at launch + t) and
// adds n_steps at the end -- a fused rollout reproduces the unfused one bit for bit (tests/test_gpu_rollout_fused.py).
#pragma once
#include "quadsim_policy.cuh"

namespace qs {

struct RolloutParams {
    StepParams S;     // env side (S.actions / S.obs / S.rew / S.done / S.flags are not used)
    PolicyParams Q;   // policy side (Q.obs / Q.actions / Q.mean / Q.raw are not used)
    float *obs_buf;   // (steps+1, n, obs_len); row block 0 = the observations of the current state (not written)
    float *act_buf;   // (steps, n, 4) clipped actions
    float *raw_buf;   // (steps, n, 4) un-clipped samples, or NULL
    float *rew_buf;   // (steps, n)
    uint8_t *done_buf;  // (steps, n)
    uint8_t *flags_buf; // (steps, n) F_* bits of every step (TimeLimit.truncated = F_TRUNC), or NULL
    int steps;
};

// the f32 staging tile of a group lives behind the layer-1 slabs of its A buffer
__host__ __device__ constexpr bool rollout_fused_fits(int k1, int obs_len) {
    return (k1 / 8) * kSlab + kPolRows * obs_len * 4 <= (kPolHidden / 8 - 1) * kSlab;  // slab 15 holds a constant
}
// dynamic shared memory = the policy kernel's + the track table
__host__ __device__ constexpr size_t rollout_smem_bytes(int k1, int n_hidden, int groups, int n_gates) {
    return policy_smem_bytes(k1, n_hidden, groups) + (size_t)n_gates * kTrackRow * 4;
}

template <int V>
__global__ void __launch_bounds__(4 * kPolRows, 1) rollout_kernel(const __grid_constant__ RolloutParams R) {
    const StepParams &P = R.S;
    const PolicyParams &Q = R.Q;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *bar_mma_all = bar_w + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + 64);
    const int groups = blockDim.x / kPolRows;
    const int group = threadIdx.x / kPolRows, tid = threadIdx.x % kPolRows, warp = tid >> 5;
    constexpr int kABytes = (kPolHidden / 8) * kSlab;
    unsigned char *s_a = smem_raw + 128 + group * kABytes;
    unsigned char *s_w = smem_raw + 128 + groups * kABytes;
    float *s_track = reinterpret_cast<float *>(s_w + Q.weight_bytes);
    uint64_t *bar_mma = bar_mma_all + group;
    const int D = P.obs_len;
    float *s_stage = reinterpret_cast<float *>(s_a + (Q.k1 / 8) * kSlab);  // [128][D] f32, idle while no hidden layer is live
    float *my_row = s_stage + tid * D;
    const long long n_tiles = (P.n + kPolRows - 1) / kPolRows;

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int g = 0; g < groups; ++g) mbar_init(bar_mma_all + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, Q.weight_bytes);
        bulk_load(s_w, Q.weights, Q.weight_bytes, bar_w);
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Q.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < P.n_gates * kTrackRow; i += blockDim.x) s_track[i] = P.track[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem = tmem_base + (uint32_t)group * kPolHidden;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    pdl_launch_dependents();
    pdl_wait();  // simulator state and epochs come from earlier kernels on the stream
    mbar_wait(bar_w, 0);

    const uint32_t a_smem = smem_u32(s_a), w_smem = smem_u32(s_w);
    const bool const_last = Q.hidden <= kPolHidden - 8;  // slab 15 = {0 x 7, 1}: written once, see epilogue_tail
    if (const_last) init_const_slab(s_a + tid * 16);
    const unsigned long long step_epoch0 = load_epoch(P.epoch);
    const unsigned long long pol_epoch0 = Q.deterministic ? 0ull : load_epoch(Q.epoch);
    const size_t nD = (size_t)P.n * D;

    float reward_acc = 0.0f;
    unsigned c_act = 0, c_done = 0, c_tr = 0, c_gp = 0, c_gc = 0, c_gr = 0, c_ob = 0;
    uint32_t phase = 0;
    const long long stride = (long long)gridDim.x * groups;
    for (long long tile = (long long)blockIdx.x * groups + group; tile < n_tiles; tile += stride) {
        const long long base = tile * kPolRows;
        const long long env = base + tid;
        const bool active = env < P.n;
        const long long rem = P.n - base;
        const int rows = rem < kPolRows ? (int)rem : kPolRows;
        const uint32_t tile_bytes = (uint32_t)rows * (uint32_t)D * 4u;

        // ---- this thread's quad: state -> registers for the whole rollout
        EnvState<V> e;
        uint32_t tg = 0, sc = 0;
        if (active) {
            load_state<V>(P.s, env, e);
            const uint32_t meta = field<V, Blk<V>::META, uint32_t>(P.s, env);
            tg = meta >> 24; sc = meta & kStepMask;
        } else {  // padding lanes of the last tile compute on a harmless state and never store
            e.x = e.y = e.vx = e.vy = e.vz = e.phi = e.th = e.psi = e.p = e.q = e.r = 0.0f; e.z = -1.0f;
#pragma unroll
            for (int j = 0; j < (V == kE2E ? 4 : 1); ++j) e.w[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < 6; ++j) e.dist[j] = 0.0f;
        }
        write_obs<V>(P, s_track, e, tg, my_row);  // = obs_buf[0] rows of this tile (already in the buffer)

        for (int t = 0; t < R.steps; ++t) {
            // ---- A operand of layer 1 from this thread's staged row: BF16, K-chunk by K-chunk; column D = 1
            for (int c = 0; c < Q.k1 / 8; ++c) {
                float x[8];
                if ((D & 3) == 0) {  // 16-byte rows: float4 reads (the same access pattern write_obs stored them with)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int k = c * 8 + q * 4;
                        float4 v = make_float4(k == D ? 1.0f : 0.0f, 0.f, 0.f, 0.f);
                        if (k < D) v = *reinterpret_cast<const float4 *>(my_row + k);
                        x[q * 4 + 0] = v.x; x[q * 4 + 1] = v.y; x[q * 4 + 2] = v.z; x[q * 4 + 3] = v.w;
                    }
                } else {  // odd row length (INDI: 17 floats): scalar reads are conflict-free
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k = c * 8 + q;
                        x[q] = k < D ? my_row[k] : (k == D ? 1.0f : 0.0f);
                    }
                }
                *reinterpret_cast<uint4 *>(s_a + c * kSlab + tid * 16) =
                    make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            }
            uint32_t w_off = 0;
            float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int layer = 0; layer <= Q.n_hidden; ++layer) {
                const bool last = layer == Q.n_hidden;
                const int k = layer == 0 ? Q.k1 : kPolHidden;
                fence_proxy_async();  // A slabs (and, before layer 0, the staged f32 rows) -> async proxy
                tc_fence_before();
                group_barrier(group);
                if (tid == 0) {
                    tc_fence_after();
                    const bool send_obs = layer == 0 && t > 0 && rows > 0;  // rows staged by step t-1 = obs_buf[t]
                    float *dst = R.obs_buf + (size_t)t * nD + (size_t)base * D;
                    const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((tile_bytes & 15u) == 0);
                    if (send_obs && bulk) bulk_store(dst, s_stage, tile_bytes);
                    const uint32_t idesc = umma_idesc_bf16(kPolRows, last ? kPolOut : kPolHidden);
                    const uint32_t w_slab = (uint32_t)(last ? kPolOut : kPolHidden) * 16u;
                    for (int j = 0; j < k / 16; ++j)
                        umma_bf16(tmem, umma_desc(a_smem + (uint32_t)j * 2u * kSlab, kSlab, 128),
                                  umma_desc(w_smem + w_off + (uint32_t)j * 2u * w_slab, w_slab, 128), idesc, j > 0);
                    // the staging tile is overwritten by the layer-1 epilogue: the bulk store must have READ it before
                    // anyone passes the MMA barrier -- so the commit is issued after the read-wait (the MMAs run meanwhile)
                    if (send_obs && bulk) bulk_store_wait_read();
                    umma_commit(bar_mma);
                }
                if (layer == 0 && t > 0) {  // unaligned observation rows (odd n * obs_len): plain stores by the group
                    float *dst = R.obs_buf + (size_t)t * nD + (size_t)base * D;
                    const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((tile_bytes & 15u) == 0);
                    if (!bulk) {
                        for (int i = tid; i < rows * D; i += kPolRows) dst[i] = s_stage[i];
                        group_barrier(group);  // all reads of the staging tile precede the epilogue's writes
                    }
                }
                w_off += layer == 0 ? policy_w1_bytes(Q.k1) : policy_wh_bytes();
                mbar_wait(bar_mma, phase);
                phase ^= 1u;
                tc_fence_after();
                if (!last) {
                    uint32_t v0[32];
#pragma unroll 1
                    for (int c = 0; c < 3; ++c) {
                        tmem_ld32(t_lane + (uint32_t)c * 32u, v0);
                        tmem_ld_wait();
                        act_pack_store<4>(v0, s_a + (c * 4) * kSlab + tid * 16, Q.activation);
                    }
                    epilogue_tail(t_lane, s_a + tid * 16, const_last, v0, Q.activation);
                } else {
                    uint32_t v[8];
                    tmem_ld8(t_lane, v);
                    tmem_ld_wait();
                    float a[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
                    if (!Q.deterministic) add_exploration_noise(Q, env, pol_epoch0 + (unsigned long long)t, a);
                    const size_t k4 = ((size_t)t * (size_t)P.n + (size_t)env) * 4;
                    if (active && R.raw_buf) *reinterpret_cast<float4 *>(R.raw_buf + k4) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) a[k2] = fminf(fmaxf(a[k2], -1.0f), 1.0f);
                    u = make_float4(a[0], a[1], a[2], a[3]);
                    if (active) *reinterpret_cast<float4 *>(R.act_buf + k4) = u;
                }
            }

            // ---- env.step(u) for this thread's quad (`3D quad race.ipynb:501-595`, normal branch, fused device reset)
            EnvState<V> n;
            euler_step<V>(P, e, u, n);
            float reward; bool dn; uint32_t fl;
            reward_and_flags<V>(P, s_track, e, n, tg, sc, reward, dn, fl);
            const bool need = dn && active;
            if (__any_sync(0xffffffffu, need)) {  // scratch: this warp's rows of the staging tile (already sent)
                draw_reset_warp<V>(P, base + warp * 32, need, reinterpret_cast<uint4 *>(s_stage + warp * 32 * D), n,
                                   step_epoch0 + (unsigned long long)t);
            }
            if (need) { tg = 0; sc = 0; }
            if (active) {
                const size_t k1 = (size_t)t * (size_t)P.n + (size_t)env;
                R.rew_buf[k1] = reward;
                R.done_buf[k1] = (uint8_t)(dn ? 1 : 0);
                if (R.flags_buf) R.flags_buf[k1] = (uint8_t)fl;
                if (P.stats) {
                    reward_acc += reward;
                    c_act += 1; c_done += (fl & F_DONE) != 0; c_tr += (fl & F_TRUNC) != 0; c_gp += (fl & F_PASSED) != 0;
                    c_gc += (fl & F_COLLISION) != 0; c_gr += (fl & F_GROUND) != 0; c_ob += (fl & F_OOB) != 0;
                }
            }
            e = n;
            __syncwarp();
            write_obs<V>(P, s_track, e, tg, my_row);  // obs_buf[t+1] rows: sent by the next iteration / the tail below
        }

        // ---- tail: the last observation rows, and the state back to HBM
        {
            fence_proxy_async();
            group_barrier(group);
            float *dst = R.obs_buf + (size_t)R.steps * nD + (size_t)base * D;
            const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((tile_bytes & 15u) == 0);
            if (bulk) {
                if (tid == 0 && rows > 0) { bulk_store(dst, s_stage, tile_bytes); bulk_store_wait_read(); }
            } else {
                for (int i = tid; i < rows * D; i += kPolRows) dst[i] = s_stage[i];
            }
            group_barrier(group);  // the next tile's first write_obs overwrites the staging tile
        }
        if (active) {
            store_world<V>(P.s, env, e);
            store_dist<V>(P.s, env, e);
            field<V, Blk<V>::META, uint32_t>(P.s, env) = (tg << 24) | sc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Q.tmem_cols) : "memory");
    if (threadIdx.x == 0) {  // both launch epochs advance by `steps`, once, after every CTA has read them
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned long long *>(P.epoch) + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
            P.epoch[1] = 0;
            P.epoch[0] = step_epoch0 + (unsigned long long)R.steps;
            if (!Q.deterministic) Q.epoch[0] = pol_epoch0 + (unsigned long long)R.steps;
        }
    }

    if (P.stats) {  // per-CTA totals into the CTA's own slot (no atomics), as step_kernel does
        __shared__ float s_red_f[16];
        __shared__ unsigned s_red_u[16][7];
        const unsigned full_mask = 0xffffffffu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) reward_acc += __shfl_xor_sync(full_mask, reward_acc, o);
        c_act = __reduce_add_sync(full_mask, c_act); c_done = __reduce_add_sync(full_mask, c_done);
        c_tr = __reduce_add_sync(full_mask, c_tr); c_gp = __reduce_add_sync(full_mask, c_gp);
        c_gc = __reduce_add_sync(full_mask, c_gc); c_gr = __reduce_add_sync(full_mask, c_gr);
        c_ob = __reduce_add_sync(full_mask, c_ob);
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
            s_red_f[w] = reward_acc;
            s_red_u[w][0] = c_act; s_red_u[w][1] = c_done; s_red_u[w][2] = c_tr; s_red_u[w][3] = c_gp;
            s_red_u[w][4] = c_gc; s_red_u[w][5] = c_gr; s_red_u[w][6] = c_ob;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float r = 0.0f;
            unsigned c[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
                r += s_red_f[k];
#pragma unroll
                for (int j = 0; j < 7; ++j) c[j] += s_red_u[k][j];
            }
            if (c[0]) {
                Stats *st = P.stats + blockIdx.x;
                st->reward_sum += (double)r;
                st->env_steps += c[0]; st->dones += c[1]; st->truncated += c[2]; st->gates_passed += c[3];
                st->gate_collisions += c[4]; st->ground_collisions += c[5]; st->out_of_bounds += c[6];
            }
        }
    }
}

}  // namespace qs
