// quadsim_kernels.cuh -- sm_100a device code of the quadrotor racing step.
//
// One thread advances one quadrotor through the reference's whole step_wait (`3D quad race.ipynb:501-595`;
// INDI `...INDI inner loop.ipynb:303-385`): residual MLPs (`:244-262`) -> equations of motion (`:65-152`) ->
// forward Euler (`:512`) -> reward / gate / termination flags (`:516-566`) -> branch logic + masked reset
// (`:568-585`, `:452-493`) -> gate-frame observation (`:365-450`).  Nothing is re-read: 285 algorithmic bytes
// per env-step (E2E, gates_ahead=1), 209 for INDI.
//
// Data layout: world state as four float4 planes  P0=(x,y,z,vx) P1=(vy,vz,phi,theta)
// P2=(psi,p,q,r) P3=(w1..w4 | INDI: scalar T_norm plane); disturbances as float4 (Mx,My,Mz,Fz) + float2 (Fx,Fy);
// counters packed in one u32 (target_gate<<24 | step_count).  Every global access of a warp is one contiguous
// 128/256/512-byte run.  Observations are staged row-major in shared memory and leave the SM as ONE
// cp.async.bulk (TMA bulk store) per thread block, because the (N,D) row-major tile of a block is contiguous.
//
// Numerics: positions, gate-plane projections and distances use explicitly rounded mul/add (no FMA
// contraction) so that done / gate flags are bit-identical to the reference's float32 NumPy arithmetic; the
// rest of the dynamics is free to contract (measured <= 2e-6 scaled error, gate is 1e-5).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qs {

constexpr int kBlock = 128;          // threads (= envs) per CTA
constexpr int kTrackRow = 12;        // floats per gate in the device track table
constexpr uint32_t kStepMask = 0x00FFFFFFu;

enum : int { kE2E = 0, kINDI = 1 };
enum : int { kModeNormal = 0, kModePauseIfCollision = 1, kModePause = 2 };
enum : int { kResetDevice = 0, kResetHost = 1 };
enum : uint32_t { F_DONE = 1, F_TRUNC = 2, F_PASSED = 4, F_COLLISION = 8, F_GROUND = 16, F_OOB = 32 };

struct Stats {  // must match qs_stats
    double reward_sum;
    unsigned long long env_steps, dones, truncated, gates_passed, gate_collisions, ground_collisions, out_of_bounds;
};

// Simulator state in HBM: array of structs of arrays, one BLOCK per 32 envs (= one warp).  Inside a block every
// field is a 32-lane plane of 16-byte words, so a warp's access to a plane is one 512-byte run, and the whole
// block is ONE contiguous run: a single TMA bulk copy brings all of a warp's state into shared memory.
template <int V> struct Blk;
template <> struct Blk<kE2E> {   // bytes: P0=(x,y,z,vx) P1=(vy,vz,phi,theta) P2=(psi,p,q,r) P3=(w1..w4) float4 planes,
    enum : int { P0 = 0, P1 = 512, P2 = 1024, P3 = 1536, META = 2048, DA = 2176, DB = 2688, BYTES = 2944 };
};                               // META u32 (target_gate<<24 | step_count), DA=(Mx,My,Mz,Fz) float4, DB=(Fx,Fy) float2
template <> struct Blk<kINDI> {  // P3 = scalar T_norm plane
    enum : int { P0 = 0, P1 = 512, P2 = 1024, P3 = 1536, META = 1664, BYTES = 1792, DA = 0, DB = 0 };
};

struct Planes {
    unsigned char *base;  // blocks, Blk<V>::BYTES apart
};
template <int V, int OFF, typename T>
__device__ __forceinline__ T &field(const Planes &s, long long env) {
    return reinterpret_cast<T *>(s.base + (env >> 5) * (long long)Blk<V>::BYTES + OFF)[env & 31];
}

struct ResetDist {   // reset_ draw ranges (`:455-489`)
    float start[3];
    float dist_lo[6], dist_span[6];  // already multiplied by disturbance_scale
};

struct StepParams {
    Planes s;
    const float4 *actions;
    float *obs;
    // fused observation all-gather: the same rows are also stored straight into the other ranks' gather buffers
    // (peer memory over NVLink): peer_obs[p] + (peer_row_offset + env) * obs_len
    float *peer_obs[7];
    long long peer_row_offset;
    int n_peers;
    float *rew;
    uint8_t *done;
    uint8_t *flags;
    const float *track;  // (n_gates, kTrackRow): gx gy gz yaw cos sin 0 0 | rel_x rel_y rel_z rel_yaw
    Stats *stats;        // one slot per CTA of the step kernel (or NULL)
    long long n;
    long long env_offset;
    long long tile_begin, tile_end;  // 128-env tiles [begin, end) this launch steps (the whole env range, or one
                                     // chunk of the host-buffer pipeline); advance_epoch marks a step's LAST launch
    int advance_epoch;
    unsigned long long seed;
    // Time half of the RNG key: number of step / reset launches so far.  It lives in DEVICE memory (epoch[0]) so that
    // a CUDA graph replaying the same launch draws fresh values; the last CTA of a launch to finish (epoch[1] counts
    // arrivals) advances it, i.e. after every CTA of this launch has read it and before the next launch may.
    unsigned long long *epoch;
    // Chained launches (consecutive full-range steps of one env inside a CUDA graph): chain[2b] counts the launches CTA b
    // has started, chain[2b+1] the ones it has finished.  CTA b steps the same tiles in every launch, so with
    // chain_wait = 1 it waits for CTA b of the previous launch only (not for the whole grid, griddepcontrol.wait):
    // the next step's loads flow while this step's last tiles drain.  NULL: classic launch.
    unsigned int *chain;
    int chain_wait;
    int n_gates, gates_ahead, obs_len;
    // 0: observations leave as float32 rows (N, obs_len), the reference's layout.  1: as BF16 in the on-device policy's
    // A-operand layout (pack_obs_row / pack_block_bytes below) -- `obs` and `peer_obs` then point to packed buffers
    int obs_packed;
    int mode, reset_source;
    int l2_hints;  // 1: state evict_last, streams evict_first (see l2_policy_*); 0: no cache hints
    long long keep_blocks;  // with l2_hints: only the first keep_blocks 32-env state blocks are pinned (what fits in L2)
    uint32_t max_steps;
    float dt;
    // disturbance observation 2*(d-lo)/(hi-lo)-1 (`:414-448`) as ((d - lo_hi) - lo_lo) * scale - 1: the subtraction
    // comes first, as in the reference, so narrow ranges far from 0 do not cancel; lo_hi + lo_lo carries a float64 lo
    float obs_lo[4], obs_lo2[4], obs_scale[4];
    ResetDist rd;
    // residual MLPs in KERNEL layout (qs_set_residual_weights transposes): layer 1 is [in][hidden] so that two
    // adjacent hidden units form one 64-bit constant-bank operand of a packed FFMA2
    alignas(16) float wt1[7 * 32];   // thrust  W1^T
    alignas(16) float bt1[32];
    alignas(16) float wt2[32];       // thrust  W2 (1x32)
    alignas(16) float wm1[10 * 32];  // moment  W1^T
    alignas(16) float bm1[32];
    alignas(16) float wm2[3 * 32];   // moment  W2 (3x32)
    alignas(16) float b2[4];         // thrust b2, moment b2[3]
};

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// np.linalg.norm(axis=1) of a float32 (N,3) array: sqrt((x*x + y*y) + z*z), each op rounded
__device__ __forceinline__ float norm3_rn(float a, float b, float c) {
    return __fsqrt_rn(add_rn(add_rn(mul_rn(a, a), mul_rn(b, b)), mul_rn(c, c)));
}

// yaw %= 2*pi ; yaw[yaw > pi] -= 2*pi ; yaw[yaw < -pi] += 2*pi   in float32 (`:393-396`).
// floor-quotient + one FMA reproduces np.remainder exactly: a - n*b is a multiple of ulp(b) below b, hence
// representable, and for |a| < b both sides perform the same single rounded add.
__device__ __noinline__ float wrap_yaw_slow(float a) {  // blown-up state: exact, out of line
    const float b = 6.283185307179586f;
    float m = fmodf(a, b);
    if (m < 0.0f) m = add_rn(m, b);
    return m;
}
__device__ __forceinline__ float wrap_yaw(float a) {
    const float b = 6.283185307179586f, pi = 3.141592653589793f;
    float m;
    if (fabsf(a) < 1.0e5f) {
        float n = floorf(a * 0.15915494309189535f);
        m = fmaf(-n, b, a);
        if (m < 0.0f) m = add_rn(m, b);
        if (m >= b) m = sub_rn(m, b);
    } else {
        m = wrap_yaw_slow(a);
    }
    if (m > pi) m = sub_rn(m, b);
    if (m < -pi) m = add_rn(m, b);
    return m;
}

// Philox4x32-10 (Salmon et al. 2011)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }  // [0,1)
__device__ __forceinline__ float uni(uint32_t x, float lo, float span) { return fmaf(u01(x), span, lo); }

template <int V>
struct EnvState {
    float x, y, z, vx, vy, vz, phi, th, psi, p, q, r;
    float w[V == kE2E ? 4 : 1];  // motor speeds | T_norm
    float dist[6];               // Mx My Mz Fx Fy Fz (E2E only)
};

template <int V>
__device__ __forceinline__ void load_state(const Planes &s, long long i, EnvState<V> &e) {
    using B = Blk<V>;
    const float4 a = field<V, B::P0, float4>(s, i), b = field<V, B::P1, float4>(s, i), c = field<V, B::P2, float4>(s, i);
    e.x = a.x; e.y = a.y; e.z = a.z; e.vx = a.w;
    e.vy = b.x; e.vz = b.y; e.phi = b.z; e.th = b.w;
    e.psi = c.x; e.p = c.y; e.q = c.z; e.r = c.w;
    if (V == kE2E) {
        const float4 d = field<V, B::P3, float4>(s, i);
        e.w[0] = d.x; e.w[1] = d.y; e.w[2] = d.z; e.w[V == kE2E ? 3 : 0] = d.w;
        const float4 da = field<V, B::DA, float4>(s, i);
        const float2 db = field<V, B::DB, float2>(s, i);
        e.dist[0] = da.x; e.dist[1] = da.y; e.dist[2] = da.z; e.dist[5] = da.w;
        e.dist[3] = db.x; e.dist[4] = db.y;
    } else {
        e.w[0] = field<V, B::P3, float>(s, i);
    }
}

template <int V>
__device__ __forceinline__ void store_world(const Planes &s, long long i, const EnvState<V> &e) {
    using B = Blk<V>;
    field<V, B::P0, float4>(s, i) = make_float4(e.x, e.y, e.z, e.vx);
    field<V, B::P1, float4>(s, i) = make_float4(e.vy, e.vz, e.phi, e.th);
    field<V, B::P2, float4>(s, i) = make_float4(e.psi, e.p, e.q, e.r);
    if (V == kE2E) field<V, B::P3, float4>(s, i) = make_float4(e.w[0], e.w[1], e.w[2], e.w[V == kE2E ? 3 : 0]);
    else field<V, B::P3, float>(s, i) = e.w[0];
}

template <int V>
__device__ __forceinline__ void store_dist(const Planes &s, long long i, const EnvState<V> &e) {
    if (V == kE2E) {
        field<V, Blk<V>::DA, float4>(s, i) = make_float4(e.dist[0], e.dist[1], e.dist[2], e.dist[5]);
        field<V, Blk<V>::DB, float2>(s, i) = make_float2(e.dist[3], e.dist[4]);
    }
}

// reset_ (`:452-489`) with the device RNG: same fields, same ranges, counter-based instead of MT19937.
// Draw c of env g reset at launch `epoch` is Philox4x32-10(counter=(g_lo, g_hi, epoch_lo, epoch_hi<<3 | c), key=seed):
// 6 draws (E2E) / 4 (INDI).  An env resets at most once per launch, so (g, epoch) names the event; nothing
// per-env has to be stored or read, and the stream does not depend on how N is sharded.
template <int V> struct ResetDraws { enum : int { N = (V == kE2E ? 6 : 4) }; };

__device__ __forceinline__ unsigned long long load_epoch(const unsigned long long *epoch) {
    return *reinterpret_cast<const volatile unsigned long long *>(epoch);
}
// `epoch` = the launch counter the draw is keyed by: load_epoch(P.epoch) for a one-step launch; the fused rollout
// kernel, which advances many steps per launch, passes (epoch at launch + step) so that both paths draw the same values
__device__ __forceinline__ uint4 reset_draw(const StepParams &P, unsigned long long g, uint32_t c, unsigned long long epoch) {
    return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)epoch, ((uint32_t)(epoch >> 32) << 3) | c),
                         make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
}
// end of a launch that may have drawn resets: called by ONE thread per CTA after its last draw
__device__ __forceinline__ void epoch_arrive(const StepParams &P) {
    __threadfence();
    if (atomicAdd(reinterpret_cast<unsigned long long *>(P.epoch) + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
        P.epoch[1] = 0;
        P.epoch[0] = P.epoch[0] + 1;
    }
}

template <int V>
__device__ __forceinline__ void fill_reset(const StepParams &P, const uint4 (&r)[6], EnvState<V> &e) {
    const float pi = 3.14159265358979f, pi9 = 0.349065850398866f;
    e.x = P.rd.start[0] + uni(r[0].x, -0.5f, 1.0f);
    e.y = P.rd.start[1] + uni(r[0].y, -0.5f, 1.0f);
    e.z = P.rd.start[2] + uni(r[0].z, -0.5f, 1.0f);
    e.vx = uni(r[0].w, -0.5f, 1.0f); e.vy = uni(r[1].x, -0.5f, 1.0f); e.vz = uni(r[1].y, -0.5f, 1.0f);
    e.phi = uni(r[1].z, -pi9, 2 * pi9); e.th = uni(r[1].w, -pi9, 2 * pi9); e.psi = uni(r[2].x, -pi, 2 * pi);
    e.p = uni(r[2].y, -0.1f, 0.2f); e.q = uni(r[2].z, -0.1f, 0.2f); e.r = uni(r[2].w, -0.1f, 0.2f);
    if (V == kE2E) {
        e.w[0] = uni(r[3].x, -1.f, 2.f); e.w[1] = uni(r[3].y, -1.f, 2.f);
        e.w[2] = uni(r[3].z, -1.f, 2.f); e.w[V == kE2E ? 3 : 0] = uni(r[3].w, -1.f, 2.f);
        e.dist[0] = uni(r[4].x, P.rd.dist_lo[0], P.rd.dist_span[0]);
        e.dist[1] = uni(r[4].y, P.rd.dist_lo[1], P.rd.dist_span[1]);
        e.dist[2] = uni(r[4].z, P.rd.dist_lo[2], P.rd.dist_span[2]);
        e.dist[3] = uni(r[4].w, P.rd.dist_lo[3], P.rd.dist_span[3]);
        e.dist[4] = uni(r[5].x, P.rd.dist_lo[4], P.rd.dist_span[4]);
        e.dist[5] = uni(r[5].y, P.rd.dist_lo[5], P.rd.dist_span[5]);
    } else {
        e.w[0] = uni(r[3].x, -0.1f, 0.2f);
    }
}

// every lane draws for itself (used by reset(), and by the step when many lanes of a warp terminate at once)
template <int V>
__device__ __forceinline__ void draw_reset(const StepParams &P, long long env, EnvState<V> &e, unsigned long long epoch) {
    const unsigned long long g = (unsigned long long)(env + P.env_offset);
    uint4 r[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) r[c] = c < ResetDraws<V>::N ? reset_draw(P, g, c, epoch) : make_uint4(0, 0, 0, 0);
    fill_reset<V>(P, r, e);
}

// Warp-cooperative form for the common case of a few terminating lanes per warp: instead of one lane running 6
// Philox evaluations while 31 wait, lane L evaluates block L % ND for the (L / ND)-th terminating lane -- ONE
// Philox pass serves up to 32/ND resets -- and the words change hands through 512 bytes of the warp's own shared
// memory (`scratch`, 16-byte aligned).  Must be called by the whole warp with warp_env0 = env index of lane 0;
// `need` marks the lanes that reset.  Same values as draw_reset.
template <int V>
__device__ __forceinline__ void draw_reset_warp(const StepParams &P, long long warp_env0, bool need, uint4 *scratch,
                                                EnvState<V> &e, unsigned long long epoch) {
    constexpr int ND = ResetDraws<V>::N;
    const unsigned full = 0xffffffffu;
    const unsigned m = __ballot_sync(full, need);
    if (m == 0) return;
    const unsigned lane = threadIdx.x & 31;
    const int cnt = __popc(m);
    if (cnt * ND > 32) {  // mass termination (e.g. synchronous time-outs): per-lane is cheaper
        if (need) draw_reset<V>(P, warp_env0 + lane, e, epoch);
        return;
    }
    const int which = lane / ND, c = lane - which * ND;
    unsigned mm = m;  // lane index of the `which`-th terminating lane
#pragma unroll
    for (int k = 0; k < 32 / ND - 1; ++k) mm = (k < which) ? (mm & (mm - 1)) : mm;
    if (which < cnt) {
        const int src = __ffs(mm) - 1;
        scratch[lane] = reset_draw(P, (unsigned long long)(warp_env0 + src + P.env_offset), c, epoch);
    }
    __syncwarp();
    if (need) {
        const int rank = __popc(m & ((1u << lane) - 1u));
        uint4 r[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) r[k] = k < ND ? scratch[rank * ND + k] : make_uint4(0, 0, 0, 0);
        fill_reset<V>(P, r, e);
    }
    __syncwarp();
}

// update_states_gate for one env (`:365-450`): the gate-frame part of the row (position, horizontal velocity, yaw)
struct ObsHead { float o0, o1, o2, o3, o4, yaw; };
template <int V>
__device__ __forceinline__ ObsHead obs_head(const float *s_track, const EnvState<V> &e, uint32_t tg) {
    const float4 ga = *reinterpret_cast<const float4 *>(s_track + tg * kTrackRow);
    const float2 gb = *reinterpret_cast<const float2 *>(s_track + tg * kTrackRow + 4);
    const float c = gb.x, sn = gb.y;
    const float dx = sub_rn(e.x, ga.x), dy = sub_rn(e.y, ga.y);
    ObsHead h;
    h.o0 = add_rn(mul_rn(dx, c), mul_rn(dy, sn));
    h.o1 = add_rn(mul_rn(dx, -sn), mul_rn(dy, c));
    h.o2 = sub_rn(e.z, ga.z);
    h.o3 = add_rn(mul_rn(e.vx, c), mul_rn(e.vy, sn));
    h.o4 = add_rn(mul_rn(e.vx, -sn), mul_rn(e.vy, c));
    h.yaw = wrap_yaw(sub_rn(e.psi, ga.w));
    return h;
}
__device__ __forceinline__ float dist_obs(const StepParams &P, float d, int k) {
    return fmaf(sub_rn(sub_rn(d, P.obs_lo[k]), P.obs_lo2[k]), P.obs_scale[k], -1.0f);
}
// ... written to a row in shared memory (16-byte aligned for E2E: obs_len is a multiple of 4 there)
template <int V>
__device__ __forceinline__ void write_obs(const StepParams &P, const float *s_track, const EnvState<V> &e,
                                          uint32_t tg, float *o) {
    const int ng = P.n_gates;
    const ObsHead h = obs_head<V>(s_track, e, tg);
    if (V == kE2E) {
        float4 *o4p = reinterpret_cast<float4 *>(o);
        o4p[0] = make_float4(h.o0, h.o1, h.o2, h.o3);
        o4p[1] = make_float4(h.o4, e.vz, e.phi, e.th);
        o4p[2] = make_float4(h.yaw, e.p, e.q, e.r);
        o4p[3] = make_float4(e.w[0], e.w[1], e.w[2], e.w[V == kE2E ? 3 : 0]);
        uint32_t nx = tg;
#pragma unroll 1
        for (int i = 0; i < P.gates_ahead; ++i) {
            nx = (nx + 1 == (uint32_t)ng) ? 0u : nx + 1;
            o4p[4 + i] = *reinterpret_cast<const float4 *>(s_track + nx * kTrackRow + 8);
        }
        o4p[4 + P.gates_ahead] = make_float4(dist_obs(P, e.dist[0], 0), dist_obs(P, e.dist[1], 1), dist_obs(P, e.dist[2], 2),
                                             dist_obs(P, e.dist[5], 3));
    } else {
        o[0] = h.o0; o[1] = h.o1; o[2] = h.o2; o[3] = h.o3; o[4] = h.o4; o[5] = e.vz; o[6] = e.phi; o[7] = e.th;
        o[8] = h.yaw; o[9] = e.p; o[10] = e.q; o[11] = e.r; o[12] = e.w[0];
        uint32_t nx = tg;
#pragma unroll 1
        for (int i = 0; i < P.gates_ahead; ++i) {
            nx = (nx + 1 == (uint32_t)ng) ? 0u : nx + 1;
            const float4 rel = *reinterpret_cast<const float4 *>(s_track + nx * kTrackRow + 8);
            o[13 + 4 * i] = rel.x; o[14 + 4 * i] = rel.y; o[15 + 4 * i] = rel.z; o[16 + 4 * i] = rel.w;
        }
    }
}
// The block's observation tile [rows][obs_len] is contiguous in global memory: one TMA bulk store moves it.
__device__ __forceinline__ void store_obs_tile(float *dst, const float *s_obs, int rows, int obs_len) {
    const uint32_t bytes = (uint32_t)rows * (uint32_t)obs_len * 4u;
    const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
    if (bulk) {
        // make this thread's generic-proxy smem writes visible to the async proxy, then one thread issues the copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_obs);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the read
        }
    } else {
        __syncthreads();
        for (int i = threadIdx.x; i < rows * obs_len; i += kBlock) dst[i] = s_obs[i];
    }
}

// ---- packed observations: what the step hands to the on-device policy when nobody needs float32 rows (a sharded job
// whose policy reads the all-gathered observations, BASELINE.json config 4).  One block per 32 envs:
//   [K chunk c = 0 .. chunks-1][row r = 0..31][8 x BF16]     = the policy kernel's K-major A-operand slabs, first layer
// The operand has kPackK = 32 columns: obs_len <= 31 values, the constant 1 that multiplies the folded bias at column
// obs_len, zeros after it.  Only the chunks that carry an observation value travel: chunks = ceil(obs_len / 8) -- the
// E2E row of 24 values is 3 chunks = 48 B per env instead of 96 B of float32, half the bytes through NVLink; a chunk that
// would hold nothing but the constant (obs_len % 8 == 0) is written by the policy kernel itself.  The policy loads
// the blocks with TMA straight into the MMA's operand buffer (no load / convert / store instructions at all).
constexpr int kPackK = 32;
constexpr int kPackChunkBytes = 32 * 16;  // one K chunk of a block: 32 rows x 8 BF16
__host__ __device__ constexpr int pack_chunks(int obs_len) { return (obs_len + 7) / 8; }
__host__ __device__ constexpr int pack_block_bytes(int obs_len) { return pack_chunks(obs_len) * kPackChunkBytes; }
__device__ __forceinline__ uint32_t pack_bf16_rn(float lo, float hi) {  // round-to-nearest-even, lo in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// this thread's float32 row (in shared memory) -> 16 BF16 pairs; inactive rows pack to zero
__device__ __forceinline__ void pack_obs_row(const float *row, int obs_len, bool active, uint32_t (&w)[16]) {
    float x[kPackK];
    if ((obs_len & 3) == 0) {
#pragma unroll
        for (int q = 0; q < kPackK / 4; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * q < obs_len) v = *reinterpret_cast<const float4 *>(row + 4 * q);
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kPackK; ++k) x[k] = k < obs_len ? row[k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kPackK; ++k) x[k] = active ? (k == obs_len ? 1.0f : x[k]) : 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) w[c] = pack_bf16_rn(x[2 * c], x[2 * c + 1]);
}

__device__ __forceinline__ void load_track(const StepParams &P, float *s_track) {
    for (int i = threadIdx.x; i < P.n_gates * kTrackRow; i += kBlock) s_track[i] = P.track[i];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ dynamics
// packed FP32: one FFMA2 issues two FMAs (Blackwell `fma.rn.f32x2`); operands are 64-bit register pairs,
// a uniform-register pair straight from the constant bank, or a scalar broadcast to both halves.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// thrust_moment_model_world_states (`:254-262`): two Linear-ReLU-Linear nets sharing their first 7 inputs.
// Weights are kernel parameters (constant bank, warp-uniform): LDCU.128 brings four of them into uniform
// registers and each FFMA2 consumes a PAIR of hidden units -- 336 FFMA2 + 168 LDCU.128 for the 672 MACs,
// measured on B200 at the same FMA/clk as scalar FFMA in half the issue slots (profiles/microbench).
// The hidden-unit loops stay FULLY unrolled although that makes the E2E step kernel 4000 SASS instructions long:
// rolled (8 trips, weights indexed by the trip counter) the constant-bank operands become indexed loads and the step
// takes 227 us instead of 60.6 us at N = 2^20 (measured, profiles/r1/experiments.md).
#ifndef QS_MLP_CHAINS
#define QS_MLP_CHAINS 4  // independent FFMA2 accumulator chains per trip (2 or 4); same sums, same order, same bits
#endif
__device__ __forceinline__ void residual_mlp(const StepParams &P, const float (&x)[10], float &thrust, float (&mom)[3]) {
    constexpr int C = QS_MLP_CHAINS;  // hidden-unit pairs in flight: a dependent FFMA2 issues every ~4 cycles, so
                                      // 2 chains cap a warp at IPC 0.5 inside the MLP; 4 chains double its ILP
    f32x2 xx[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) xx[k] = pack2(x[k], x[k]);
    f32x2 at = pack2(P.b2[0], 0.0f);
#pragma unroll
    for (int j = 0; j < 32; j += 2 * C) {
        f32x2 h[C];
#pragma unroll
        for (int c = 0; c < C; ++c) h[c] = pack2(P.bt1[j + 2 * c], P.bt1[j + 2 * c + 1]);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                h[c] = fma2(pack2(P.wt1[k * 32 + j + 2 * c], P.wt1[k * 32 + j + 2 * c + 1]), xx[k], h[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {  // ascending hidden units into ONE output chain: the order fixes the bits
            float h0, h1;
            unpack2(h[c], h0, h1);
            at = fma2(pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f)), pack2(P.wt2[j + 2 * c], P.wt2[j + 2 * c + 1]), at);
        }
    }
    float lo, hi;
    unpack2(at, lo, hi);
    thrust = add_rn(lo, hi);
    f32x2 a0 = pack2(P.b2[1], 0.0f), a1 = pack2(P.b2[2], 0.0f), a2 = pack2(P.b2[3], 0.0f);
#pragma unroll
    for (int j = 0; j < 32; j += 2 * C) {
        f32x2 h[C];
#pragma unroll
        for (int c = 0; c < C; ++c) h[c] = pack2(P.bm1[j + 2 * c], P.bm1[j + 2 * c + 1]);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                h[c] = fma2(pack2(P.wm1[k * 32 + j + 2 * c], P.wm1[k * 32 + j + 2 * c + 1]), xx[k], h[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float h0, h1;
            unpack2(h[c], h0, h1);
            const f32x2 r = pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f));
            a0 = fma2(r, pack2(P.wm2[j + 2 * c], P.wm2[j + 2 * c + 1]), a0);
            a1 = fma2(r, pack2(P.wm2[32 + j + 2 * c], P.wm2[32 + j + 2 * c + 1]), a1);
            a2 = fma2(r, pack2(P.wm2[64 + j + 2 * c], P.wm2[64 + j + 2 * c + 1]), a2);
        }
    }
    unpack2(a0, lo, hi); mom[0] = add_rn(lo, hi);
    unpack2(a1, lo, hi); mom[1] = add_rn(lo, hi);
    unpack2(a2, lo, hi); mom[2] = add_rn(lo, hi);
}

// The same two nets for TWO envs of one thread (step_kernel_x2): every weight pair is fetched once and used for both,
// which halves the LDCU / LDC traffic per env; per env the operations and their order are exactly residual_mlp's.
__device__ __forceinline__ void residual_mlp2(const StepParams &P, const float (&xa)[10], const float (&xb)[10], float &thrust_a,
                                              float (&mom_a)[3], float &thrust_b, float (&mom_b)[3]) {
    constexpr int C = QS_MLP_CHAINS;
    f32x2 xxa[10], xxb[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) { xxa[k] = pack2(xa[k], xa[k]); xxb[k] = pack2(xb[k], xb[k]); }
    f32x2 ata = pack2(P.b2[0], 0.0f), atb = ata;
#pragma unroll
    for (int j = 0; j < 32; j += 2 * C) {
        f32x2 ha[C], hb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ha[c] = hb[c] = pack2(P.bt1[j + 2 * c], P.bt1[j + 2 * c + 1]);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const f32x2 w = pack2(P.wt1[k * 32 + j + 2 * c], P.wt1[k * 32 + j + 2 * c + 1]);
                ha[c] = fma2(w, xxa[k], ha[c]);
                hb[c] = fma2(w, xxb[k], hb[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const f32x2 w = pack2(P.wt2[j + 2 * c], P.wt2[j + 2 * c + 1]);
            float h0, h1;
            unpack2(ha[c], h0, h1);
            ata = fma2(pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f)), w, ata);
            unpack2(hb[c], h0, h1);
            atb = fma2(pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f)), w, atb);
        }
    }
    float lo, hi;
    unpack2(ata, lo, hi); thrust_a = add_rn(lo, hi);
    unpack2(atb, lo, hi); thrust_b = add_rn(lo, hi);
    f32x2 a0 = pack2(P.b2[1], 0.0f), a1 = pack2(P.b2[2], 0.0f), a2 = pack2(P.b2[3], 0.0f), b0 = a0, b1 = a1, b2 = a2;
#pragma unroll
    for (int j = 0; j < 32; j += 2 * C) {
        f32x2 ha[C], hb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ha[c] = hb[c] = pack2(P.bm1[j + 2 * c], P.bm1[j + 2 * c + 1]);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const f32x2 w = pack2(P.wm1[k * 32 + j + 2 * c], P.wm1[k * 32 + j + 2 * c + 1]);
                ha[c] = fma2(w, xxa[k], ha[c]);
                hb[c] = fma2(w, xxb[k], hb[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const f32x2 w0 = pack2(P.wm2[j + 2 * c], P.wm2[j + 2 * c + 1]);
            const f32x2 w1 = pack2(P.wm2[32 + j + 2 * c], P.wm2[32 + j + 2 * c + 1]);
            const f32x2 w2 = pack2(P.wm2[64 + j + 2 * c], P.wm2[64 + j + 2 * c + 1]);
            float h0, h1;
            unpack2(ha[c], h0, h1);
            const f32x2 ra = pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f));
            a0 = fma2(ra, w0, a0); a1 = fma2(ra, w1, a1); a2 = fma2(ra, w2, a2);
            unpack2(hb[c], h0, h1);
            const f32x2 rb = pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f));
            b0 = fma2(rb, w0, b0); b1 = fma2(rb, w1, b1); b2 = fma2(rb, w2, b2);
        }
    }
    unpack2(a0, lo, hi); mom_a[0] = add_rn(lo, hi);
    unpack2(a1, lo, hi); mom_a[1] = add_rn(lo, hi);
    unpack2(a2, lo, hi); mom_a[2] = add_rn(lo, hi);
    unpack2(b0, lo, hi); mom_b[0] = add_rn(lo, hi);
    unpack2(b1, lo, hi); mom_b[1] = add_rn(lo, hi);
    unpack2(b2, lo, hi); mom_b[2] = add_rn(lo, hi);
}

// sin and cos together: 3-term Cody-Waite reduction by pi/2 and degree-7/8 minimax polynomials (Cephes
// coefficients), <= 1.5 ulp on |x| < 1e5 -- the same error class as NumPy's float32 sin/cos (1.45 ulp), see
// tests/test_kernel_math_models.py.  Huge arguments (a blown-up yaw) take the library's Payne-Hanek path
// out of line, so the hot code stays small and needs no stack frame.
__device__ __noinline__ float2 sincos_slow(float x) {
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

__device__ __forceinline__ void sincos_core(float x, float &sn, float &cs) {  // |x| <= 1e5 (anything else: see callers)
    const float j = rintf(mul_rn(x, 0.636619772367581343f));
    const int q = (int)j;
    float a = fmaf(j, -1.5707962512969970703f, x);
    a = fmaf(j, -7.5497894158615963534e-08f, a);
    a = fmaf(j, -5.3903029534742383927e-15f, a);
    const float z = mul_rn(a, a);
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float s = fmaf(mul_rn(a, z), ps, a);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    pc = fmaf(pc, z, -0.5f);
    const float c = fmaf(pc, z, 1.0f);
    const bool swap = q & 1;
    float ss = swap ? c : s, cc = swap ? s : c;
    sn = (q & 2) ? -ss : ss;
    cs = ((q + 1) & 2) ? -cc : cc;
}
__device__ __forceinline__ void sincos_fast(float x, float &sn, float &cs) {
    if (fabsf(x) > 1.0e5f) { const float2 r = sincos_slow(x); sn = r.x; cs = r.y; return; }   // also Inf
    sincos_core(x, sn, cs);
}
// The three Euler angles of an env at once: the polynomial path runs unconditionally for all three -- ONE basic block,
// so the scheduler interleaves the three dependency chains -- and a single, rarely taken branch redoes the angles whose
// magnitude needs the Payne-Hanek path.  Same values as three sincos_fast calls.
__device__ __forceinline__ void sincos_fix(float x, float &sn, float &cs) {
    if (fabsf(x) > 1.0e5f) { const float2 r = sincos_slow(x); sn = r.x; cs = r.y; }
}
__device__ __forceinline__ void sincos3_fast(float a, float b, float c, float &sa, float &ca, float &sb, float &cb, float &sc,
                                             float &cc) {
    sincos_core(a, sa, ca);
    sincos_core(b, sb, cb);
    sincos_core(c, sc, cc);
    if (fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(c)) > 1.0e5f) {  // fmaxf drops NaNs, like the three separate tests
        sincos_fix(a, sa, ca); sincos_fix(b, sb, cb); sincos_fix(c, sc, cc);
    }
}

// new = state + dt * f(state, action[, residual + disturbance])  (`:503-512`; INDI `:304`)
// Every operation is spelled out (fmaf / __fmul_rn / __fadd_rn / __fsub_rn): nothing is left for the compiler to
// contract one way in one kernel and another way in the next, so step_kernel and the fused rollout_kernel, which both
// inline this function, produce the same bits.  The FMA placement is the one the compiler chose when free to (one
// multiply-add per product term), measured <= 2e-6 scaled error against the reference's unfused float32 evaluation.
// The three stages of one Euler step.  euler_pre: rotation matrix, body velocity and the residual nets' input vector;
// [residual nets: residual_mlp for one env, residual_mlp2 for the two envs of a step_kernel_x2 thread]; euler_post:
// equations of motion and the update.  euler_step = the three in a row; both kernels execute the same operations.
struct EulerMid {
    float sph, cph, sth, cth;
    float r00, r10, r20, r01, r11, r21, r02, r12, r22;
    float vbx, vby, vbz;
};
template <int V>
__device__ __forceinline__ void euler_pre(const EnvState<V> &e, EulerMid &m, float (&x)[10]) {
    float sps, cps;
    sincos3_fast(e.phi, e.th, e.psi, m.sph, m.cph, m.sth, m.cth, sps, cps);
    const float sph = m.sph, cph = m.cph, sth = m.sth, cth = m.cth;
    // R = Rz*Ry*Rx
    m.r00 = mul_rn(cps, cth); m.r10 = mul_rn(sps, cth); m.r20 = -sth;
    m.r01 = fmaf(mul_rn(sph, sth), cps, -mul_rn(sps, cph));   // sph*sth*cps - sps*cph
    m.r11 = fmaf(mul_rn(sph, sps), sth, mul_rn(cph, cps));    // sph*sps*sth + cph*cps
    m.r21 = mul_rn(sph, cth);
    m.r02 = fmaf(mul_rn(sth, cph), cps, mul_rn(sph, sps));    // sph*sps + sth*cph*cps
    m.r12 = fmaf(mul_rn(sps, sth), cph, -mul_rn(sph, cps));   // -sph*cps + sps*sth*cph
    m.r22 = mul_rn(cph, cth);
    m.vbx = fmaf(e.vz, m.r20, fmaf(e.vy, m.r10, mul_rn(e.vx, m.r00)));
    m.vby = fmaf(e.vz, m.r21, fmaf(e.vy, m.r11, mul_rn(e.vx, m.r01)));
    m.vbz = fmaf(e.vz, m.r22, fmaf(e.vy, m.r12, mul_rn(e.vx, m.r02)));
    if (V == kE2E) {
        x[0] = e.w[0]; x[1] = e.w[1]; x[2] = e.w[2]; x[3] = e.w[V == kE2E ? 3 : 0];
        x[4] = m.vbx; x[5] = m.vby; x[6] = m.vbz; x[7] = e.p; x[8] = e.q; x[9] = e.r;
    }
}

template <int V>
__device__ __forceinline__ void euler_post(const StepParams &P, const EnvState<V> &e, const float4 u, const EulerMid &m,
                                           const float thr, const float (&mom)[3], EnvState<V> &n) {
    const float dt = P.dt;
    const float sph = m.sph, cph = m.cph, sth = m.sth, cth = m.cth;
    const float r00 = m.r00, r10 = m.r10, r20 = m.r20, r01 = m.r01, r11 = m.r11, r21 = m.r21, r02 = m.r02, r12 = m.r12, r22 = m.r22;
    const float vbx = m.vbx, vby = m.vby, vbz = m.vbz;
    float Dx, Dy, T, dp, dq, dr;
    if (V == kE2E) {
        const float w1 = e.w[0], w2 = e.w[1], w3 = e.w[2], w4 = e.w[V == kE2E ? 3 : 0];
        const float Mx = add_rn(mom[0], e.dist[0]), My = add_rn(mom[1], e.dist[1]), Mz = add_rn(mom[2], e.dist[2]);
        const float Fz = add_rn(thr, e.dist[5]);
        const float W1 = fmaf(4000.f, w1, 7000.f), W2 = fmaf(4000.f, w2, 7000.f);
        const float W3 = fmaf(4000.f, w3, 7000.f), W4 = fmaf(4000.f, w4, 7000.f);
        const float sumW = add_rn(add_rn(W1, W2), add_rn(W3, W4));
        const float q1 = mul_rn(W1, W1), q2 = mul_rn(W2, W2), q3 = mul_rn(W3, W3), q4 = mul_rn(W4, W4);
        // T = Fz - k_w*sum(W^2) - k_h*(vbx^2 + vby^2) - k_z*vbz*sum(W)
        T = fmaf(-4.36301076e-8f, add_rn(add_rn(q1, q2), add_rn(q3, q4)), Fz);
        T = fmaf(-0.0625501332f, fmaf(vbx, vbx, mul_rn(vby, vby)), T);
        T = fmaf(-mul_rn(2.7862899e-5f, vbz), sumW, T);
        Dx = fmaf(-mul_rn(1.07933887e-5f, vbx), sumW, e.dist[3]);
        Dy = fmaf(-mul_rn(9.65250793e-6f, vby), sumW, e.dist[4]);
        dp = mul_rn(1103.7527593819f, Mx);
        dp = fmaf(-mul_rn(0.896247240618101f, e.q), e.r, dp);
        dp = fmaf(-8.79803364238411f, vby, dp);
        dp = fmaf(1.55842505518764e-6f, add_rn(sub_rn(q1, q2), sub_rn(q4, q3)), dp);
        dq = mul_rn(805.152979066023f, My);
        dq = fmaf(mul_rn(0.924315619967794f, e.p), e.r, dq);
        dq = fmaf(10.4077084541063f, vbx, dq);
        dq = fmaf(9.79081191626409e-7f, add_rn(sub_rn(q1, q3), sub_rn(q2, q4)), dq);
        dr = mul_rn(486.854917234664f, Mz);
        dr = fmaf(-mul_rn(0.163583252190847f, e.p), e.q, dr);
        dr = fmaf(-0.395780237098345f, e.r, dr);
        dr = fmaf(13.3373373580007f, add_rn(sub_rn(u.y, u.x), sub_rn(u.w, u.z)), dr);
        dr = fmaf(8.33177659850698f, add_rn(sub_rn(w1, w2), sub_rn(w3, w4)), dr);
        n.w[0] = fmaf(dt, mul_rn(16.6666666666667f, sub_rn(u.x, w1)), w1);
        n.w[1] = fmaf(dt, mul_rn(16.6666666666667f, sub_rn(u.y, w2)), w2);
        n.w[2] = fmaf(dt, mul_rn(16.6666666666667f, sub_rn(u.z, w3)), w3);
        n.w[V == kE2E ? 3 : 0] = fmaf(dt, mul_rn(16.6666666666667f, sub_rn(u.w, w4)), w4);
#pragma unroll
        for (int k = 0; k < 6; ++k) n.dist[k] = e.dist[k];
    } else {
        const float Tn = e.w[0];
        T = fmaf(-8.0f, Tn, -8.0f);
        Dx = mul_rn(-0.33915248f, vbx);
        Dy = mul_rn(-0.4314916f, vby);
        dp = fmaf(-33.3333333333333f, e.p, mul_rn(100.0f, u.x));
        dq = fmaf(-33.3333333333333f, e.q, mul_rn(100.0f, u.y));
        dr = fmaf(-33.3333333333333f, e.r, mul_rn(66.6666666666667f, u.z));
        n.w[0] = fmaf(dt, mul_rn(33.3333333333333f, sub_rn(u.w, Tn)), Tn);
    }
    const float dvx = fmaf(r02, T, fmaf(r01, Dy, mul_rn(r00, Dx)));
    const float dvy = fmaf(r12, T, fmaf(r11, Dy, mul_rn(r10, Dx)));
    const float dvz = add_rn(fmaf(r22, T, fmaf(r21, Dy, mul_rn(r20, Dx))), 9.81f);
    // Euler-angle kinematics (`:140-142`); tan = sin/cos with one IEEE reciprocal
    const float rc = __frcp_rn(cth);
    const float tth = mul_rn(sth, rc);
    const float qs_rc = fmaf(e.r, cph, mul_rn(e.q, sph));
    const float dphi = fmaf(qs_rc, tth, e.p);
    const float dth = fmaf(-e.r, sph, mul_rn(e.q, cph));
    const float dpsi = mul_rn(qs_rc, rc);

    // positions: exactly the reference's two rounded float32 operations, so every threshold test agrees bit for bit
    n.x = add_rn(e.x, mul_rn(dt, e.vx));
    n.y = add_rn(e.y, mul_rn(dt, e.vy));
    n.z = add_rn(e.z, mul_rn(dt, e.vz));
    n.vx = fmaf(dt, dvx, e.vx); n.vy = fmaf(dt, dvy, e.vy); n.vz = fmaf(dt, dvz, e.vz);
    n.phi = fmaf(dt, dphi, e.phi); n.th = fmaf(dt, dth, e.th); n.psi = fmaf(dt, dpsi, e.psi);
    n.p = fmaf(dt, dp, e.p); n.q = fmaf(dt, dq, e.q); n.r = fmaf(dt, dr, e.r);
}

template <int V>
__device__ __forceinline__ void euler_step(const StepParams &P, const EnvState<V> &e, const float4 u, EnvState<V> &n) {
    EulerMid m;
    float x[10];
    euler_pre<V>(e, m, x);
    float thr = 0.0f, mom[3] = {0.0f, 0.0f, 0.0f};
    if (V == kE2E) {
#ifdef QS_EXP_NOMLP  // experiment only: what the step costs without the residual MLPs
        thr = mul_rn(x[4], P.b2[0]); mom[0] = mul_rn(x[5], P.b2[1]); mom[1] = mul_rn(x[6], P.b2[2]); mom[2] = mul_rn(x[7], P.b2[3]);
#else
        residual_mlp(P, x, thr, mom);
#endif
    }
    euler_post<V>(P, e, u, m, thr, mom, n);
}

// ------------------------------------------------------------------------------------------------ async-proxy primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// programmatic dependent launch: let the next step's CTAs be scheduled while this grid drains / wait for the
// previous grid's memory before touching simulator state
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// ---- L2 residency control (126 MB L2 on B200).  The simulator state is re-read and re-written every step while
// observations / rewards / flags stream out and actions stream in: with the state's lines marked evict_last and the
// streams evict_first the state stays resident in L2 across consecutive step launches and stops costing HBM traffic.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_hint(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void *dst, const void *src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void st_hint(float4 *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float2 *p, float2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(uint8_t *p, uint8_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(p), "r"((uint32_t)v), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// step_counts += 1, reward, gate logic and termination flags of one env (`3D quad race.ipynb:514-566`): `e` is the
// state before the Euler update, `n` after it.  Advances `tg` on a gate pass; shared by the step kernel and the fused
// rollout kernel so that both produce the same bits.
template <int V>
__device__ __forceinline__ void reward_and_flags(const StepParams &P, const float *s_track, const EnvState<V> &e,
                                                 const EnvState<V> &n, uint32_t &tg, uint32_t &sc, float &reward, bool &dn,
                                                 uint32_t &fl) {
    sc = sc < kStepMask ? sc + 1 : sc;  // step_counts += 1 (`:514`), saturating in 24 bits
    const float4 ga = *reinterpret_cast<const float4 *>(s_track + tg * kTrackRow);
    const float2 gcs = *reinterpret_cast<const float2 *>(s_track + tg * kTrackRow + 4);
    const float ox = sub_rn(e.x, ga.x), oy = sub_rn(e.y, ga.y), oz = sub_rn(e.z, ga.z);
    const float nx = sub_rn(n.x, ga.x), ny = sub_rn(n.y, ga.y), nz = sub_rn(n.z, ga.z);
    const float d_old = norm3_rn(ox, oy, oz), d_new = norm3_rn(nx, ny, nz);
    reward = sub_rn(d_old, d_new);
    const float proj_old = add_rn(mul_rn(ox, gcs.x), mul_rn(oy, gcs.y));
    const float proj_new = add_rn(mul_rn(nx, gcs.x), mul_rn(ny, gcs.y));
    const bool plane = (proj_old < 0.0f) && (proj_new > 0.0f);
    const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
    const bool passed = plane && (ax < 0.5f) && (ay < 0.5f) && (az < 0.5f);
    const bool collided = plane && ((ax > 0.5f) || (ay > 0.5f) || (az > 0.5f));
    const bool ground = n.z > 0.0f;
    const bool oob = (fabsf(n.x) > 10.0f) || (fabsf(n.y) > 10.0f) || (fabsf(n.p) > 1000.0f) ||
                     (fabsf(n.q) > 1000.0f) || (fabsf(n.r) > 1000.0f);
    const bool trunc = sc >= P.max_steps;
    if (passed) reward = sub_rn(10.0f, mul_rn(10.0f, d_new));
    if (collided | ground | oob) reward = -10.0f;
    if (passed) tg = (tg + 1 == (uint32_t)P.n_gates) ? 0u : tg + 1;  // (`:556-557`)
    dn = trunc | ground | collided | oob;
    fl = (dn ? F_DONE : 0u) | (trunc ? F_TRUNC : 0u) | (passed ? F_PASSED : 0u) | (collided ? F_COLLISION : 0u) |
         (ground ? F_GROUND : 0u) | (oob ? F_OOB : 0u);
}

// ------------------------------------------------------------------------------------------------ the step kernel
// Input stage of ONE WARP: the warp's state block exactly as it lies in HBM (one bulk copy) followed by its 32
// action rows (a second bulk copy from the caller's buffer); nobody spends an instruction on global loads.
template <int V> struct Stage : Blk<V> {
    enum : int { ACT = Blk<V>::BYTES, BYTES = Blk<V>::BYTES + 512 };
};
constexpr int kBarBytes = 256;  // mbarriers live in the first 256 bytes of dynamic shared memory
// Worker warps per CTA of the step kernel.  A warp is an independent worker over 32-env warp-tiles, so the CTA size only
// decides how many warps fit on an SM.  E2E (91 registers): 5 CTAs x 4 warps = 20 warps per SM.  Measured alternative
// (QS_E2E_WARPS=7, 3 CTAs x 7 warps = 21 per SM, 148 x 21 workers would step 2^20 envs in 11 rounds instead of 12):
// registers are allocated to a CTA in units of 4 warps, so 7-warp CTAs only fit with 80 registers per thread and the
// tighter allocation costs more than the 21st warp gives back (57.8 vs 55.7 us per step at N = 2^20).
// INDI (63 registers): 8 CTAs x 4 warps = 32 warps per SM.
#ifndef QS_E2E_WARPS
#define QS_E2E_WARPS 4
#endif
template <int V> struct StepCta {
    enum : int { WARPS = (V == kE2E ? QS_E2E_WARPS : 4), THREADS = WARPS * 32,
                 MIN_CTAS = (V == kE2E ? (QS_E2E_WARPS == 7 ? 3 : QS_E2E_WARPS == 4 ? 5 : QS_E2E_WARPS == 2 ? 10 : 1) : 8) };
};
__host__ __device__ constexpr int step_warps(int variant) { return variant == kE2E ? (int)StepCta<kE2E>::WARPS : (int)StepCta<kINDI>::WARPS; }

// a warp's observation staging slice: its 32 float32 rows, and never less than the largest packed BF16 block (2 KB)
__host__ __device__ constexpr int step_slice_floats(int obs_len) { return 32 * obs_len > 512 ? 32 * obs_len : 512; }
__host__ __device__ constexpr size_t step_smem_bytes(int variant, int stages, int obs_len, int n_gates) {
    return kBarBytes + (size_t)stages * step_warps(variant) * (variant == kE2E ? (int)Stage<kE2E>::BYTES : (int)Stage<kINDI>::BYTES) +
           (size_t)step_warps(variant) * step_slice_floats(obs_len) * 4 + (size_t)n_gates * kTrackRow * 4;
}

// lane 0 of a warp: fill one of the warp's stages with the 32 envs starting at `first`
template <int V, bool kHints>
__device__ __forceinline__ void issue_warp_tile(const StepParams &P, unsigned char *st, uint64_t *bar, long long first,
                                                uint64_t pol_state, uint64_t pol_stream) {
    using S = Stage<V>;
    const long long rem = P.n - first;
    const uint32_t act_bytes = rem >= 32 ? 512u : (rem > 0 ? (uint32_t)rem * 16u : 0u);  // caller's buffer is not padded
    mbar_expect_tx(bar, (uint32_t)Blk<V>::BYTES + act_bytes);
    if (kHints) {
        bulk_load_hint(st, P.s.base + (first >> 5) * (long long)Blk<V>::BYTES, Blk<V>::BYTES, bar,
                       (first >> 5) < P.keep_blocks ? pol_state : pol_stream);
        if (act_bytes) bulk_load_hint(st + S::ACT, P.actions + first, act_bytes, bar, pol_stream);
    } else {
        bulk_load(st, P.s.base + (first >> 5) * (long long)Blk<V>::BYTES, Blk<V>::BYTES, bar);
        if (act_bytes) bulk_load(st + S::ACT, P.actions + first, act_bytes, bar);
    }
}

// Start of a step launch, after the CTA's own prologue (barrier init, track table): wait for what this CTA depends on
// and return the launch's RNG epoch.  Classic launch: trigger the next grid's early start, wait for the whole previous
// grid (griddepcontrol.wait), read the epoch (advanced by the last CTA of this launch to FINISH, step_launch_end).
// Chained launch (P.chain): the epoch is read and advanced -- by the last CTA to ARRIVE -- BEFORE this CTA lets the
// next launch start, so every CTA of the next launch reads epoch + 1 although this launch is still running; then only
// CTA b of the previous launch is awaited (it stepped the very tiles this CTA is about to load).  chain_seq (thread 0):
// how many chained launches this CTA slot had started before this one.
__device__ __forceinline__ unsigned long long step_launch_begin(const StepParams &P, unsigned &chain_seq) {
    const int tid = threadIdx.x;
    chain_seq = 0;
    if (P.chain == nullptr) {
        pdl_launch_dependents();
        pdl_wait();
        return load_epoch(P.epoch);
    }
    __shared__ unsigned long long s_epoch;
    if (!P.chain_wait) pdl_wait();  // the previous kernel is not a chained step of this env: wait for all of it
    if (tid == 0) {
        const unsigned long long ep = load_epoch(P.epoch);
        s_epoch = ep;
        if (atomicAdd(P.epoch + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
            *reinterpret_cast<volatile unsigned long long *>(P.epoch + 1) = 0ull;
            *reinterpret_cast<volatile unsigned long long *>(P.epoch) = ep + 1ull;
        }
        chain_seq = atomicAdd(P.chain + 2 * blockIdx.x, 1u);
        __threadfence();
    }
    __syncthreads();
    const unsigned long long launch_epoch = s_epoch;
    pdl_launch_dependents();
    if (P.chain_wait) {  // CTA b of the previous launch wrote the state blocks this CTA is about to load
        if (tid == 0)
            while (ld_acquire_gpu(P.chain + 2 * blockIdx.x + 1) != chain_seq) __nanosleep(40);
        __syncthreads();
        if ((tid & 31) == 0) fence_proxy_async_all();  // its generic-proxy stores, before this warp's bulk (async-proxy) loads
    }
    return launch_epoch;
}

// Per-thread tallies of a launch -> the CTA's own statistics slot (warp-shuffle reduction, then ONE plain read-modify-
// write per CTA: no atomics -- 740-1184 CTAs hammering one address cost 1-4 us per launch; qs_get_stats sums the slots).
struct StepTally {
    float reward = 0.0f;
    unsigned act = 0, done = 0, tr = 0, gp = 0, gc = 0, gr = 0, ob = 0;
    __device__ __forceinline__ void add(float r, uint32_t fl) {
        reward += r;
        act += 1; done += (fl & F_DONE) != 0; tr += (fl & F_TRUNC) != 0; gp += (fl & F_PASSED) != 0;
        gc += (fl & F_COLLISION) != 0; gr += (fl & F_GROUND) != 0; ob += (fl & F_OOB) != 0;
    }
};
template <int kWarps>
__device__ __forceinline__ void step_stats_flush(const StepParams &P, StepTally t) {
    __shared__ float s_red_f[kWarps];
    __shared__ unsigned s_red_u[kWarps][7];
    const int tid = threadIdx.x;
    const unsigned full_mask = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t.reward += __shfl_xor_sync(full_mask, t.reward, o);
    t.act = __reduce_add_sync(full_mask, t.act); t.done = __reduce_add_sync(full_mask, t.done);
    t.tr = __reduce_add_sync(full_mask, t.tr); t.gp = __reduce_add_sync(full_mask, t.gp);
    t.gc = __reduce_add_sync(full_mask, t.gc); t.gr = __reduce_add_sync(full_mask, t.gr);
    t.ob = __reduce_add_sync(full_mask, t.ob);
    if ((tid & 31) == 0) {
        const int w = tid >> 5;
        s_red_f[w] = t.reward;
        s_red_u[w][0] = t.act; s_red_u[w][1] = t.done; s_red_u[w][2] = t.tr; s_red_u[w][3] = t.gp;
        s_red_u[w][4] = t.gc; s_red_u[w][5] = t.gr; s_red_u[w][6] = t.ob;
    }
    __syncthreads();
    if (tid == 0) {
        float r = 0.0f;
        unsigned c[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            r += s_red_f[w];
#pragma unroll
            for (int k = 0; k < 7; ++k) c[k] += s_red_u[w][k];
        }
        if (c[0]) {
            Stats *st = P.stats + blockIdx.x;
            st->reward_sum += (double)r;
            st->env_steps += c[0]; st->dones += c[1]; st->truncated += c[2]; st->gates_passed += c[3];
            st->gate_collisions += c[4]; st->ground_collisions += c[5]; st->out_of_bounds += c[6];
        }
    }
}

// End of a step launch: called by every thread after the CTA's last store / draw (lane 0 of each warp has already
// waited for its bulk stores: .read for a classic launch, full completion for a chained one).
template <int kWarps>
__device__ __forceinline__ void step_launch_end(const StepParams &P, const StepTally &tally, unsigned chain_seq) {
    const int tid = threadIdx.x;
    __syncthreads();  // every warp of the CTA is past its last reset draw (and, chained, its last store)
    if (tid == 0 && P.advance_epoch && !P.chain) epoch_arrive(P);
    if (P.stats) step_stats_flush<kWarps>(P, tally);
    if (tid == 0 && P.chain) {  // chained: hand this CTA's tiles (and its stats slot) to CTA b of the next launch
        __threadfence();
        st_release_gpu(P.chain + 2 * blockIdx.x + 1, chain_seq + 1u);
    }
}

// Persistent CTAs (grid = SMs x resident CTAs) of four INDEPENDENT warps.  Each warp owns 32 envs of the CTA's
// 128-env tile and runs its own kStages-deep ring of shared-memory stages: lane 0 refills a stage with TMA bulk
// loads the moment the warp has copied it to registers, and waits on the stage's mbarrier for the bytes to land.
// There is no block barrier anywhere in the loop, so the warps of an SM drift apart: while one computes, others
// load or store.  Observations leave through the warp's slice of the staging tile as one TMA bulk store (32 rows
// of a row-major (N,D) array are contiguous).  kHints: the L2 residency policies (see l2_policy_*) are compiled in.
template <int V, int kStages, bool kHints>
__global__ void __launch_bounds__(StepCta<V>::THREADS, StepCta<V>::MIN_CTAS) step_kernel(const __grid_constant__ StepParams P) {
    using S = Stage<V>;
    constexpr int kWarps = StepCta<V>::WARPS, kStepThreads = StepCta<V>::THREADS;
    static_assert(kWarps * kStages * 8 <= kBarBytes, "mbarriers do not fit");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    // INDI: the shuffle tells the compiler that `warp` is warp-uniform, so everything derived from it (stage, barrier,
    // block and tile addresses) is computed once per warp on the uniform datapath and the TMA issues need no per-lane
    // address loop: 33.6 -> 31.0 us per step at N = 2^20.  E2E, whose arithmetic already fills the issue slots between
    // the address chains, measured 1 % slower with it (58.2 -> 58.8 us; profiles/r2/step_kernel_ablation.md).
    const int lane = tid & 31;
    const int warp = (V == kINDI) ? __shfl_sync(0xffffffffu, tid >> 5, 0) : (tid >> 5);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw) + warp * kStages;          // this warp's barriers
    unsigned char *stages = smem_raw + kBarBytes + warp * (kStages * S::BYTES);          // this warp's ring
    float *s_obs = reinterpret_cast<float *>(smem_raw + kBarBytes + kWarps * kStages * S::BYTES);
    const int slice = step_slice_floats(P.obs_len);
    float *s_track = s_obs + kWarps * slice;
    // work unit = a 32-env warp-tile (one state block); worker = a warp; the launch covers the 128-env tiles
    // [tile_begin, tile_end) = warp-tiles [4*tile_begin, 4*tile_end), dealt round-robin to the grid's warps
    const long long n_tiles = P.tile_end * (kBlock / 32);
    const long long tile0 = P.tile_begin * (kBlock / 32) + (long long)blockIdx.x * kWarps + warp;
    const long long stride = (long long)gridDim.x * kWarps;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_gates * kTrackRow; i += kStepThreads) s_track[i] = P.track[i];
    const uint64_t pol_keep = kHints ? l2_policy_evict_last() : 0ull, pol_stream = kHints ? l2_policy_evict_first() : 0ull;
    __syncthreads();  // barrier init and track table visible to everyone
    // Everything above touched only launch constants.  From here on we read and write simulator state that the
    // previous step's grid may still be producing (programmatic dependent launch).
    unsigned chain_seq;
    const unsigned long long launch_epoch = step_launch_begin(P, chain_seq);  // the RNG epoch of this launch
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            const long long t = tile0 + (long long)s * stride;
            if (t < n_tiles) issue_warp_tile<V, kHints>(P, stages + s * S::BYTES, &full[s], t * 32, pol_keep, pol_stream);
        }
    }

    // ------------------------------------------------------------------------------------------ consumers
    StepTally tally;  // stats are reduced once per CTA lifetime, not per tile
    const bool write_obs_tile = P.mode != kModePause;
    const bool fused_reset = P.mode == kModeNormal && P.reset_source == kResetDevice;
    float *const warp_obs = s_obs + warp * slice;
    float *const my_obs = warp_obs + lane * P.obs_len;
    bool obs_in_flight = false;  // warp-uniform: a bulk store of this warp's observation slice may still be reading it
    int it = 0;
    for (long long tile = tile0; tile < n_tiles; tile += stride, ++it) {
        const int stage = it % kStages;
        unsigned char *st = stages + stage * S::BYTES;
        const long long base = tile * 32;  // first env of this warp-tile
        const long long env = base + lane;
        const bool active = env < P.n;
        mbar_wait(&full[stage], (uint32_t)(it / kStages) & 1u);

        // ---- stage -> registers (conflict-free: consecutive lanes read consecutive 16-byte words)
        EnvState<V> e, n;
        {
            const float4 a = reinterpret_cast<const float4 *>(st + S::P0)[lane];
            const float4 b = reinterpret_cast<const float4 *>(st + S::P1)[lane];
            const float4 c = reinterpret_cast<const float4 *>(st + S::P2)[lane];
            e.x = a.x; e.y = a.y; e.z = a.z; e.vx = a.w; e.vy = b.x; e.vz = b.y; e.phi = b.z; e.th = b.w;
            e.psi = c.x; e.p = c.y; e.q = c.z; e.r = c.w;
            if (V == kE2E) {
                const float4 d = reinterpret_cast<const float4 *>(st + S::P3)[lane];
                const float4 da = reinterpret_cast<const float4 *>(st + S::DA)[lane];
                const float2 db = reinterpret_cast<const float2 *>(st + S::DB)[lane];
                e.w[0] = d.x; e.w[1] = d.y; e.w[2] = d.z; e.w[V == kE2E ? 3 : 0] = d.w;
                e.dist[0] = da.x; e.dist[1] = da.y; e.dist[2] = da.z; e.dist[5] = da.w; e.dist[3] = db.x; e.dist[4] = db.y;
            } else {
                e.w[0] = reinterpret_cast<const float *>(st + S::P3)[lane];
            }
        }
        const float4 u = reinterpret_cast<const float4 *>(st + S::ACT)[lane];
        const uint32_t meta = reinterpret_cast<const uint32_t *>(st + S::META)[lane];
        uint32_t tg = meta >> 24, sc = meta & kStepMask;
        __syncwarp();  // the warp's inputs are in registers: refill the stage with the tile kStages ahead
        if (lane == 0) {
            const long long nt = tile + (long long)kStages * stride;
            if (nt < n_tiles) issue_warp_tile<V, kHints>(P, st, &full[stage], nt * 32, pol_keep, pol_stream);
        }

#ifdef QS_EXP_NOCOMPUTE  // experiment only: the memory pipeline alone (same loads and stores, no arithmetic)
        n = e; n.x = e.x + u.x; n.y = e.y + u.y + u.z + u.w;
        float reward = u.y; const bool dn = false; const uint32_t fl = 0;
#else
        euler_step<V>(P, e, u, n);
        float reward; bool dn; uint32_t fl;
        reward_and_flags<V>(P, s_track, e, n, tg, sc, reward, dn, fl);
#endif

        // ---- branch logic (`:568-585`)
        bool write_world = active, write_dist = false;
        if (fused_reset) {  // warp-uniform branch: the whole warp takes part in the draw
#ifdef QS_EXP_NORESET  // experiment only: what the fused reset path costs
            const bool need = false;
#else
            const bool need = dn && active;
#endif
            // scratch = the head of this warp's observation slice: its previous bulk store must have been read
            if (__any_sync(0xffffffffu, need)) {
                if (lane == 0 && obs_in_flight) bulk_store_wait_read();
                obs_in_flight = false;
                __syncwarp();
                draw_reset_warp<V>(P, base, need, reinterpret_cast<uint4 *>(warp_obs), n, launch_epoch);
            }
            if (need) {
                tg = 0; sc = 0;
                write_dist = (V == kE2E);
            }
        } else if (P.mode == kModePauseIfCollision) {
            if (dn) { n = e; write_world = false; }
        } else if (P.mode == kModePause) {
            write_world = false;
        }
        unsigned char *const gblk = P.s.base + tile * (long long)Blk<V>::BYTES;  // this warp's block
        const uint8_t dn8 = (P.mode == kModePause) ? (uint8_t)0 : (uint8_t)(dn ? 1 : 0);
        if (!kHints) {
            if (active) {
                reinterpret_cast<uint32_t *>(gblk + S::META)[lane] = (tg << 24) | sc;
                P.rew[env] = reward;
                P.done[env] = dn8;
                if (P.flags) P.flags[env] = (uint8_t)fl;
            }
            if (write_world) {
                reinterpret_cast<float4 *>(gblk + S::P0)[lane] = make_float4(n.x, n.y, n.z, n.vx);
                reinterpret_cast<float4 *>(gblk + S::P1)[lane] = make_float4(n.vy, n.vz, n.phi, n.th);
                reinterpret_cast<float4 *>(gblk + S::P2)[lane] = make_float4(n.psi, n.p, n.q, n.r);
                if (V == kE2E) reinterpret_cast<float4 *>(gblk + S::P3)[lane] = make_float4(n.w[0], n.w[1], n.w[2], n.w[V == kE2E ? 3 : 0]);
                else reinterpret_cast<float *>(gblk + S::P3)[lane] = n.w[0];
            }
            if (V == kE2E && write_dist) {
                reinterpret_cast<float4 *>(gblk + S::DA)[lane] = make_float4(n.dist[0], n.dist[1], n.dist[2], n.dist[5]);
                reinterpret_cast<float2 *>(gblk + S::DB)[lane] = make_float2(n.dist[3], n.dist[4]);
            }
        } else {  // the same stores with L2 priorities: state lines stay, result streams leave first
            const uint64_t pol_state = tile < P.keep_blocks ? pol_keep : pol_stream;
            if (active) {
                st_hint(reinterpret_cast<uint32_t *>(gblk + S::META) + lane, (tg << 24) | sc, pol_state);
                st_hint(P.rew + env, reward, pol_stream);
                st_hint(P.done + env, dn8, pol_stream);
                if (P.flags) st_hint(P.flags + env, (uint8_t)fl, pol_stream);
            }
            if (write_world) {
                st_hint(reinterpret_cast<float4 *>(gblk + S::P0) + lane, make_float4(n.x, n.y, n.z, n.vx), pol_state);
                st_hint(reinterpret_cast<float4 *>(gblk + S::P1) + lane, make_float4(n.vy, n.vz, n.phi, n.th), pol_state);
                st_hint(reinterpret_cast<float4 *>(gblk + S::P2) + lane, make_float4(n.psi, n.p, n.q, n.r), pol_state);
                if (V == kE2E) st_hint(reinterpret_cast<float4 *>(gblk + S::P3) + lane, make_float4(n.w[0], n.w[1], n.w[2], n.w[V == kE2E ? 3 : 0]), pol_state);
                else st_hint(reinterpret_cast<float *>(gblk + S::P3) + lane, n.w[0], pol_state);
            }
            if (V == kE2E && write_dist) {
                st_hint(reinterpret_cast<float4 *>(gblk + S::DA) + lane, make_float4(n.dist[0], n.dist[1], n.dist[2], n.dist[5]), pol_state);
                st_hint(reinterpret_cast<float2 *>(gblk + S::DB) + lane, make_float2(n.dist[3], n.dist[4]), pol_state);
            }
        }

        if (write_obs_tile) {
            if (lane == 0 && obs_in_flight) bulk_store_wait_read();  // the previous tile's rows have left this warp's slice
            __syncwarp();
#ifdef QS_EXP_NOCOMPUTE
            if (V == kE2E) {
                for (int k = 0; k < P.obs_len; k += 4) *reinterpret_cast<float4 *>(my_obs + k) = make_float4(n.x, n.y, n.z, n.vx);
            } else {
                for (int k = 0; k < P.obs_len; ++k) my_obs[k] = n.x;
            }
#else
            write_obs<V>(P, s_track, n, tg, my_obs);
#endif
            if (P.obs_packed) {  // warp-uniform: the rows leave as ONE packed 2 KB block (see pack_obs_row)
                uint32_t w[16];
                pack_obs_row(my_obs, P.obs_len, active, w);
                __syncwarp();  // every lane has read its float32 row: the block is built over them
                uint4 *blk = reinterpret_cast<uint4 *>(warp_obs);
#pragma unroll
                for (int c = 0; c < kPackK / 8; ++c)
                    if (c < pack_chunks(P.obs_len)) blk[c * 32 + lane] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const long long b32 = tile;  // 32-env block index within this launch's env range
                    const uint32_t pkb = (uint32_t)pack_block_bytes(P.obs_len);
                    unsigned char *pk = reinterpret_cast<unsigned char *>(P.obs) + b32 * (long long)pkb;
                    if (kHints) bulk_store_hint(pk, warp_obs, pkb, pol_stream);
                    else bulk_store(pk, warp_obs, pkb);
#pragma unroll 1
                    for (int p = 0; p < P.n_peers; ++p)
                        bulk_store(reinterpret_cast<unsigned char *>(P.peer_obs[p]) + ((P.peer_row_offset >> 5) + b32) * (long long)pkb, warp_obs, pkb);
                }
                obs_in_flight = true;
                if (P.stats && active) tally.add(reward, fl);
                continue;
            }
            const long long rem = P.n - base;
            const int rows = rem < 32 ? (rem < 0 ? 0 : (int)rem) : 32;
            float *dst = P.obs + base * P.obs_len;
            const uint32_t bytes = (uint32_t)rows * (uint32_t)P.obs_len * 4u;
            // (peer destinations: base pointers are 16-byte aligned and qs_set_obs_peers rejects a row offset that
            // would misalign them, so the local test covers every destination of the tile)
            const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
            if (bulk) {  // the warp's rows are contiguous in the (N,D) output: one TMA bulk store
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && bytes) {
                    if (kHints) bulk_store_hint(dst, warp_obs, bytes, pol_stream);
                    else bulk_store(dst, warp_obs, bytes);
#pragma unroll 1
                    for (int p = 0; p < P.n_peers; ++p)  // NVLink: the tile leaves for every peer while the next one is computed
                        bulk_store(P.peer_obs[p] + (P.peer_row_offset + base) * P.obs_len, warp_obs, bytes);
                }
                obs_in_flight = true;
            } else {
                __syncwarp();
#pragma unroll 1
                for (int i = lane; i < rows * P.obs_len; i += 32) dst[i] = warp_obs[i];
#pragma unroll 1
                for (int p = 0; p < P.n_peers; ++p) {
                    float *pd = P.peer_obs[p] + (P.peer_row_offset + base) * P.obs_len;
#pragma unroll 1
                    for (int i = lane; i < rows * P.obs_len; i += 32) pd[i] = warp_obs[i];
                }
                __syncwarp();
            }
        }
        if (P.stats && active) tally.add(reward, fl);
    }
    if (lane == 0 && write_obs_tile) {
        if (P.chain) bulk_store_wait_all();  // chained: the rows have LANDED before the next launch's CTA may rewrite them
        else bulk_store_wait_read();         // shared memory must outlive the last bulk read
    }
    step_launch_end<kWarps>(P, tally, chain_seq);
}

// ------------------------------------------------------------------------------------------------ E2E, two envs per thread
// EXPERIMENT, opt-in (QS_STEP_X2=1), measured SLOWER than step_kernel<e2e> on B200 (75 vs 57 us per step at N = 2^20).
// The E2E step is bound by instruction issue, not by HBM (with chained launches its memory pipeline alone runs at 0.96 of
// the roofline, the full kernel at 0.80): 615 of its 1430 warp-instructions per 32 envs are the residual nets, and a third
// of those only fetch weights (LDCU / LDC).  Here a thread steps TWO envs (a lane of warp-tile a and the same lane of
// warp-tile b): every weight fetch feeds both (residual_mlp2: 173 LDCU.128 for 672 FFMA2) and the per-tile bookkeeping
// is paid once per 64 envs -- 15 % fewer instructions per env.  But the loop body doubles (the nets must stay fully
// unrolled, see residual_mlp), only 12-16 warps fit per SM (138 registers), and the per-warp issue rate does not rise
// with the second independent chain: the instruction fetch, not the dependency latency, is what limits a warp here.
// Kept because it is bit-identical per env to step_kernel (tests/test_gpu_chain.py) and documents the dead end.
// Shared memory per warp: a ring of FOUR stage buffers, used in pairs.  Iteration i waits for pair i&1, copies it to
// registers, and -- the buffers now being free -- stages its two observation tiles in them for the TMA bulk stores;
// half an iteration later (behind the dynamics of iteration i+1) those stores have long read the buffers and lane 0
// refills them with the tiles of iteration i+2.  No separate observation staging: 13.5 KB per warp.
// Scope: float32 observations that fit a stage buffer (obs_len <= 27: gates_ahead <= 1), no peer stores, no L2 hints;
// everything else runs step_kernel<kE2E>.
constexpr int kX2Warps = 4, kX2Bufs = 4;
__host__ __device__ constexpr size_t step_x2_smem_bytes(int n_gates) {
    return kBarBytes + (size_t)kX2Warps * kX2Bufs * Stage<kE2E>::BYTES + (size_t)n_gates * kTrackRow * 4;
}
__host__ __device__ constexpr bool step_x2_supports(int obs_len) { return 32 * obs_len * 4 <= (int)Stage<kE2E>::BYTES; }

#ifndef QS_X2_MIN_CTAS
#define QS_X2_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(kX2Warps * 32, QS_X2_MIN_CTAS) step_kernel_x2(const __grid_constant__ StepParams P) {
    constexpr int V = kE2E;
    using S = Stage<V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw) + warp * kX2Bufs;                  // this warp's barriers
    unsigned char *bufs = smem_raw + kBarBytes + warp * (kX2Bufs * S::BYTES);                  // this warp's ring
    float *s_track = reinterpret_cast<float *>(smem_raw + kBarBytes + kX2Warps * kX2Bufs * S::BYTES);
    const long long n_tiles = P.tile_end * (kBlock / 32);                                      // in 32-env warp-tiles
    const long long tile0 = P.tile_begin * (kBlock / 32) + (long long)blockIdx.x * kX2Warps + warp;
    const long long stride = (long long)gridDim.x * kX2Warps;

    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kX2Bufs; ++b) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_gates * kTrackRow; i += kX2Warps * 32) s_track[i] = P.track[i];
    __syncthreads();
    unsigned chain_seq;
    const unsigned long long launch_epoch = step_launch_begin(P, chain_seq);
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kX2Bufs; ++b) {
            const long long t = tile0 + (long long)b * stride;
            if (t < n_tiles) issue_warp_tile<V, false>(P, bufs + b * S::BYTES, &full[b], t * 32, 0ull, 0ull);
        }
    }

    StepTally tally;
    const bool write_obs_tile = P.mode != kModePause;
    const bool fused_reset = P.mode == kModeNormal && P.reset_source == kResetDevice;
    const uint32_t row_bytes = (uint32_t)P.obs_len * 4u;
    int it = 0;
    for (long long tile_a = tile0; tile_a < n_tiles; tile_a += 2 * stride, ++it) {
        const int pair = it & 1;
        const uint32_t phase = (uint32_t)(it >> 1) & 1u;
        unsigned char *const buf[2] = {bufs + (2 * pair) * S::BYTES, bufs + (2 * pair + 1) * S::BYTES};
        const long long tile[2] = {tile_a, tile_a + stride};
        const bool have_b = tile[1] < n_tiles;  // warp-uniform: the last iteration of a warp may hold one tile only
        mbar_wait(&full[2 * pair], phase);
        if (have_b) mbar_wait(&full[2 * pair + 1], phase);

        // ---- stage buffers -> registers
        EnvState<V> e[2], n[2];
        float4 u[2];
        uint32_t tg[2], sc[2];
        bool active[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned char *st = buf[h];
            const float4 a = reinterpret_cast<const float4 *>(st + S::P0)[lane];
            const float4 b = reinterpret_cast<const float4 *>(st + S::P1)[lane];
            const float4 c = reinterpret_cast<const float4 *>(st + S::P2)[lane];
            const float4 d = reinterpret_cast<const float4 *>(st + S::P3)[lane];
            const float4 da = reinterpret_cast<const float4 *>(st + S::DA)[lane];
            const float2 db = reinterpret_cast<const float2 *>(st + S::DB)[lane];
            e[h].x = a.x; e[h].y = a.y; e[h].z = a.z; e[h].vx = a.w; e[h].vy = b.x; e[h].vz = b.y; e[h].phi = b.z; e[h].th = b.w;
            e[h].psi = c.x; e[h].p = c.y; e[h].q = c.z; e[h].r = c.w;
            e[h].w[0] = d.x; e[h].w[1] = d.y; e[h].w[2] = d.z; e[h].w[3] = d.w;
            e[h].dist[0] = da.x; e[h].dist[1] = da.y; e[h].dist[2] = da.z; e[h].dist[5] = da.w; e[h].dist[3] = db.x; e[h].dist[4] = db.y;
            u[h] = reinterpret_cast<const float4 *>(st + S::ACT)[lane];
            const uint32_t meta = reinterpret_cast<const uint32_t *>(st + S::META)[lane];
            tg[h] = meta >> 24; sc[h] = meta & kStepMask;
            active[h] = (h == 0 || have_b) && (tile[h] * 32 + lane < P.n);
        }
        if (!have_b) { tg[1] = 0; sc[1] = 0; }  // whatever bytes an unused buffer holds: keep the table index in range
        __syncwarp();  // the warp's inputs are in registers: both buffers are free

        // ---- the OTHER pair of buffers staged the previous iteration's observation tiles: as soon as those bulk stores
        // have read them (they were issued at the end of that iteration) refill them with the tiles of the next one --
        // a whole iteration of lead for the loads
        if (it > 0 && lane == 0) {
            bulk_store_wait_read();
            const long long na = tile_a + 2 * stride;
            unsigned char *nb = bufs + (2 * (pair ^ 1)) * S::BYTES;
            if (na < n_tiles) issue_warp_tile<V, false>(P, nb, &full[2 * (pair ^ 1)], na * 32, 0ull, 0ull);
            if (na + stride < n_tiles) issue_warp_tile<V, false>(P, nb + S::BYTES, &full[2 * (pair ^ 1) + 1], (na + stride) * 32, 0ull, 0ull);
        }

        // ---- dynamics: the two envs share every weight fetch
        EulerMid mid[2];
        float xin[2][10], thr[2], mom[2][3];
        euler_pre<V>(e[0], mid[0], xin[0]);
        euler_pre<V>(e[1], mid[1], xin[1]);
        residual_mlp2(P, xin[0], xin[1], thr[0], mom[0], thr[1], mom[1]);
        euler_post<V>(P, e[0], u[0], mid[0], thr[0], mom[0], n[0]);
        euler_post<V>(P, e[1], u[1], mid[1], thr[1], mom[1], n[1]);

        // The rest runs in phases over both envs (not env by env), so that each phase is one large basic block in which
        // the two independent instruction streams interleave.
        float reward[2]; bool dn[2]; uint32_t fl[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) reward_and_flags<V>(P, s_track, e[h], n[h], tg[h], sc[h], reward[h], dn[h], fl[h]);
        // ---- branch logic (`:568-585`)
        bool write_world[2] = {active[0], active[1]}, write_dist[2] = {false, false};
        if (fused_reset) {  // warp-uniform branch: the whole warp takes part in the draws
            const bool need0 = dn[0] && active[0], need1 = dn[1] && active[1];
            if (__any_sync(0xffffffffu, need0 | need1)) {  // scratch = a tile's own (free) stage buffer
                draw_reset_warp<V>(P, tile[0] * 32, need0, reinterpret_cast<uint4 *>(buf[0]), n[0], launch_epoch);
                draw_reset_warp<V>(P, tile[1] * 32, need1, reinterpret_cast<uint4 *>(buf[1]), n[1], launch_epoch);
            }
            if (need0) { tg[0] = 0; sc[0] = 0; write_dist[0] = true; }
            if (need1) { tg[1] = 0; sc[1] = 0; write_dist[1] = true; }
        } else if (P.mode == kModePauseIfCollision) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (dn[h]) { n[h] = e[h]; write_world[h] = false; }
        } else if (P.mode == kModePause) {
            write_world[0] = write_world[1] = false;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long env = tile[h] * 32 + lane;
            unsigned char *const gblk = P.s.base + tile[h] * (long long)Blk<V>::BYTES;  // this warp-tile's block
            const uint8_t dn8 = (P.mode == kModePause) ? (uint8_t)0 : (uint8_t)(dn[h] ? 1 : 0);
            if (active[h]) {
                reinterpret_cast<uint32_t *>(gblk + S::META)[lane] = (tg[h] << 24) | sc[h];
                P.rew[env] = reward[h];
                P.done[env] = dn8;
                if (P.flags) P.flags[env] = (uint8_t)fl[h];
            }
            if (write_world[h]) {
                reinterpret_cast<float4 *>(gblk + S::P0)[lane] = make_float4(n[h].x, n[h].y, n[h].z, n[h].vx);
                reinterpret_cast<float4 *>(gblk + S::P1)[lane] = make_float4(n[h].vy, n[h].vz, n[h].phi, n[h].th);
                reinterpret_cast<float4 *>(gblk + S::P2)[lane] = make_float4(n[h].psi, n[h].p, n[h].q, n[h].r);
                reinterpret_cast<float4 *>(gblk + S::P3)[lane] = make_float4(n[h].w[0], n[h].w[1], n[h].w[2], n[h].w[3]);
            }
            if (write_dist[h]) {
                reinterpret_cast<float4 *>(gblk + S::DA)[lane] = make_float4(n[h].dist[0], n[h].dist[1], n[h].dist[2], n[h].dist[5]);
                reinterpret_cast<float2 *>(gblk + S::DB)[lane] = make_float2(n[h].dist[3], n[h].dist[4]);
            }
            if (P.stats && active[h]) tally.add(reward[h], fl[h]);
        }
        if (write_obs_tile) {
            __syncwarp();  // (a reset draw may just have used the buffers as scratch)
            write_obs<V>(P, s_track, n[0], tg[0], reinterpret_cast<float *>(buf[0]) + lane * P.obs_len);
            if (have_b) write_obs<V>(P, s_track, n[1], tg[1], reinterpret_cast<float *>(buf[1]) + lane * P.obs_len);
            fence_proxy_async();
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !have_b) break;
                const long long base = tile[h] * 32;
                const float *tile_obs = reinterpret_cast<const float *>(buf[h]);
                const long long rem = P.n - base;
                const int rows = rem < 32 ? (rem < 0 ? 0 : (int)rem) : 32;
                float *dst = P.obs + base * P.obs_len;
                const uint32_t bytes = (uint32_t)rows * row_bytes;
                const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
                if (bulk) {  // the warp's rows are contiguous in the (N,D) output: one TMA bulk store
                    if (lane == 0 && bytes) bulk_store(dst, tile_obs, bytes);
                } else {
#pragma unroll 1
                    for (int i = lane; i < rows * P.obs_len; i += 32) dst[i] = tile_obs[i];
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && write_obs_tile) {
        if (P.chain) bulk_store_wait_all();
        else bulk_store_wait_read();
    }
    step_launch_end<kX2Warps>(P, tally, chain_seq);
}

// ------------------------------------------------------------------------------------------------ observe / reset kernels
// update_states() on its own (after qs_set_state / reset), and reset() with the device RNG when reset_all != 0.
template <int V>
__global__ void __launch_bounds__(kBlock) observe_kernel(const __grid_constant__ StepParams P, int reset_all) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_obs = reinterpret_cast<float *>(smem_raw);
    float *s_track = s_obs + kBlock * P.obs_len;
    load_track(P, s_track);
    const long long base = (long long)blockIdx.x * kBlock;
    const long long env = base + threadIdx.x;
    if (env < P.n) {
        EnvState<V> e;
        uint32_t tg;
        if (reset_all) {
            draw_reset<V>(P, env, e, load_epoch(P.epoch));
            tg = 0;
            field<V, Blk<V>::META, uint32_t>(P.s, env) = 0;
            store_world<V>(P.s, env, e);
            store_dist<V>(P.s, env, e);
        } else {
            load_state<V>(P.s, env, e);
            tg = field<V, Blk<V>::META, uint32_t>(P.s, env) >> 24;
        }
        write_obs<V>(P, s_track, e, tg, s_obs + threadIdx.x * P.obs_len);
    }
    if (P.obs_packed) {  // packed blocks: four coalesced 512-byte stores per warp, straight from registers
        uint32_t w[16];
        pack_obs_row(s_obs + threadIdx.x * P.obs_len, P.obs_len, env < P.n, w);
        uint4 *blk = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(P.obs) + (env >> 5) * (long long)pack_block_bytes(P.obs_len));
#pragma unroll
        for (int c = 0; c < kPackK / 8; ++c)
            if (c < pack_chunks(P.obs_len)) blk[c * 32 + (threadIdx.x & 31)] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        if (reset_all && threadIdx.x == 0) epoch_arrive(P);
        return;
    }
    const long long rem = P.n - base;
    store_obs_tile(P.obs + base * P.obs_len, s_obs, rem < kBlock ? (int)rem : kBlock, P.obs_len);  // block barrier inside
    if (reset_all && threadIdx.x == 0) epoch_arrive(P);
}

// Masked stores of reset_ with host-drawn values (`:476-489`) + the observation rows of those envs.
template <int V>
__global__ void __launch_bounds__(kBlock) apply_reset_kernel(const __grid_constant__ StepParams P, long long count,
                                                             const int *idx, const float *ws, const float *dist) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_obs = reinterpret_cast<float *>(smem_raw);
    float *s_track = s_obs + kBlock * P.obs_len;
    load_track(P, s_track);
    const long long k = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (k >= count) return;
    const long long env = idx[k];
    constexpr int NS = V == kE2E ? 16 : 13;
    const float *r = ws + k * NS;
    EnvState<V> e;
    e.x = r[0]; e.y = r[1]; e.z = r[2]; e.vx = r[3]; e.vy = r[4]; e.vz = r[5];
    e.phi = r[6]; e.th = r[7]; e.psi = r[8]; e.p = r[9]; e.q = r[10]; e.r = r[11];
#pragma unroll
    for (int j = 0; j < NS - 12; ++j) e.w[j] = r[12 + j];
    if (V == kE2E) {
        if (dist) {
#pragma unroll
            for (int j = 0; j < 6; ++j) e.dist[j] = dist[k * 6 + j];
            store_dist<V>(P.s, env, e);
        } else {
            const float4 da = field<V, Blk<V>::DA, float4>(P.s, env);
            const float2 db = field<V, Blk<V>::DB, float2>(P.s, env);
            e.dist[0] = da.x; e.dist[1] = da.y; e.dist[2] = da.z; e.dist[5] = da.w; e.dist[3] = db.x; e.dist[4] = db.y;
        }
    }
    store_world<V>(P.s, env, e);
    field<V, Blk<V>::META, uint32_t>(P.s, env) = 0;
    float *row = s_obs + threadIdx.x * P.obs_len;
    write_obs<V>(P, s_track, e, 0u, row);
    float *dst = P.obs + env * P.obs_len;
    for (int j = 0; j < P.obs_len; ++j) dst[j] = row[j];
}

// What step_wait's per-env `infos` loop computes (`3D quad race.ipynb:589-594`), from the flag bytes of a step: the highest
// done env index (its observation row becomes the ONE aliased dict's "terminal_observation"), the number of done envs
// and whether any env was truncated.  out: [0] last done index + 1 (0 = none), [1] done count, [2] any-truncated.
__global__ void step_info_kernel(const uint8_t *flags, long long n, unsigned long long *out) {
    unsigned long long last1 = 0, cnt = 0, tr = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t f = flags[i];
        if (f & F_DONE) { last1 = (unsigned long long)i + 1ull; ++cnt; }
        tr |= (f & F_TRUNC) ? 1ull : 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, last1, o);
        last1 = a > last1 ? a : last1;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        tr |= __shfl_xor_sync(0xffffffffu, tr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (last1) atomicMax(out, last1);
        if (cnt) atomicAdd(out + 1, cnt);
        if (tr) atomicOr(out + 2, 1ull);
    }
}

// ------------------------------------------------------------------------------------------------ AoS <-> planes
template <int V>
__global__ void import_kernel(Planes s, long long first, long long count, const float *ws, const float *dist,
                              const long long *tg, const long long *sc, int n_gates) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long env = first + k;
    using B = Blk<V>;
    constexpr int NS = V == kE2E ? 16 : 13;
    if (ws) {
        const float *r = ws + k * NS;
        field<V, B::P0, float4>(s, env) = make_float4(r[0], r[1], r[2], r[3]);
        field<V, B::P1, float4>(s, env) = make_float4(r[4], r[5], r[6], r[7]);
        field<V, B::P2, float4>(s, env) = make_float4(r[8], r[9], r[10], r[11]);
        if (V == kE2E) field<V, B::P3, float4>(s, env) = make_float4(r[12], r[13], r[14], r[V == kE2E ? 15 : 12]);
        else field<V, B::P3, float>(s, env) = r[12];
    }
    if (V == kE2E && dist) {
        const float *d = dist + k * 6;
        field<V, B::DA, float4>(s, env) = make_float4(d[0], d[1], d[2], d[5]);
        field<V, B::DB, float2>(s, env) = make_float2(d[3], d[4]);
    }
    if (tg || sc) {
        const uint32_t m = field<V, B::META, uint32_t>(s, env);
        uint32_t g = m >> 24, c = m & kStepMask;
        if (tg) { long long t = tg[k] % n_gates; if (t < 0) t += n_gates; g = (uint32_t)t; }
        if (sc) { long long v = sc[k]; c = v < 0 ? 0u : (v > (long long)kStepMask ? kStepMask : (uint32_t)v); }
        field<V, B::META, uint32_t>(s, env) = (g << 24) | c;
    }
}

template <int V>
__global__ void export_kernel(Planes s, long long first, long long count, float *ws, float *dist, long long *tg,
                              long long *sc) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long env = first + k;
    using B = Blk<V>;
    constexpr int NS = V == kE2E ? 16 : 13;
    if (ws) {
        float *r = ws + k * NS;
        const float4 a = field<V, B::P0, float4>(s, env), b = field<V, B::P1, float4>(s, env), c = field<V, B::P2, float4>(s, env);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
        r[8] = c.x; r[9] = c.y; r[10] = c.z; r[11] = c.w;
        if (V == kE2E) { const float4 d = field<V, B::P3, float4>(s, env); r[12] = d.x; r[13] = d.y; r[14] = d.z; r[V == kE2E ? 15 : 12] = d.w; }
        else r[12] = field<V, B::P3, float>(s, env);
    }
    if (V == kE2E && dist) {
        const float4 da = field<V, B::DA, float4>(s, env);
        const float2 db = field<V, B::DB, float2>(s, env);
        float *d = dist + k * 6;
        d[0] = da.x; d[1] = da.y; d[2] = da.z; d[3] = db.x; d[4] = db.y; d[5] = da.w;
    }
    if (tg || sc) {
        const uint32_t m = field<V, B::META, uint32_t>(s, env);
        if (tg) tg[k] = m >> 24;
        if (sc) sc[k] = m & kStepMask;
    }
}

}  // namespace qs
