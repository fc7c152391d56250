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

struct Planes {
    float4 *p0, *p1, *p2;
    float4 *p3;      // E2E motor speeds
    float *p3s;      // INDI T_norm
    float4 *da;      // E2E (Mx,My,Mz,Fz)
    float2 *db;      // E2E (Fx,Fy)
    uint32_t *meta;  // target_gate<<24 | step_count
    uint32_t *episode;
};

struct ResetDist {   // reset_ draw ranges (`:455-489`)
    float start[3];
    float dist_lo[6], dist_span[6];  // already multiplied by disturbance_scale
};

struct StepParams {
    Planes s;
    const float4 *actions;
    float *obs;
    float *rew;
    uint8_t *done;
    uint8_t *flags;
    const float *track;  // (n_gates, kTrackRow): gx gy gz yaw cos sin 0 0 | rel_x rel_y rel_z rel_yaw
    Stats *stats;
    long long n;
    long long env_offset;
    unsigned long long seed;
    int n_gates, gates_ahead, obs_len;
    int mode, reset_source;
    uint32_t max_steps;
    float dt;
    float obs_scale[4], obs_off[4];  // disturbance observation: d*scale+off  (`:414-448`)
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
__device__ __forceinline__ float wrap_yaw(float a) {
    const float b = 6.283185307179586f, pi = 3.141592653589793f;
    float m;
    if (fabsf(a) < 1.0e5f) {
        float n = floorf(a * 0.15915494309189535f);
        m = fmaf(-n, b, a);
        if (m < 0.0f) m = add_rn(m, b);
        if (m >= b) m = sub_rn(m, b);
    } else {  // blown-up state: exact but slow path
        m = fmodf(a, b);
        if (m < 0.0f) m = add_rn(m, b);
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
    const float4 a = s.p0[i], b = s.p1[i], c = s.p2[i];
    e.x = a.x; e.y = a.y; e.z = a.z; e.vx = a.w;
    e.vy = b.x; e.vz = b.y; e.phi = b.z; e.th = b.w;
    e.psi = c.x; e.p = c.y; e.q = c.z; e.r = c.w;
    if (V == kE2E) {
        const float4 d = s.p3[i];
        e.w[0] = d.x; e.w[1] = d.y; e.w[2] = d.z; e.w[V == kE2E ? 3 : 0] = d.w;
        const float4 da = s.da[i];
        const float2 db = s.db[i];
        e.dist[0] = da.x; e.dist[1] = da.y; e.dist[2] = da.z; e.dist[5] = da.w;
        e.dist[3] = db.x; e.dist[4] = db.y;
    } else {
        e.w[0] = s.p3s[i];
    }
}

template <int V>
__device__ __forceinline__ void store_world(const Planes &s, long long i, const EnvState<V> &e) {
    s.p0[i] = make_float4(e.x, e.y, e.z, e.vx);
    s.p1[i] = make_float4(e.vy, e.vz, e.phi, e.th);
    s.p2[i] = make_float4(e.psi, e.p, e.q, e.r);
    if (V == kE2E) s.p3[i] = make_float4(e.w[0], e.w[1], e.w[2], e.w[V == kE2E ? 3 : 0]);
    else s.p3s[i] = e.w[0];
}

template <int V>
__device__ __forceinline__ void store_dist(const Planes &s, long long i, const EnvState<V> &e) {
    if (V == kE2E) {
        s.da[i] = make_float4(e.dist[0], e.dist[1], e.dist[2], e.dist[5]);
        s.db[i] = make_float2(e.dist[3], e.dist[4]);
    }
}

// reset_ (`:452-489`) with the device RNG: same fields, same ranges, counter-based instead of MT19937.
// Draw c of env g in episode ep is Philox4x32-10(counter=(g_lo,g_hi,ep,c), key=seed): 6 draws (E2E) / 4 (INDI).
template <int V> struct ResetDraws { enum : int { N = (V == kE2E ? 6 : 4) }; };

__device__ __forceinline__ uint4 reset_draw(const StepParams &P, unsigned long long g, uint32_t episode, uint32_t c) {
    return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), episode, c),
                         make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
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
__device__ __forceinline__ void draw_reset(const StepParams &P, long long env, uint32_t episode, EnvState<V> &e) {
    const unsigned long long g = (unsigned long long)(env + P.env_offset);
    uint4 r[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) r[c] = c < ResetDraws<V>::N ? reset_draw(P, g, episode, c) : make_uint4(0, 0, 0, 0);
    fill_reset<V>(P, r, e);
}

// Warp-cooperative form for the common case of one or two terminating lanes per warp: instead of one lane
// running 6 Philox evaluations while 31 wait, lanes 0..5 each run ONE for that env and hand the words over by
// shuffle.  Must be called by the whole warp; `need` marks the lanes that reset.  Same values as draw_reset.
template <int V>
__device__ __forceinline__ void draw_reset_warp(const StepParams &P, long long env, uint32_t episode, bool need,
                                                EnvState<V> &e) {
    const unsigned full = 0xffffffffu;
    unsigned m = __ballot_sync(full, need);
    if (m == 0) return;
    if (__popc(m) > 4) {  // mass termination (e.g. synchronous time-outs): per-lane is cheaper
        if (need) draw_reset<V>(P, env, episode, e);
        return;
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long g_own = (unsigned long long)(env + P.env_offset);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t g_lo = __shfl_sync(full, (uint32_t)g_own, src), g_hi = __shfl_sync(full, (uint32_t)(g_own >> 32), src);
        const uint32_t ep = __shfl_sync(full, episode, src);
        const uint4 mine = reset_draw(P, ((unsigned long long)g_hi << 32) | g_lo, ep, lane & 7);
        uint4 r[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            if (c < ResetDraws<V>::N) {
                r[c].x = __shfl_sync(full, mine.x, c); r[c].y = __shfl_sync(full, mine.y, c);
                r[c].z = __shfl_sync(full, mine.z, c); r[c].w = __shfl_sync(full, mine.w, c);
            } else {
                r[c] = make_uint4(0, 0, 0, 0);
            }
        }
        if ((int)lane == src) fill_reset<V>(P, r, e);
    }
}

// update_states_gate for one env (`:365-450`), written to a row in shared memory.
template <int V>
__device__ __forceinline__ void write_obs(const StepParams &P, const float *s_track, const EnvState<V> &e,
                                          uint32_t tg, float *o) {
    const int ng = P.n_gates;
    const float4 ga = *reinterpret_cast<const float4 *>(s_track + tg * kTrackRow);
    const float2 gb = *reinterpret_cast<const float2 *>(s_track + tg * kTrackRow + 4);
    const float c = gb.x, sn = gb.y;
    const float dx = sub_rn(e.x, ga.x), dy = sub_rn(e.y, ga.y);
    const float o0 = add_rn(mul_rn(dx, c), mul_rn(dy, sn));
    const float o1 = add_rn(mul_rn(dx, -sn), mul_rn(dy, c));
    const float o2 = sub_rn(e.z, ga.z);
    const float o3 = add_rn(mul_rn(e.vx, c), mul_rn(e.vy, sn));
    const float o4 = add_rn(mul_rn(e.vx, -sn), mul_rn(e.vy, c));
    const float yaw = wrap_yaw(sub_rn(e.psi, ga.w));
    if (V == kE2E) {
        float4 *o4p = reinterpret_cast<float4 *>(o);
        o4p[0] = make_float4(o0, o1, o2, o3);
        o4p[1] = make_float4(o4, e.vz, e.phi, e.th);
        o4p[2] = make_float4(yaw, e.p, e.q, e.r);
        o4p[3] = make_float4(e.w[0], e.w[1], e.w[2], e.w[V == kE2E ? 3 : 0]);
        uint32_t nx = tg;
        for (int i = 0; i < P.gates_ahead; ++i) {
            nx = (nx + 1 == (uint32_t)ng) ? 0u : nx + 1;
            o4p[4 + i] = *reinterpret_cast<const float4 *>(s_track + nx * kTrackRow + 8);
        }
        o4p[4 + P.gates_ahead] = make_float4(fmaf(e.dist[0], P.obs_scale[0], P.obs_off[0]),
                                             fmaf(e.dist[1], P.obs_scale[1], P.obs_off[1]),
                                             fmaf(e.dist[2], P.obs_scale[2], P.obs_off[2]),
                                             fmaf(e.dist[5], P.obs_scale[3], P.obs_off[3]));
    } else {
        o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = e.vz; o[6] = e.phi; o[7] = e.th;
        o[8] = yaw; o[9] = e.p; o[10] = e.q; o[11] = e.r; o[12] = e.w[0];
        uint32_t nx = tg;
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
__device__ __forceinline__ void residual_mlp(const StepParams &P, const float (&x)[10], float &thrust, float (&mom)[3]) {
    f32x2 xx[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) xx[k] = pack2(x[k], x[k]);
    // four hidden units per trip: each LDCU.128 of W1^T feeds two FFMA2
    f32x2 at = pack2(P.b2[0], 0.0f);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        f32x2 h01 = pack2(P.bt1[j], P.bt1[j + 1]), h23 = pack2(P.bt1[j + 2], P.bt1[j + 3]);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            h01 = fma2(pack2(P.wt1[k * 32 + j], P.wt1[k * 32 + j + 1]), xx[k], h01);
            h23 = fma2(pack2(P.wt1[k * 32 + j + 2], P.wt1[k * 32 + j + 3]), xx[k], h23);
        }
        float h0, h1, h2, h3;
        unpack2(h01, h0, h1);
        unpack2(h23, h2, h3);
        at = fma2(pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f)), pack2(P.wt2[j], P.wt2[j + 1]), at);
        at = fma2(pack2(fmaxf(h2, 0.0f), fmaxf(h3, 0.0f)), pack2(P.wt2[j + 2], P.wt2[j + 3]), at);
    }
    float lo, hi;
    unpack2(at, lo, hi);
    thrust = lo + hi;
    f32x2 a0 = pack2(P.b2[1], 0.0f), a1 = pack2(P.b2[2], 0.0f), a2 = pack2(P.b2[3], 0.0f);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        f32x2 h01 = pack2(P.bm1[j], P.bm1[j + 1]), h23 = pack2(P.bm1[j + 2], P.bm1[j + 3]);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            h01 = fma2(pack2(P.wm1[k * 32 + j], P.wm1[k * 32 + j + 1]), xx[k], h01);
            h23 = fma2(pack2(P.wm1[k * 32 + j + 2], P.wm1[k * 32 + j + 3]), xx[k], h23);
        }
        float h0, h1, h2, h3;
        unpack2(h01, h0, h1);
        unpack2(h23, h2, h3);
        const f32x2 r01 = pack2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f)), r23 = pack2(fmaxf(h2, 0.0f), fmaxf(h3, 0.0f));
        a0 = fma2(r01, pack2(P.wm2[j], P.wm2[j + 1]), a0);
        a0 = fma2(r23, pack2(P.wm2[j + 2], P.wm2[j + 3]), a0);
        a1 = fma2(r01, pack2(P.wm2[32 + j], P.wm2[32 + j + 1]), a1);
        a1 = fma2(r23, pack2(P.wm2[32 + j + 2], P.wm2[32 + j + 3]), a1);
        a2 = fma2(r01, pack2(P.wm2[64 + j], P.wm2[64 + j + 1]), a2);
        a2 = fma2(r23, pack2(P.wm2[64 + j + 2], P.wm2[64 + j + 3]), a2);
    }
    unpack2(a0, lo, hi); mom[0] = lo + hi;
    unpack2(a1, lo, hi); mom[1] = lo + hi;
    unpack2(a2, lo, hi); mom[2] = lo + hi;
}

// sin and cos together: 3-term Cody-Waite reduction by pi/2 and degree-7/8 minimax polynomials (Cephes
// coefficients), <= 1.5 ulp on |x| < 1e5 -- the same error class as NumPy's float32 sin/cos (1.45 ulp), see
// tests/test_kernel_math_models.py.  Huge arguments (a blown-up yaw) take the library's Payne-Hanek path
// out of line, so the hot code stays small and needs no stack frame.
__device__ __noinline__ void sincos_slow(float x, float *s, float *c) { sincosf(x, s, c); }

__device__ __forceinline__ void sincos_fast(float x, float &sn, float &cs) {
    if (fabsf(x) > 1.0e5f) { sincos_slow(x, &sn, &cs); return; }   // also NaN/Inf
    const float j = rintf(x * 0.636619772367581343f);
    const int q = (int)j;
    float a = fmaf(j, -1.5707962512969970703f, x);
    a = fmaf(j, -7.5497894158615963534e-08f, a);
    a = fmaf(j, -5.3903029534742383927e-15f, a);
    const float z = a * a;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float s = fmaf(a * z, ps, a);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    pc = fmaf(pc, z, -0.5f);
    const float c = fmaf(pc, z, 1.0f);
    const bool swap = q & 1;
    float ss = swap ? c : s, cc = swap ? s : c;
    sn = (q & 2) ? -ss : ss;
    cs = ((q + 1) & 2) ? -cc : cc;
}

// new = state + dt * f(state, action[, residual + disturbance])  (`:503-512`; INDI `:304`)
template <int V>
__device__ __forceinline__ void euler_step(const StepParams &P, const EnvState<V> &e, const float4 u, EnvState<V> &n) {
    const float dt = P.dt;
    float sph, cph, sth, cth, sps, cps;
    sincos_fast(e.phi, sph, cph);
    sincos_fast(e.th, sth, cth);
    sincos_fast(e.psi, sps, cps);
    // R = Rz*Ry*Rx
    const float r00 = cps * cth, r10 = sps * cth, r20 = -sth;
    const float r01 = sph * sth * cps - sps * cph, r11 = sph * sps * sth + cph * cps, r21 = sph * cth;
    const float r02 = sph * sps + sth * cph * cps, r12 = -sph * cps + sps * sth * cph, r22 = cph * cth;
    const float vbx = e.vx * r00 + e.vy * r10 + e.vz * r20;
    const float vby = e.vx * r01 + e.vy * r11 + e.vz * r21;
    const float vbz = e.vx * r02 + e.vy * r12 + e.vz * r22;

    float Dx, Dy, T, dp, dq, dr;
    if (V == kE2E) {
        const float w1 = e.w[0], w2 = e.w[1], w3 = e.w[2], w4 = e.w[V == kE2E ? 3 : 0];
        const float x[10] = {w1, w2, w3, w4, vbx, vby, vbz, e.p, e.q, e.r};
        float thr, mom[3];
        residual_mlp(P, x, thr, mom);
        const float Mx = mom[0] + e.dist[0], My = mom[1] + e.dist[1], Mz = mom[2] + e.dist[2];
        const float Fz = thr + e.dist[5];
        const float W1 = fmaf(4000.f, w1, 7000.f), W2 = fmaf(4000.f, w2, 7000.f);
        const float W3 = fmaf(4000.f, w3, 7000.f), W4 = fmaf(4000.f, w4, 7000.f);
        const float sumW = (W1 + W2) + (W3 + W4);
        const float q1 = W1 * W1, q2 = W2 * W2, q3 = W3 * W3, q4 = W4 * W4;
        T = Fz - 4.36301076e-8f * ((q1 + q2) + (q3 + q4)) - 0.0625501332f * (vbx * vbx + vby * vby) -
            2.7862899e-5f * vbz * sumW;
        Dx = e.dist[3] - 1.07933887e-5f * vbx * sumW;
        Dy = e.dist[4] - 9.65250793e-6f * vby * sumW;
        dp = 1103.7527593819f * Mx - 0.896247240618101f * e.q * e.r - 8.79803364238411f * vby +
             1.55842505518764e-6f * ((q1 - q2) + (q4 - q3));
        dq = 805.152979066023f * My + 0.924315619967794f * e.p * e.r + 10.4077084541063f * vbx +
             9.79081191626409e-7f * ((q1 - q3) + (q2 - q4));
        dr = 486.854917234664f * Mz - 0.163583252190847f * e.p * e.q - 0.395780237098345f * e.r +
             13.3373373580007f * ((u.y - u.x) + (u.w - u.z)) + 8.33177659850698f * ((w1 - w2) + (w3 - w4));
        n.w[0] = fmaf(dt, 16.6666666666667f * (u.x - w1), w1);
        n.w[1] = fmaf(dt, 16.6666666666667f * (u.y - w2), w2);
        n.w[2] = fmaf(dt, 16.6666666666667f * (u.z - w3), w3);
        n.w[V == kE2E ? 3 : 0] = fmaf(dt, 16.6666666666667f * (u.w - w4), w4);
#pragma unroll
        for (int k = 0; k < 6; ++k) n.dist[k] = e.dist[k];
    } else {
        const float Tn = e.w[0];
        T = fmaf(-8.0f, Tn, -8.0f);
        Dx = -0.33915248f * vbx;
        Dy = -0.4314916f * vby;
        dp = fmaf(-33.3333333333333f, e.p, 100.0f * u.x);
        dq = fmaf(-33.3333333333333f, e.q, 100.0f * u.y);
        dr = fmaf(-33.3333333333333f, e.r, 66.6666666666667f * u.z);
        n.w[0] = fmaf(dt, 33.3333333333333f * (u.w - Tn), Tn);
    }
    const float dvx = r00 * Dx + r01 * Dy + r02 * T;
    const float dvy = r10 * Dx + r11 * Dy + r12 * T;
    const float dvz = r20 * Dx + r21 * Dy + r22 * T + 9.81f;
    // Euler-angle kinematics (`:140-142`); tan = sin/cos with one IEEE reciprocal
    const float rc = __frcp_rn(cth);
    const float tth = sth * rc;
    const float qs_rc = e.q * sph + e.r * cph;
    const float dphi = fmaf(qs_rc, tth, e.p);
    const float dth = e.q * cph - e.r * sph;
    const float dpsi = qs_rc * rc;

    // positions: exactly the reference's two rounded float32 operations, so every threshold test agrees bit for bit
    n.x = add_rn(e.x, mul_rn(dt, e.vx));
    n.y = add_rn(e.y, mul_rn(dt, e.vy));
    n.z = add_rn(e.z, mul_rn(dt, e.vz));
    n.vx = fmaf(dt, dvx, e.vx); n.vy = fmaf(dt, dvy, e.vy); n.vz = fmaf(dt, dvz, e.vz);
    n.phi = fmaf(dt, dphi, e.phi); n.th = fmaf(dt, dth, e.th); n.psi = fmaf(dt, dpsi, e.psi);
    n.p = fmaf(dt, dp, e.p); n.q = fmaf(dt, dq, e.q); n.r = fmaf(dt, dr, e.r);
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
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ the step kernel
// Input tile of one pipeline stage: the 128 envs of a tile are 2 KB-contiguous in every plane, so a stage is
// filled by 6-8 bulk copies issued by ONE thread; nobody else spends an instruction on global loads.
template <int V> struct Stage;
template <> struct Stage<kE2E> {
    enum : int { P0 = 0, P1 = 2048, P2 = 4096, P3 = 6144, DA = 8192, DB = 10240, META = 11264, ACT = 11776, BYTES = 13824 };
};
template <> struct Stage<kINDI> {
    enum : int { P0 = 0, P1 = 2048, P2 = 4096, P3 = 6144, META = 6656, ACT = 7168, BYTES = 9216, DA = 0, DB = 0 };
};
constexpr int kBarBytes = 128;  // mbarriers live in the first 128 bytes of dynamic shared memory

__host__ __device__ constexpr size_t step_smem_bytes(int variant, int stages, int obs_len, int n_gates) {
    return kBarBytes + (size_t)stages * (variant == kE2E ? (int)Stage<kE2E>::BYTES : (int)Stage<kINDI>::BYTES) +
           (size_t)kBlock * obs_len * 4 + (size_t)n_gates * kTrackRow * 4;
}

template <int V>
__device__ __forceinline__ void issue_tile(const StepParams &P, unsigned char *st, uint64_t *bar, long long tile) {
    using S = Stage<V>;
    const long long base = tile * kBlock;
    const long long rem = P.n - base;
    const uint32_t act_bytes = (uint32_t)(rem < kBlock ? rem : kBlock) * 16u;  // caller's buffer is not padded
    mbar_expect_tx(bar, (uint32_t)S::BYTES - 2048u + act_bytes);
    bulk_load(st + S::P0, P.s.p0 + base, 2048, bar);
    bulk_load(st + S::P1, P.s.p1 + base, 2048, bar);
    bulk_load(st + S::P2, P.s.p2 + base, 2048, bar);
    if (V == kE2E) {
        bulk_load(st + S::P3, P.s.p3 + base, 2048, bar);
        bulk_load(st + S::DA, P.s.da + base, 2048, bar);
        bulk_load(st + S::DB, P.s.db + base, 1024, bar);
    } else {
        bulk_load(st + S::P3, P.s.p3s + base, 512, bar);
    }
    bulk_load(st + S::META, P.s.meta + base, 512, bar);
    bulk_load(st + S::ACT, P.actions + base, act_bytes, bar);
}

// Persistent CTAs (grid = SMs x resident CTAs), each looping over 128-env tiles through a kStages-deep ring of
// shared-memory stages: while tile i is being computed, the TMA engine is already filling the stages of the next
// tiles, so HBM latency never sits on a warp's scoreboard.  Per tile: 1 mbarrier wait, 2 block barriers.
template <int V, int kStages>
__global__ void __launch_bounds__(kBlock) step_kernel(const __grid_constant__ StepParams P) {
    using S = Stage<V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *stages = smem_raw + kBarBytes;
    float *s_obs = reinterpret_cast<float *>(stages + kStages * S::BYTES);
    float *s_track = s_obs + kBlock * P.obs_len;
    const int tid = threadIdx.x;
    const long long n_tiles = (P.n + kBlock - 1) / kBlock;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_track(P, s_track);  // ends with __syncthreads(): barrier init is visible to everyone
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_tiles) issue_tile<V>(P, stages + s * S::BYTES, &full[s], t);
        }
    }

    float reward_acc = 0.0f;  // stats are reduced once per CTA lifetime, not per tile
    unsigned c_act = 0, c_done = 0, c_tr = 0, c_gp = 0, c_gc = 0, c_gr = 0, c_ob = 0;
    const bool write_obs_tile = P.mode != kModePause;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = it % kStages;
        unsigned char *st = stages + stage * S::BYTES;
        const long long base = tile * kBlock;
        const long long env = base + tid;
        const bool active = env < P.n;
        mbar_wait(&full[stage], (uint32_t)(it / kStages) & 1u);

        // ---- stage -> registers (conflict-free: consecutive threads read consecutive 16-byte words)
        EnvState<V> e, n;
        {
            const float4 a = reinterpret_cast<const float4 *>(st + S::P0)[tid];
            const float4 b = reinterpret_cast<const float4 *>(st + S::P1)[tid];
            const float4 c = reinterpret_cast<const float4 *>(st + S::P2)[tid];
            e.x = a.x; e.y = a.y; e.z = a.z; e.vx = a.w; e.vy = b.x; e.vz = b.y; e.phi = b.z; e.th = b.w;
            e.psi = c.x; e.p = c.y; e.q = c.z; e.r = c.w;
            if (V == kE2E) {
                const float4 d = reinterpret_cast<const float4 *>(st + S::P3)[tid];
                const float4 da = reinterpret_cast<const float4 *>(st + S::DA)[tid];
                const float2 db = reinterpret_cast<const float2 *>(st + S::DB)[tid];
                e.w[0] = d.x; e.w[1] = d.y; e.w[2] = d.z; e.w[V == kE2E ? 3 : 0] = d.w;
                e.dist[0] = da.x; e.dist[1] = da.y; e.dist[2] = da.z; e.dist[5] = da.w; e.dist[3] = db.x; e.dist[4] = db.y;
            } else {
                e.w[0] = reinterpret_cast<const float *>(st + S::P3)[tid];
            }
        }
        const float4 u = reinterpret_cast<const float4 *>(st + S::ACT)[tid];
        const uint32_t meta = reinterpret_cast<const uint32_t *>(st + S::META)[tid];
        uint32_t tg = meta >> 24, sc = meta & kStepMask;
        if (tid == 0 && it > 0 && write_obs_tile) bulk_store_wait_read();  // previous tile's obs has left s_obs
        __syncthreads();  // [A] every thread holds its inputs: the stage and s_obs may be overwritten
        if (tid == 0) {
            const long long nt = tile + (long long)kStages * gridDim.x;
            if (nt < n_tiles) issue_tile<V>(P, st, &full[stage], nt);
        }

        euler_step<V>(P, e, u, n);
        sc = sc < kStepMask ? sc + 1 : sc;  // step_counts += 1 (`:514`), saturating in 24 bits

        // ---- reward and flags (`:516-566`)
        const float4 ga = *reinterpret_cast<const float4 *>(s_track + tg * kTrackRow);
        const float2 gcs = *reinterpret_cast<const float2 *>(s_track + tg * kTrackRow + 4);
        const float ox = sub_rn(e.x, ga.x), oy = sub_rn(e.y, ga.y), oz = sub_rn(e.z, ga.z);
        const float nx = sub_rn(n.x, ga.x), ny = sub_rn(n.y, ga.y), nz = sub_rn(n.z, ga.z);
        const float d_old = norm3_rn(ox, oy, oz), d_new = norm3_rn(nx, ny, nz);
        float reward = sub_rn(d_old, d_new);
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
        const bool dn = trunc | ground | collided | oob;
        const uint32_t fl = (dn ? F_DONE : 0u) | (trunc ? F_TRUNC : 0u) | (passed ? F_PASSED : 0u) |
                            (collided ? F_COLLISION : 0u) | (ground ? F_GROUND : 0u) | (oob ? F_OOB : 0u);

        // ---- branch logic (`:568-585`)
        bool write_world = active, write_dist = false;
        if (P.mode == kModeNormal) {
            if (P.reset_source == kResetDevice) {  // block-uniform branch: the whole warp takes part in the draw
                const bool need = dn && active;
                uint32_t ep = 0;
                if (need) ep = P.s.episode[env];
                draw_reset_warp<V>(P, env, ep, need, n);
                if (need) {
                    P.s.episode[env] = ep + 1;
                    tg = 0; sc = 0;
                    write_dist = (V == kE2E);
                }
            }
        } else if (P.mode == kModePauseIfCollision) {
            if (dn) { n = e; write_world = false; }
        } else {  // env.pause
            write_world = false;
        }
        if (active) {
            P.s.meta[env] = (tg << 24) | sc;
            P.rew[env] = reward;
            P.done[env] = (P.mode == kModePause) ? (uint8_t)0 : (uint8_t)(dn ? 1 : 0);
            if (P.flags) P.flags[env] = (uint8_t)fl;
        }
        if (write_world) store_world<V>(P.s, env, n);
        if (write_dist) store_dist<V>(P.s, env, n);

        if (write_obs_tile) {
            write_obs<V>(P, s_track, n, tg, s_obs + tid * P.obs_len);
            const long long rem = P.n - base;
            const int rows = rem < kBlock ? (int)rem : kBlock;
            float *dst = P.obs + base * P.obs_len;
            const uint32_t bytes = (uint32_t)rows * (uint32_t)P.obs_len * 4u;
            const bool bulk = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((bytes & 15u) == 0);
            if (bulk) {  // the tile's rows are contiguous in the (N,D) output: one TMA bulk store
                fence_proxy_async();
                __syncthreads();  // [B]
                if (tid == 0) bulk_store(dst, s_obs, bytes);
            } else {
                __syncthreads();
                for (int i = tid; i < rows * P.obs_len; i += kBlock) dst[i] = s_obs[i];
            }
        }
        if (P.stats && active) {
            reward_acc += reward;
            c_act += 1; c_done += (fl & F_DONE) != 0; c_tr += (fl & F_TRUNC) != 0; c_gp += (fl & F_PASSED) != 0;
            c_gc += (fl & F_COLLISION) != 0; c_gr += (fl & F_GROUND) != 0; c_ob += (fl & F_OOB) != 0;
        }
    }
    if (tid == 0 && write_obs_tile) bulk_store_wait_read();  // shared memory must outlive the last bulk read

    if (P.stats) {  // warp-level reduction of the reward and the flag counters, one atomic set per warp
        const unsigned full_mask = 0xffffffffu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) reward_acc += __shfl_xor_sync(full_mask, reward_acc, o);
        c_act = __reduce_add_sync(full_mask, c_act); c_done = __reduce_add_sync(full_mask, c_done);
        c_tr = __reduce_add_sync(full_mask, c_tr); c_gp = __reduce_add_sync(full_mask, c_gp);
        c_gc = __reduce_add_sync(full_mask, c_gc); c_gr = __reduce_add_sync(full_mask, c_gr);
        c_ob = __reduce_add_sync(full_mask, c_ob);
        if ((tid & 31) == 0 && c_act) {
            atomicAdd(&P.stats->reward_sum, (double)reward_acc);
            atomicAdd(&P.stats->env_steps, (unsigned long long)c_act);
            if (c_done) atomicAdd(&P.stats->dones, (unsigned long long)c_done);
            if (c_tr) atomicAdd(&P.stats->truncated, (unsigned long long)c_tr);
            if (c_gp) atomicAdd(&P.stats->gates_passed, (unsigned long long)c_gp);
            if (c_gc) atomicAdd(&P.stats->gate_collisions, (unsigned long long)c_gc);
            if (c_gr) atomicAdd(&P.stats->ground_collisions, (unsigned long long)c_gr);
            if (c_ob) atomicAdd(&P.stats->out_of_bounds, (unsigned long long)c_ob);
        }
    }
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
            const uint32_t ep = P.s.episode[env];
            draw_reset<V>(P, env, ep, e);
            P.s.episode[env] = ep + 1;
            tg = 0;
            P.s.meta[env] = 0;
            store_world<V>(P.s, env, e);
            store_dist<V>(P.s, env, e);
        } else {
            load_state<V>(P.s, env, e);
            tg = P.s.meta[env] >> 24;
        }
        write_obs<V>(P, s_track, e, tg, s_obs + threadIdx.x * P.obs_len);
    }
    const long long rem = P.n - base;
    store_obs_tile(P.obs + base * P.obs_len, s_obs, rem < kBlock ? (int)rem : kBlock, P.obs_len);
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
            const float4 da = P.s.da[env];
            const float2 db = P.s.db[env];
            e.dist[0] = da.x; e.dist[1] = da.y; e.dist[2] = da.z; e.dist[5] = da.w; e.dist[3] = db.x; e.dist[4] = db.y;
        }
    }
    store_world<V>(P.s, env, e);
    P.s.meta[env] = 0;
    float *row = s_obs + threadIdx.x * P.obs_len;
    write_obs<V>(P, s_track, e, 0u, row);
    float *dst = P.obs + env * P.obs_len;
    for (int j = 0; j < P.obs_len; ++j) dst[j] = row[j];
}

// ------------------------------------------------------------------------------------------------ AoS <-> planes
template <int V>
__global__ void import_kernel(Planes s, long long first, long long count, const float *ws, const float *dist,
                              const long long *tg, const long long *sc, int n_gates) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long env = first + k;
    constexpr int NS = V == kE2E ? 16 : 13;
    if (ws) {
        const float *r = ws + k * NS;
        s.p0[env] = make_float4(r[0], r[1], r[2], r[3]);
        s.p1[env] = make_float4(r[4], r[5], r[6], r[7]);
        s.p2[env] = make_float4(r[8], r[9], r[10], r[11]);
        if (V == kE2E) s.p3[env] = make_float4(r[12], r[13], r[14], r[V == kE2E ? 15 : 12]);
        else s.p3s[env] = r[12];
    }
    if (V == kE2E && dist) {
        const float *d = dist + k * 6;
        s.da[env] = make_float4(d[0], d[1], d[2], d[5]);
        s.db[env] = make_float2(d[3], d[4]);
    }
    if (tg || sc) {
        const uint32_t m = s.meta[env];
        uint32_t g = m >> 24, c = m & kStepMask;
        if (tg) { long long t = tg[k] % n_gates; if (t < 0) t += n_gates; g = (uint32_t)t; }
        if (sc) { long long v = sc[k]; c = v < 0 ? 0u : (v > (long long)kStepMask ? kStepMask : (uint32_t)v); }
        s.meta[env] = (g << 24) | c;
    }
}

template <int V>
__global__ void export_kernel(Planes s, long long first, long long count, float *ws, float *dist, long long *tg,
                              long long *sc) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long env = first + k;
    constexpr int NS = V == kE2E ? 16 : 13;
    if (ws) {
        float *r = ws + k * NS;
        const float4 a = s.p0[env], b = s.p1[env], c = s.p2[env];
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
        r[8] = c.x; r[9] = c.y; r[10] = c.z; r[11] = c.w;
        if (V == kE2E) { const float4 d = s.p3[env]; r[12] = d.x; r[13] = d.y; r[14] = d.z; r[V == kE2E ? 15 : 12] = d.w; }
        else r[12] = s.p3s[env];
    }
    if (V == kE2E && dist) {
        const float4 da = s.da[env];
        const float2 db = s.db[env];
        float *d = dist + k * 6;
        d[0] = da.x; d[1] = da.y; d[2] = da.z; d[3] = db.x; d[4] = db.y; d[5] = da.w;
    }
    if (tg || sc) {
        const uint32_t m = s.meta[env];
        if (tg) tg[k] = m >> 24;
        if (sc) sc[k] = m & kStepMask;
    }
}

}  // namespace qs
