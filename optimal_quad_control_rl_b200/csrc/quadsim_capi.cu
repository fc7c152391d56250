// quadsim_capi.cu -- host side of libquadsim.so: the C ABI declared in include/quadsim.h.
// Owns the device planes, builds the kernel parameter block, launches the sm_100a kernels of quadsim_kernels.cuh.
// No torch, no Python: plain pointers and sizes only.
#include "quadsim_kernels.cuh"
#include "quadsim_policy.cuh"
#include "quadsim_rollout.cuh"
#include "quadsim_train.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "quadsim.h"

using qs::Planes;
using qs::StepParams;

static_assert(sizeof(qs::Stats) == sizeof(qs_stats), "qs_stats layout");
static_assert(sizeof(StepParams) < 32000, "kernel parameter block too large");

struct qs_env {
    int variant = 0, device = 0, n_gates = 0, gates_ahead = 0, obs_len = 0, state_len = 0, block_bytes = 0;
    int64_t n = 0;
    cudaStream_t stream = nullptr;
    Planes planes{};
    float *track_dev = nullptr;
    qs::Stats *stats_dev = nullptr;
    bool stats_on = false, have_weights = false, track_dirty = true;
    std::vector<float> gate_pos, gate_yaw, gate_cos, gate_sin, pos_rel, yaw_rel;
    float start_pos[3] = {0, 0, 0};
    double ranges[12] = {0};
    int ranges_f64 = 0;
    double dist_scale = 1.0;
    int64_t max_steps = 1200;
    float dt = 0.01f;
    uint64_t seed = 0;
    unsigned long long *epoch_dev = nullptr;  // [0] launches that may draw resets so far (device RNG key), [1] CTA arrivals
    int64_t env_offset = 0;
    StepParams P{};  // weights live here
    // staging for the host-buffer entry points
    float *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr;
    float *act_stage = nullptr;  // pinned host staging for pageable / float64 actions (qs_step_host_ex)
    int host_threads = 1;        // OpenMP threads that fill it
    unsigned long long *info_dev = nullptr, *info_host = nullptr;  // step_info_kernel result (3 words) + its pinned copy
    cudaEvent_t ev_info = nullptr;
    uint8_t *h_done = nullptr, *h_flags = nullptr;
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    uint64_t launches = 0, chained_launches = 0;
    cudaStream_t copy_a = nullptr, copy_b = nullptr;  // host-buffer pipeline: upload+kernel / download
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
    int host_first_div = 4;  // first chunk = 1/div of an equal share (qs_step_host_ex): 2^20 envs 2092 -> 2037 us per step, 8: 2073, 2: 2080
    int host_chunks = 4;   // measured on B200 + PCIe Gen5: 1 -> 4.67e8, 4 -> 5.13e8 env-steps/s at N = 2^20
    int stages = 2, step_grid = 0;  // pipeline depth and persistent grid of the step kernel
    int stats_slots = 0;            // statistics / chain-ticket slots: one per CTA of the larger of the two step grids
    int l2_hints = 0;               // QS_L2_HINTS: state evict_last / streams evict_first in the step kernel
    double l2_keep_mb = 56.0;       // QS_L2_KEEP_MB: how much of the state to pin (one die's share of the 126 MB L2)
    bool pdl = true;                // programmatic dependent launch of consecutive steps
    // Chained launches (QS_CHAIN, default on): consecutive full-range steps captured into ONE CUDA graph depend on each
    // other CTA by CTA instead of grid by grid (StepParams::chain).  chain_tag / chain_capture say whether the previous
    // early-triggering launch of this library was this env's chained step inside the same capture.
    bool chain = true;
    unsigned int *chain_dev = nullptr;
    uint64_t chain_tag = 0;
    unsigned long long chain_capture = 0;
    cudaGraphNode_t chain_node = nullptr;  // the graph node of that launch: the next step chains only if it depends on exactly this node
    bool chain_x2 = false;          // kernel of the last chainable launch (a chain never spans two kernels)
    // E2E with two envs per thread (step_kernel_x2; QS_STEP_X2=0 turns it off): used when every warp of its grid gets
    // at least two warp-tiles and the launch needs none of the features only step_kernel has (see launch_step)
    bool x2 = false;
    int step_grid_x2 = 0;
    size_t step_smem_x2 = 0;
    size_t step_smem = 0;
    std::string err;
};

static std::string g_create_error;
// serial number of the last kernel this library launched with an early programmatic trigger (step, policy, rollout):
// a chained step may skip griddepcontrol.wait only if nothing else of ours was launched since its predecessor
static std::atomic<uint64_t> g_pdl_serial{0};

#define QS_CHECK_ENV(e) \
    if (!(e)) return QS_ERR_ARG
#define QS_CUDA(e, call)                                                                       \
    do {                                                                                       \
        cudaError_t _c = (call);                                                               \
        if (_c != cudaSuccess) {                                                               \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_c);                     \
            return QS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static int fail(qs_env *e, int code, const char *msg) {
    if (e) e->err = msg; else g_create_error = msg;
    return code;
}

extern "C" {

int qs_state_len(int variant) { return variant == QS_E2E ? 16 : 13; }
int qs_obs_len(int variant, int ga) { return (variant == QS_E2E ? 20 : 13) + 4 * ga; }
int qs_algorithmic_bytes_per_env_step(int variant, int ga) {
    return variant == QS_E2E ? 189 + 4 * (20 + 4 * ga) : 141 + 4 * (13 + 4 * ga);
}
const char *qs_version(void) { return "quadsim-b200 0.1 (sm_100a)"; }
const char *qs_last_error(const qs_env *e) { return e ? e->err.c_str() : g_create_error.c_str(); }
uint64_t qs_launch_count(const qs_env *e) { return e ? e->launches : 0; }
uint64_t qs_chained_launch_count(const qs_env *e) { return e ? e->chained_launches : 0; }

static void compute_track_tables(qs_env *e) {
    // Quadcopter3DGates.__init__ (`3D quad race.ipynb:309-319`): gate i expressed in the frame of gate i-1 (cyclic)
    const int ng = e->n_gates;
    e->gate_cos.resize(ng); e->gate_sin.resize(ng); e->pos_rel.resize(3 * ng); e->yaw_rel.resize(ng);
    for (int i = 0; i < ng; ++i) { e->gate_cos[i] = cosf(e->gate_yaw[i]); e->gate_sin[i] = sinf(e->gate_yaw[i]); }
    for (int i = 0; i < ng; ++i) {
        const int j = (i + ng - 1) % ng;
        const float dx = e->gate_pos[3 * i] - e->gate_pos[3 * j], dy = e->gate_pos[3 * i + 1] - e->gate_pos[3 * j + 1];
        const float c = e->gate_cos[j], s = e->gate_sin[j];
        volatile float a = c * dx, b = s * dy, cc = (-s) * dx, d = c * dy;  // separately rounded products
        e->pos_rel[3 * i] = a + b;
        e->pos_rel[3 * i + 1] = cc + d;
        e->pos_rel[3 * i + 2] = e->gate_pos[3 * i + 2] - e->gate_pos[3 * j + 2];
        e->yaw_rel[i] = e->gate_yaw[i] - e->gate_yaw[j];
    }
}

static int upload_track(qs_env *e) {
    if (!e->track_dirty) return QS_OK;
    std::vector<float> t((size_t)e->n_gates * qs::kTrackRow, 0.f);
    for (int g = 0; g < e->n_gates; ++g) {
        float *r = &t[(size_t)g * qs::kTrackRow];
        r[0] = e->gate_pos[3 * g]; r[1] = e->gate_pos[3 * g + 1]; r[2] = e->gate_pos[3 * g + 2]; r[3] = e->gate_yaw[g];
        r[4] = e->gate_cos[g]; r[5] = e->gate_sin[g];
        r[8] = e->pos_rel[3 * g]; r[9] = e->pos_rel[3 * g + 1]; r[10] = e->pos_rel[3 * g + 2]; r[11] = e->yaw_rel[g];
    }
    // stream-ordered with respect to kernels that read the old table
    QS_CUDA(e, cudaStreamSynchronize(e->stream));
    QS_CUDA(e, cudaMemcpy(e->track_dev, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    e->track_dirty = false;
    return QS_OK;
}

static void refresh_params(qs_env *e) {
    StepParams &P = e->P;
    P.s = e->planes;
    P.track = e->track_dev;
    P.stats = e->stats_on ? e->stats_dev : nullptr;
    P.n = e->n;
    P.env_offset = e->env_offset;
    P.seed = e->seed;
    P.epoch = e->epoch_dev;
    P.n_gates = e->n_gates;
    P.gates_ahead = e->gates_ahead;
    P.obs_len = e->obs_len;
    P.l2_hints = e->l2_hints;
    P.keep_blocks = (long long)(e->l2_keep_mb * 1048576.0 / (double)e->block_bytes);
    P.max_steps = e->max_steps < 0 ? 0u : (e->max_steps > 0xFFFFFF ? 0xFFFFFFFFu : (uint32_t)e->max_steps);
    P.dt = e->dt;
    static const int rows[4] = {0, 1, 2, 5};
    for (int k = 0; k < 4; ++k) {  // 2*(d-lo)/(hi-lo)-1 with the lo==hi widening of `:419-442`
        double lo = e->ranges[2 * rows[k]], hi = e->ranges[2 * rows[k] + 1];
        if (lo == hi) { lo -= 1; hi += 1; }
        // a float32 ranges array (the ctor default) is evaluated in float32 by the reference, a float64 one in float64
        const double span = e->ranges_f64 ? hi - lo : (double)((float)hi - (float)lo);
        P.obs_lo[k] = (float)lo;
        P.obs_lo2[k] = e->ranges_f64 ? (float)(lo - (double)P.obs_lo[k]) : 0.0f;
        P.obs_scale[k] = (float)(2.0 / span);
    }
    for (int k = 0; k < 3; ++k) P.rd.start[k] = e->start_pos[k];
    for (int k = 0; k < 6; ++k) {
        P.rd.dist_lo[k] = (float)(e->dist_scale * e->ranges[2 * k]);
        P.rd.dist_span[k] = (float)(e->dist_scale * (e->ranges[2 * k + 1] - e->ranges[2 * k]));
    }
}

#define QS_STEP_FN(V, H) \
    (e->stages == 3 ? (const void *)qs::step_kernel<V, 3, H> : e->stages == 4 ? (const void *)qs::step_kernel<V, 4, H> : (const void *)qs::step_kernel<V, 2, H>)
static const void *step_function(const qs_env *e) {
    if (e->variant == QS_E2E) return e->l2_hints == 1 ? QS_STEP_FN(qs::kE2E, true) : QS_STEP_FN(qs::kE2E, false);
    return e->l2_hints == 1 ? QS_STEP_FN(qs::kINDI, true) : QS_STEP_FN(qs::kINDI, false);
}
#undef QS_STEP_FN

static size_t smem_bytes(const qs_env *e) {
    return ((size_t)qs::kBlock * e->obs_len + (size_t)e->n_gates * qs::kTrackRow) * sizeof(float);
}
static unsigned grid_for(int64_t n) { return (unsigned)((n + qs::kBlock - 1) / qs::kBlock); }

static int ensure_scratch(qs_env *e, size_t bytes) {
    if (bytes <= e->scratch_bytes) return QS_OK;
    if (e->scratch) { QS_CUDA(e, cudaStreamSynchronize(e->stream)); cudaFree(e->scratch); e->scratch = nullptr; e->scratch_bytes = 0; }
    QS_CUDA(e, cudaMalloc(&e->scratch, bytes));
    e->scratch_bytes = bytes;
    return QS_OK;
}

int qs_create(qs_env **out, int variant, int64_t num_envs, int n_gates, const float *gate_pos, const float *gate_yaw,
              const float *start_pos, int gates_ahead, int device, void *stream) {
    if (!out) return fail(nullptr, QS_ERR_ARG, "qs_create: out is NULL");
    *out = nullptr;
    if (variant != QS_E2E && variant != QS_INDI) return fail(nullptr, QS_ERR_ARG, "qs_create: unknown variant");
    if (num_envs <= 0 || num_envs > (int64_t)1 << 31) return fail(nullptr, QS_ERR_ARG, "qs_create: num_envs out of range");
    if (n_gates <= 0 || n_gates > QS_MAX_GATES) return fail(nullptr, QS_ERR_ARG, "qs_create: n_gates must be 1..255");
    if (gates_ahead < 0 || gates_ahead > 16) return fail(nullptr, QS_ERR_ARG, "qs_create: gates_ahead must be 0..16");
    if (!gate_pos || !gate_yaw || !start_pos) return fail(nullptr, QS_ERR_ARG, "qs_create: NULL track pointer");
    qs_env *e = new (std::nothrow) qs_env();
    if (!e) return fail(nullptr, QS_ERR_NOMEM, "qs_create: out of host memory");
    e->variant = variant; e->device = device; e->n = num_envs; e->n_gates = n_gates; e->gates_ahead = gates_ahead;
    e->state_len = qs_state_len(variant); e->obs_len = qs_obs_len(variant, gates_ahead);
    e->stream = (cudaStream_t)stream;
    e->gate_pos.assign(gate_pos, gate_pos + 3 * n_gates);
    e->gate_yaw.assign(gate_yaw, gate_yaw + n_gates);
    memcpy(e->start_pos, start_pos, sizeof e->start_pos);
    compute_track_tables(e);

    auto bail = [&](cudaError_t c, const char *what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(c);
        qs_destroy(e);
        return (int)QS_ERR_CUDA;
    };
    cudaError_t c;
    if ((c = cudaSetDevice(device)) != cudaSuccess) return bail(c, "cudaSetDevice");
    // state: one AoSoA block per 32 envs, padded (zero-filled) to whole 128-env tiles so that every bulk load of
    // the step kernel is full-size
    const size_t n = ((size_t)num_envs + qs::kBlock - 1) / qs::kBlock * qs::kBlock;
    e->block_bytes = variant == QS_E2E ? (int)qs::Blk<qs::kE2E>::BYTES : (int)qs::Blk<qs::kINDI>::BYTES;
    Planes &s = e->planes;
    if ((c = cudaMalloc(&s.base, n / 32 * e->block_bytes)) != cudaSuccess) return bail(c, "cudaMalloc");
    cudaMemset(s.base, 0, n / 32 * e->block_bytes);
    if ((c = cudaMalloc(&e->epoch_dev, 16)) != cudaSuccess) return bail(c, "cudaMalloc");
    cudaMemset(e->epoch_dev, 0, 16);
    if ((c = cudaMalloc(&e->track_dev, (size_t)n_gates * qs::kTrackRow * 4)) != cudaSuccess) return bail(c, "cudaMalloc");
    if ((c = cudaDeviceSynchronize()) != cudaSuccess) return bail(c, "cudaDeviceSynchronize");
    // the observation tile + track table must fit in dynamic shared memory
    const size_t smem = smem_bytes(e);
    if (smem > 48 * 1024) {
        const void *fns[] = {(const void *)qs::observe_kernel<qs::kE2E>, (const void *)qs::observe_kernel<qs::kINDI>,
                             (const void *)qs::apply_reset_kernel<qs::kE2E>, (const void *)qs::apply_reset_kernel<qs::kINDI>};
        for (const void *f : fns)
            if ((c = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                return bail(c, "cudaFuncSetAttribute(smem)");
    }
    // persistent step kernel: pipeline depth, shared memory, grid = SMs x resident CTAs
    if (const char *sv = getenv("QS_STAGES")) e->stages = atoi(sv) == 3 ? 3 : (atoi(sv) == 4 ? 4 : 2);
    if (const char *pv = getenv("QS_PDL")) e->pdl = atoi(pv) != 0;
    if (const char *cv = getenv("QS_CHAIN")) e->chain = atoi(cv) != 0;
    // measured on B200 (profiles/r1/l2_pinning.md): INDI N = 2^20 41.9 -> 32.8 us/step, DRAM traffic per launch 177 ->
    // 136 MB.  E2E (96 MB of state + 122 MB of streams per step): 272 -> 258 MB and, with chained launches, 57.4 -> 56.6 us
    // (1 - 2 %, keep 40 - 72 MB alike; nothing without chaining, nothing at N = 262144: profiles/r2/step_kernel_ablation.md)
    e->l2_hints = 1;
    if (const char *hv = getenv("QS_L2_HINTS")) e->l2_hints = atoi(hv) != 0;
    if (const char *kv = getenv("QS_L2_KEEP_MB")) { double v = atof(kv); if (v >= 0.0 && v <= 4096.0) e->l2_keep_mb = v; }
    const void *step_fn = step_function(e);
    e->step_smem = qs::step_smem_bytes(variant, e->stages, e->obs_len, n_gates);
    if ((c = cudaFuncSetAttribute(step_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->step_smem)) != cudaSuccess)
        return bail(c, "cudaFuncSetAttribute(step smem)");
    int per_sm = 0, sms = 0;
    if ((c = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_fn, qs::step_warps(variant) * 32, e->step_smem)) != cudaSuccess)
        return bail(c, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if ((c = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
        return bail(c, "cudaDeviceGetAttribute");
    if (per_sm < 1) { g_create_error = "step kernel does not fit on an SM (gates_ahead / n_gates too large)"; qs_destroy(e); return QS_ERR_ARG; }
    if (const char *cv = getenv("QS_CTAS_PER_SM")) { int v = atoi(cv); if (v >= 1 && v < per_sm) per_sm = v; }
    const long long tiles = (num_envs + qs::kBlock - 1) / qs::kBlock;
    const long long want = (tiles * (qs::kBlock / 32) + qs::step_warps(variant) - 1) / qs::step_warps(variant);  // one warp-tile per warp
    e->step_grid = (int)(want < (long long)sms * per_sm ? want : (long long)sms * per_sm);
    // opt-in (QS_STEP_X2=1): bit-identical to step_kernel<e2e> but measured slower on B200 (75 vs 57 us per step at
    // N = 2^20: its 2x larger loop body does not stay in the instruction cache), see profiles/r2/step_kernel_ablation.md
    if (variant == QS_E2E && e->l2_hints == 0 && qs::step_x2_supports(e->obs_len))
        if (const char *xv = getenv("QS_STEP_X2")) e->x2 = atoi(xv) != 0;
    if (e->x2) {
        e->step_smem_x2 = qs::step_x2_smem_bytes(n_gates);
        int per_sm2 = 0;
        if ((c = cudaFuncSetAttribute((const void *)qs::step_kernel_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->step_smem_x2)) != cudaSuccess ||
            (c = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, (const void *)qs::step_kernel_x2, qs::kX2Warps * 32, e->step_smem_x2)) != cudaSuccess)
            return bail(c, "step_kernel_x2 setup");
        if (const char *cv = getenv("QS_X2_CTAS_PER_SM")) { int v = atoi(cv); if (v >= 1 && v < per_sm2) per_sm2 = v; }
        e->step_grid_x2 = sms * per_sm2;
        if (per_sm2 < 1) e->x2 = false;
    }
    // device-side totals: one slot per CTA, updated without atomics, summed on read
    const int slots = e->step_grid > e->step_grid_x2 ? e->step_grid : e->step_grid_x2;
    if ((c = cudaMalloc(&e->stats_dev, sizeof(qs::Stats) * slots)) != cudaSuccess) return bail(c, "cudaMalloc");
    if ((c = cudaMemset(e->stats_dev, 0, sizeof(qs::Stats) * slots)) != cudaSuccess) return bail(c, "cudaMemset");
    e->stats_slots = slots;
    if ((c = cudaMalloc(&e->chain_dev, sizeof(unsigned int) * 2 * slots)) != cudaSuccess) return bail(c, "cudaMalloc");
    if ((c = cudaMemset(e->chain_dev, 0, sizeof(unsigned int) * 2 * slots)) != cudaSuccess) return bail(c, "cudaMemset");
    *out = e;
    return QS_OK;
}

int qs_destroy(qs_env *e) {
    if (!e) return QS_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream); else cudaDeviceSynchronize();
    Planes &s = e->planes;
    cudaFree(s.base); cudaFree(e->epoch_dev); cudaFree(e->track_dev); cudaFree(e->stats_dev); cudaFree(e->chain_dev); cudaFree(e->scratch);
    cudaFree(e->h_act); cudaFree(e->h_obs); cudaFree(e->h_rew); cudaFree(e->h_done); cudaFree(e->h_flags);
    if (e->act_stage) cudaFreeHost(e->act_stage);
    cudaFree(e->info_dev);
    if (e->info_host) cudaFreeHost(e->info_host);
    if (e->ev_info) cudaEventDestroy(e->ev_info);
    if (e->copy_a) cudaStreamDestroy(e->copy_a);
    if (e->copy_b) cudaStreamDestroy(e->copy_b);
    if (e->ev_in) cudaEventDestroy(e->ev_in);
    if (e->ev_k) cudaEventDestroy(e->ev_k);
    if (e->ev_out) cudaEventDestroy(e->ev_out);
    delete e;
    return QS_OK;
}

int qs_set_stream(qs_env *e, void *stream) { QS_CHECK_ENV(e); e->stream = (cudaStream_t)stream; return QS_OK; }

int qs_set_track_tables(qs_env *e, const float *gc, const float *gs, const float *pr, const float *yr) {
    QS_CHECK_ENV(e);
    const int ng = e->n_gates;
    if (gc) e->gate_cos.assign(gc, gc + ng);
    if (gs) e->gate_sin.assign(gs, gs + ng);
    if (pr) e->pos_rel.assign(pr, pr + 3 * ng);
    if (yr) e->yaw_rel.assign(yr, yr + ng);
    e->track_dirty = true;
    return QS_OK;
}

int qs_get_track_tables(qs_env *e, float *gc, float *gs, float *pr, float *yr) {
    QS_CHECK_ENV(e);
    const int ng = e->n_gates;
    if (gc) memcpy(gc, e->gate_cos.data(), ng * 4);
    if (gs) memcpy(gs, e->gate_sin.data(), ng * 4);
    if (pr) memcpy(pr, e->pos_rel.data(), 3 * ng * 4);
    if (yr) memcpy(yr, e->yaw_rel.data(), ng * 4);
    return QS_OK;
}

int qs_set_max_steps(qs_env *e, int64_t m) { QS_CHECK_ENV(e); e->max_steps = m; return QS_OK; }
int qs_set_dt(qs_env *e, float dt) { QS_CHECK_ENV(e); e->dt = dt; return QS_OK; }
int qs_seed(qs_env *e, uint64_t seed) {  // also rewinds the launch epoch: same seed, same calls => same draws
    QS_CHECK_ENV(e);
    e->seed = seed;
    QS_CUDA(e, cudaSetDevice(e->device));
    QS_CUDA(e, cudaMemsetAsync(e->epoch_dev, 0, 16, e->stream));
    return QS_OK;
}
int qs_set_env_offset(qs_env *e, int64_t off) { QS_CHECK_ENV(e); e->env_offset = off; return QS_OK; }

int qs_set_disturbance_ranges(qs_env *e, const double *r, int is_f64, double scale) {
    QS_CHECK_ENV(e);
    if (!r) return fail(e, QS_ERR_ARG, "qs_set_disturbance_ranges: NULL");
    if (e->variant != QS_E2E) return fail(e, QS_ERR_STATE, "disturbances exist only in the E2E variant");
    memcpy(e->ranges, r, sizeof e->ranges);
    e->ranges_f64 = is_f64;
    e->dist_scale = scale;
    return QS_OK;
}

int qs_set_residual_weights(qs_env *e, const float *t, const float *m) {
    QS_CHECK_ENV(e);
    if (!t || !m) return fail(e, QS_ERR_ARG, "qs_set_residual_weights: NULL");
    if (e->variant != QS_E2E) return fail(e, QS_ERR_STATE, "residual MLPs exist only in the E2E variant");
    // torch layout (row-major [out][in], then bias) -> kernel layout (layer 1 transposed to [in][hidden])
    StepParams &P = e->P;
    for (int j = 0; j < 32; ++j) {
        for (int k = 0; k < 7; ++k) P.wt1[k * 32 + j] = t[j * 7 + k];
        for (int k = 0; k < 10; ++k) P.wm1[k * 32 + j] = m[j * 10 + k];
        P.bt1[j] = t[224 + j];
        P.wt2[j] = t[256 + j];
        P.bm1[j] = m[320 + j];
    }
    memcpy(P.wm2, m + 352, 96 * sizeof(float));
    P.b2[0] = t[288]; P.b2[1] = m[448]; P.b2[2] = m[449]; P.b2[3] = m[450];
    e->have_weights = true;
    return QS_OK;
}

int qs_set_obs_peers(qs_env *e, int n_peers, void *const *peer_obs_bases, int64_t row_offset) {
    QS_CHECK_ENV(e);
    if (n_peers < 0 || n_peers > 7 || (n_peers && !peer_obs_bases)) return fail(e, QS_ERR_ARG, "qs_set_obs_peers: 0..7 peers");
    for (int p = 0; p < 7; ++p) e->P.peer_obs[p] = p < n_peers ? (float *)peer_obs_bases[p] : nullptr;
    for (int p = 0; p < n_peers; ++p)
        if (!peer_obs_bases[p] || ((uintptr_t)peer_obs_bases[p] & 15)) return fail(e, QS_ERR_ARG, "qs_set_obs_peers: NULL or misaligned peer buffer");
    // the kernel sends a tile to the peers with the same TMA bulk store as to the local buffer, whose 16-byte
    // alignment it tests on the LOCAL destination only: the peer rows must be aligned the same way
    if (n_peers && (row_offset < 0 || ((row_offset * e->obs_len * 4) & 15)))
        return fail(e, QS_ERR_ARG, "qs_set_obs_peers: row_offset * obs_len * 4 must be a non-negative multiple of 16");
    if (n_peers && e->P.obs_packed && (row_offset % qs::kPolRows))
        return fail(e, QS_ERR_ARG, "qs_set_obs_peers: packed observations need row_offset % 128 == 0");
    e->P.n_peers = n_peers;
    e->P.peer_row_offset = row_offset;
    return QS_OK;
}

int qs_set_obs_format(qs_env *e, int format) {
    QS_CHECK_ENV(e);
    if (format != QS_OBS_F32 && format != QS_OBS_BF16_K32) return fail(e, QS_ERR_ARG, "qs_set_obs_format: unknown format");
    if (format == QS_OBS_BF16_K32 && e->obs_len > qs::kPackK - 1)
        return fail(e, QS_ERR_ARG, "qs_set_obs_format: packed observations need obs_len <= 31 (gates_ahead too large)");
    if (format == QS_OBS_BF16_K32 && e->P.n_peers && (e->P.peer_row_offset % qs::kPolRows))
        return fail(e, QS_ERR_ARG, "qs_set_obs_format: packed observations need a peer row_offset % 128 == 0");
    e->P.obs_packed = format == QS_OBS_BF16_K32 ? 1 : 0;
    return QS_OK;
}
int64_t qs_obs_packed_bytes(int obs_len, int64_t n) {
    if (n <= 0 || obs_len < 1 || obs_len > qs::kPackK - 1) return 0;
    return (n + qs::kPolRows - 1) / qs::kPolRows * (4 * (int64_t)qs::pack_block_bytes(obs_len));
}

int qs_enable_stats(qs_env *e, int on) { QS_CHECK_ENV(e); e->stats_on = on != 0; return QS_OK; }

int qs_get_stats(qs_env *e, qs_stats *out, int reset) {
    QS_CHECK_ENV(e);
    if (!out) return fail(e, QS_ERR_ARG, "qs_get_stats: NULL");
    QS_CUDA(e, cudaSetDevice(e->device));
    std::vector<qs_stats> slots((size_t)e->stats_slots);
    const size_t bytes = sizeof(qs_stats) * slots.size();
    QS_CUDA(e, cudaMemcpyAsync(slots.data(), e->stats_dev, bytes, cudaMemcpyDeviceToHost, e->stream));
    if (reset) QS_CUDA(e, cudaMemsetAsync(e->stats_dev, 0, bytes, e->stream));
    QS_CUDA(e, cudaStreamSynchronize(e->stream));
    qs_stats t{};
    for (const qs_stats &s : slots) {
        t.reward_sum += s.reward_sum; t.env_steps += s.env_steps; t.dones += s.dones; t.truncated += s.truncated;
        t.gates_passed += s.gates_passed; t.gate_collisions += s.gate_collisions;
        t.ground_collisions += s.ground_collisions; t.out_of_bounds += s.out_of_bounds;
    }
    *out = t;
    return QS_OK;
}

int qs_get_state_layout(qs_env *e, void **base, int *block_bytes, int *offsets7) {
    QS_CHECK_ENV(e);
    if (base) *base = e->planes.base;
    if (block_bytes) *block_bytes = e->block_bytes;
    if (offsets7) {
        using E = qs::Blk<qs::kE2E>;
        using I = qs::Blk<qs::kINDI>;
        const int oe[7] = {E::P0, E::P1, E::P2, E::P3, E::META, E::DA, E::DB}, oi[7] = {I::P0, I::P1, I::P2, I::P3, I::META, -1, -1};
        memcpy(offsets7, e->variant == QS_E2E ? oe : oi, sizeof oe);
    }
    return QS_OK;
}

// ---------------------------------------------------------------------------------------------- state import / export
static inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
static int check_slice(qs_env *e, int64_t first, int64_t count) {
    if (first < 0 || count < 0 || first + count > e->n) return fail(e, QS_ERR_ARG, "env slice out of range");
    return QS_OK;
}

int qs_set_state(qs_env *e, int64_t first, int64_t count, const float *ws, const float *dist, const int64_t *tg,
                 const int64_t *sc) {
    QS_CHECK_ENV(e);
    if (int r = check_slice(e, first, count)) return r;
    if (dist && e->variant != QS_E2E) return fail(e, QS_ERR_STATE, "disturbances exist only in the E2E variant");
    if (count == 0) return QS_OK;
    QS_CUDA(e, cudaSetDevice(e->device));
    const size_t b_ws = ws ? (size_t)count * e->state_len * 4 : 0, b_d = dist ? (size_t)count * 24 : 0;
    const size_t b_tg = tg ? (size_t)count * 8 : 0, b_sc = sc ? (size_t)count * 8 : 0;
    // staging sub-buffers start on 16-byte boundaries (13 floats x an odd count would misalign the int64 arrays)
    const size_t o_d = align16(b_ws), o_tg = o_d + align16(b_d), o_sc = o_tg + align16(b_tg);
    if (int r = ensure_scratch(e, o_sc + b_sc + 64)) return r;
    char *base = (char *)e->scratch;
    float *d_ws = (float *)base; float *d_d = (float *)(base + o_d);
    long long *d_tg = (long long *)(base + o_tg), *d_sc = (long long *)(base + o_sc);
    if (ws) QS_CUDA(e, cudaMemcpyAsync(d_ws, ws, b_ws, cudaMemcpyHostToDevice, e->stream));
    if (dist) QS_CUDA(e, cudaMemcpyAsync(d_d, dist, b_d, cudaMemcpyHostToDevice, e->stream));
    if (tg) QS_CUDA(e, cudaMemcpyAsync(d_tg, tg, b_tg, cudaMemcpyHostToDevice, e->stream));
    if (sc) QS_CUDA(e, cudaMemcpyAsync(d_sc, sc, b_sc, cudaMemcpyHostToDevice, e->stream));
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (e->variant == QS_E2E)
        qs::import_kernel<qs::kE2E><<<grid, 256, 0, e->stream>>>(e->planes, first, count, ws ? d_ws : nullptr,
                                                               dist ? d_d : nullptr, tg ? d_tg : nullptr,
                                                               sc ? d_sc : nullptr, e->n_gates);
    else
        qs::import_kernel<qs::kINDI><<<grid, 256, 0, e->stream>>>(e->planes, first, count, ws ? d_ws : nullptr, nullptr,
                                                                tg ? d_tg : nullptr, sc ? d_sc : nullptr, e->n_gates);
    e->launches++;
    QS_CUDA(e, cudaGetLastError());
    QS_CUDA(e, cudaStreamSynchronize(e->stream));
    return QS_OK;
}

int qs_get_state(qs_env *e, int64_t first, int64_t count, float *ws, float *dist, int64_t *tg, int64_t *sc) {
    QS_CHECK_ENV(e);
    if (int r = check_slice(e, first, count)) return r;
    if (dist && e->variant != QS_E2E) return fail(e, QS_ERR_STATE, "disturbances exist only in the E2E variant");
    if (count == 0) return QS_OK;
    QS_CUDA(e, cudaSetDevice(e->device));
    const size_t b_ws = ws ? (size_t)count * e->state_len * 4 : 0, b_d = dist ? (size_t)count * 24 : 0;
    const size_t b_tg = tg ? (size_t)count * 8 : 0, b_sc = sc ? (size_t)count * 8 : 0;
    // staging sub-buffers start on 16-byte boundaries (13 floats x an odd count would misalign the int64 arrays)
    const size_t o_d = align16(b_ws), o_tg = o_d + align16(b_d), o_sc = o_tg + align16(b_tg);
    if (int r = ensure_scratch(e, o_sc + b_sc + 64)) return r;
    char *base = (char *)e->scratch;
    float *d_ws = (float *)base; float *d_d = (float *)(base + o_d);
    long long *d_tg = (long long *)(base + o_tg), *d_sc = (long long *)(base + o_sc);
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (e->variant == QS_E2E)
        qs::export_kernel<qs::kE2E><<<grid, 256, 0, e->stream>>>(e->planes, first, count, ws ? d_ws : nullptr,
                                                               dist ? d_d : nullptr, tg ? d_tg : nullptr,
                                                               sc ? d_sc : nullptr);
    else
        qs::export_kernel<qs::kINDI><<<grid, 256, 0, e->stream>>>(e->planes, first, count, ws ? d_ws : nullptr, nullptr,
                                                                tg ? d_tg : nullptr, sc ? d_sc : nullptr);
    e->launches++;
    QS_CUDA(e, cudaGetLastError());
    if (ws) QS_CUDA(e, cudaMemcpyAsync(ws, d_ws, b_ws, cudaMemcpyDeviceToHost, e->stream));
    if (dist) QS_CUDA(e, cudaMemcpyAsync(dist, d_d, b_d, cudaMemcpyDeviceToHost, e->stream));
    if (tg) QS_CUDA(e, cudaMemcpyAsync(tg, d_tg, b_tg, cudaMemcpyDeviceToHost, e->stream));
    if (sc) QS_CUDA(e, cudaMemcpyAsync(sc, d_sc, b_sc, cudaMemcpyDeviceToHost, e->stream));
    QS_CUDA(e, cudaStreamSynchronize(e->stream));
    return QS_OK;
}

// ---------------------------------------------------------------------------------------------- hot path
static int prep_launch(qs_env *e, const char *who) {
    if (e->variant == QS_E2E && !e->have_weights) {
        e->err = std::string(who) + ": residual MLP weights not set (qs_set_residual_weights)";
        return QS_ERR_STATE;
    }
    QS_CUDA(e, cudaSetDevice(e->device));
    if (int r = upload_track(e)) return r;
    refresh_params(e);
    return QS_OK;
}

static int launch_observe(qs_env *e, float *obs_dev, int reset_all, const char *who) {
    if (!obs_dev) return fail(e, QS_ERR_ARG, "obs_dev is NULL");
    if (e->P.obs_packed && ((uintptr_t)obs_dev & 15)) return fail(e, QS_ERR_ARG, "a packed obs_dev must be 16-byte aligned");
    if (int r = prep_launch(e, who)) return r;
    e->P.obs = obs_dev;
    if (e->variant == QS_E2E)
        qs::observe_kernel<qs::kE2E><<<grid_for(e->n), qs::kBlock, smem_bytes(e), e->stream>>>(e->P, reset_all);
    else
        qs::observe_kernel<qs::kINDI><<<grid_for(e->n), qs::kBlock, smem_bytes(e), e->stream>>>(e->P, reset_all);
    e->launches++;
    QS_CUDA(e, cudaGetLastError());
    return QS_OK;
}

int qs_observe(qs_env *e, float *obs_dev) { QS_CHECK_ENV(e); return launch_observe(e, obs_dev, 0, "qs_observe"); }
int qs_reset_all(qs_env *e, float *obs_dev) { QS_CHECK_ENV(e); return launch_observe(e, obs_dev, 1, "qs_reset_all"); }

static int check_step_args(qs_env *e, const void *act, const void *obs, const void *rew, const void *done, int mode,
                           int reset_source, const char *who) {
    if (!act || !rew || !done) { e->err = std::string(who) + ": NULL buffer"; return QS_ERR_ARG; }
    if (mode < QS_MODE_NORMAL || mode > QS_MODE_PAUSE) { e->err = std::string(who) + ": bad mode"; return QS_ERR_ARG; }
    if (reset_source != QS_RESET_DEVICE && reset_source != QS_RESET_HOST) { e->err = std::string(who) + ": bad reset_source"; return QS_ERR_ARG; }
    if (mode != QS_MODE_PAUSE && !obs) { e->err = std::string(who) + ": obs buffer is NULL"; return QS_ERR_ARG; }
    return QS_OK;
}

// one launch of the step kernel over the 128-env tiles [t0, t1) on `stream`; e->P already holds the buffers
static int launch_step(qs_env *e, long long t0, long long t1, int advance_epoch, cudaStream_t stream, bool pdl) {
    StepParams &P = e->P;
    P.tile_begin = t0; P.tile_end = t1; P.advance_epoch = advance_epoch;
    const long long tiles = t1 - t0;
    // chained launch: the whole env range on the env's own stream, nothing stored into peers.  It skips the grid-wide
    // wait only inside a stream capture, behind another chained launch of this env in the SAME capture (a graph is
    // replayed in contexts the host cannot see, so its first step always waits for everything before it).
    const bool chainable = e->chain && e->chain_dev && pdl && advance_epoch && stream == e->stream && P.n_peers == 0 &&
                           t0 == 0 && t1 == (e->n + qs::kBlock - 1) / qs::kBlock;
    P.chain = chainable ? e->chain_dev : nullptr;
    P.chain_wait = 0;
    // two envs per thread (E2E): float32 rows, nothing stored into peers, and enough work for two warp-tiles per warp
    const bool x2 = e->x2 && !P.obs_packed && P.n_peers == 0 && P.l2_hints == 0 &&
                    tiles * (qs::kBlock / 32) >= 2LL * e->step_grid_x2 * qs::kX2Warps;
    if (chainable) {
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        unsigned long long cid = 0;
        const cudaGraphNode_t *deps = nullptr;
        size_t n_deps = 0;
        if (cudaStreamGetCaptureInfo(stream, &cst, &cid, nullptr, &deps, &n_deps) != cudaSuccess) { cudaGetLastError(); cst = cudaStreamCaptureStatusNone; }
        const bool capturing = cst == cudaStreamCaptureStatusActive;
        // ... and only if that launch is the ONLY thing this one depends on: a foreign kernel captured in between (it may
        // produce the actions) is then waited for as a whole, like everything else
        if (capturing && cid == e->chain_capture && e->chain_tag != 0 && g_pdl_serial.load() == e->chain_tag && e->chain_x2 == x2 &&
            n_deps == 1 && deps && deps[0] == e->chain_node && e->chain_node != nullptr)
            P.chain_wait = 1;
        e->chain_capture = capturing ? cid : 0ull;
        e->chain_x2 = x2;
    }
    // programmatic dependent launch: this grid's prologue may overlap the tail of the previous kernel on the
    // stream; the kernel itself waits (griddepcontrol.wait) before it touches simulator state
    void *args[] = {&P};
    cudaLaunchConfig_t cfg{};
    const int cta_warps = qs::step_warps(e->variant == QS_E2E ? qs::kE2E : qs::kINDI);
    const long long want = (tiles * (qs::kBlock / 32) + cta_warps - 1) / cta_warps;  // CTAs that give every warp-tile a warp
    cfg.gridDim = dim3((unsigned)(want < e->step_grid ? want : e->step_grid));
    cfg.blockDim = dim3((unsigned)cta_warps * 32);
    cfg.dynamicSmemBytes = e->step_smem;
    if (x2) {
        cfg.gridDim = dim3((unsigned)e->step_grid_x2);
        cfg.blockDim = dim3((unsigned)qs::kX2Warps * 32);
        cfg.dynamicSmemBytes = e->step_smem_x2;
    }
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QS_CUDA(e, cudaLaunchKernelExC(&cfg, x2 ? (const void *)qs::step_kernel_x2 : step_function(e), args));
    const uint64_t tag = ++g_pdl_serial;
    e->chain_tag = chainable ? tag : 0;
    e->chain_node = nullptr;
    if (chainable && e->chain_capture != 0) {  // remember the node this launch became
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        const cudaGraphNode_t *deps = nullptr;
        size_t n_deps = 0;
        if (cudaStreamGetCaptureInfo(stream, &cst, nullptr, nullptr, &deps, &n_deps) == cudaSuccess && n_deps == 1 && deps) e->chain_node = deps[0];
        else cudaGetLastError();
    }
    e->launches++;
    e->chained_launches += P.chain_wait ? 1 : 0;
    return QS_OK;
}

int qs_step(qs_env *e, const float *actions_dev, float *obs_dev, float *rew_dev, uint8_t *done_dev, uint8_t *flags_dev,
            int mode, int reset_source) {
    QS_CHECK_ENV(e);
    if (int r = check_step_args(e, actions_dev, obs_dev, rew_dev, done_dev, mode, reset_source, "qs_step")) return r;
    if ((uintptr_t)actions_dev & 15) return fail(e, QS_ERR_ARG, "qs_step: actions_dev must be 16-byte aligned");
    if (e->P.obs_packed && ((uintptr_t)obs_dev & 15)) return fail(e, QS_ERR_ARG, "qs_step: a packed obs_dev must be 16-byte aligned");
    if (e->P.obs_packed && reset_source == QS_RESET_HOST) return fail(e, QS_ERR_STATE, "qs_step: packed observations need the device reset");
    if (int r = prep_launch(e, "qs_step")) return r;
    StepParams &P = e->P;
    P.actions = reinterpret_cast<const float4 *>(actions_dev);
    P.obs = obs_dev; P.rew = rew_dev; P.done = done_dev; P.flags = flags_dev;
    P.mode = mode; P.reset_source = reset_source;
    if (int r = launch_step(e, 0, (e->n + qs::kBlock - 1) / qs::kBlock, 1, e->stream, e->pdl)) return r;
    QS_CUDA(e, cudaGetLastError());
    return QS_OK;
}

int qs_apply_reset(qs_env *e, int64_t count, const int32_t *idx, const float *ws, const float *dist, float *obs_dev) {
    QS_CHECK_ENV(e);
    if (count == 0) return QS_OK;
    if (count < 0 || count > e->n || !idx || !ws || !obs_dev) return fail(e, QS_ERR_ARG, "qs_apply_reset: bad argument");
    if (dist && e->variant != QS_E2E) return fail(e, QS_ERR_STATE, "disturbances exist only in the E2E variant");
    for (int64_t k = 0; k < count; ++k)
        if (idx[k] < 0 || idx[k] >= e->n) return fail(e, QS_ERR_ARG, "qs_apply_reset: env index out of range");
    if (int r = prep_launch(e, "qs_apply_reset")) return r;
    const size_t b_i = ((size_t)count * 4 + 15) & ~(size_t)15, b_ws = (size_t)count * e->state_len * 4;
    const size_t b_d = dist ? (size_t)count * 24 : 0;
    if (int r = ensure_scratch(e, b_i + b_ws + b_d + 64)) return r;
    char *base = (char *)e->scratch;
    int *d_i = (int *)base; float *d_ws = (float *)(base + b_i); float *d_d = (float *)(base + b_i + b_ws);
    QS_CUDA(e, cudaMemcpyAsync(d_i, idx, (size_t)count * 4, cudaMemcpyHostToDevice, e->stream));
    QS_CUDA(e, cudaMemcpyAsync(d_ws, ws, b_ws, cudaMemcpyHostToDevice, e->stream));
    if (dist) QS_CUDA(e, cudaMemcpyAsync(d_d, dist, b_d, cudaMemcpyHostToDevice, e->stream));
    e->P.obs = obs_dev;
    if (e->variant == QS_E2E)
        qs::apply_reset_kernel<qs::kE2E><<<grid_for(count), qs::kBlock, smem_bytes(e), e->stream>>>(e->P, count, d_i, d_ws,
                                                                                                 dist ? d_d : nullptr);
    else
        qs::apply_reset_kernel<qs::kINDI><<<grid_for(count), qs::kBlock, smem_bytes(e), e->stream>>>(e->P, count, d_i, d_ws,
                                                                                                  nullptr);
    e->launches++;
    QS_CUDA(e, cudaGetLastError());
    QS_CUDA(e, cudaStreamSynchronize(e->stream));
    return QS_OK;
}

// ---------------------------------------------------------------------------------------------- host-buffer entry points
static int ensure_io(qs_env *e) {
    if (e->h_act) return QS_OK;
    const size_t n = (size_t)e->n;
    QS_CUDA(e, cudaSetDevice(e->device));
    QS_CUDA(e, cudaMalloc(&e->h_act, n * 16));
    QS_CUDA(e, cudaMalloc(&e->h_obs, n * e->obs_len * 4));
    QS_CUDA(e, cudaMalloc(&e->h_rew, n * 4));
    QS_CUDA(e, cudaMalloc(&e->h_done, n));
    QS_CUDA(e, cudaMalloc(&e->h_flags, n));
    QS_CUDA(e, cudaStreamCreateWithFlags(&e->copy_a, cudaStreamNonBlocking));
    QS_CUDA(e, cudaStreamCreateWithFlags(&e->copy_b, cudaStreamNonBlocking));
    QS_CUDA(e, cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
    QS_CUDA(e, cudaEventCreateWithFlags(&e->ev_k, cudaEventDisableTiming));
    QS_CUDA(e, cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
    if (const char *cv = getenv("QS_HOST_CHUNKS")) { int v = atoi(cv); if (v >= 1 && v <= 64) e->host_chunks = v; }
    if (const char *cv = getenv("QS_HOST_FIRST_DIV")) { int v = atoi(cv); if (v >= 1 && v <= 64) e->host_first_div = v; }
    {   // staging threads of qs_step_host_ex: a few are enough to outrun PCIe; QS_HOST_THREADS overrides
        const unsigned hc = std::thread::hardware_concurrency();
        e->host_threads = hc >= 16 ? 6 : (hc >= 4 ? (int)hc / 2 : 1);
        if (const char *tv = getenv("QS_HOST_THREADS")) { int v = atoi(tv); if (v >= 1 && v <= 64) e->host_threads = v; }
    }
    return QS_OK;
}

// One step with HOST buffers, software-pipelined over chunks of envs so that PCIe runs in both directions at once:
// stream A carries  H2D(actions c) -> step kernel(chunk c)  and stream B the D2H of chunk c's outputs, so the
// upload of chunk c+1 overlaps the download of chunk c.  All chunk launches read the same RNG epoch; the last one
// advances it, so the result is bit-identical to one qs_step over device buffers.
// Actions may be float32 or float64 (the reference keeps whatever array the caller passed, `3D quad race.ipynb:498-499`)
// in pageable or pinned memory: pageable / float64 input is converted chunk by chunk into an internal pinned staging
// buffer while the previous chunk is on the bus.  `info` (optional) receives what step_wait's `infos` loop needs
// (`:589-594`): the highest done env index and whether any env was truncated, so the caller scans nothing.
static bool is_pinned_host(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int qs_step_host_ex(qs_env *e, const void *act, int act_dtype, float *obs, float *rew, uint8_t *done, uint8_t *flags,
                    int mode, int reset_source, qs_step_info *info) {
    QS_CHECK_ENV(e);
    if (int r = check_step_args(e, act, obs, rew, done, mode, reset_source, "qs_step_host")) return r;
    if (act_dtype != QS_F32 && act_dtype != QS_F64) return fail(e, QS_ERR_ARG, "qs_step_host: actions must be float32 or float64");
    if (info && !flags) return fail(e, QS_ERR_ARG, "qs_step_host: info needs the flags buffer");
    if (e->P.obs_packed) return fail(e, QS_ERR_STATE, "qs_step_host: the env's observation format is packed BF16 (qs_set_obs_format)");
    if (int r = ensure_io(e)) return r;
    if (int r = prep_launch(e, "qs_step_host")) return r;
    StepParams &P = e->P;
    P.actions = reinterpret_cast<const float4 *>(e->h_act);
    P.obs = e->h_obs; P.rew = e->h_rew; P.done = e->h_done; P.flags = flags ? e->h_flags : nullptr;
    P.mode = mode; P.reset_source = reset_source;
    const long long tiles = (e->n + qs::kBlock - 1) / qs::kBlock;
    const bool stage = act_dtype == QS_F64 || !is_pinned_host(act);
    // staging the actions on the host puts the first chunk's memcpy on the critical path: twice as many chunks then
    const int chunks = stage ? 2 * e->host_chunks : e->host_chunks;
    long long per = (tiles + chunks - 1) / chunks;
    if (per < 256) per = tiles < 256 ? tiles : 256;  // >= 32768 envs per chunk: below that the copies are latency-bound
    // The download is the long pole (101 B per env against 16 B up): nothing comes back before the first chunk's
    // upload + kernel are through, so the first chunk is a fraction of the others (QS_HOST_FIRST_DIV, 1 = equal chunks)
    long long first = per;
    if (e->host_first_div > 1 && chunks > 1 && per / e->host_first_div >= 256) {
        first = per / e->host_first_div;
        per = (tiles - first + chunks - 2) / (chunks - 1);
    }
    const bool want_obs = obs && mode != QS_MODE_PAUSE;
    if (stage && !e->act_stage) QS_CUDA(e, cudaHostAlloc((void **)&e->act_stage, (size_t)e->n * 16, cudaHostAllocDefault));
    QS_CUDA(e, cudaEventRecord(e->ev_in, e->stream));
    QS_CUDA(e, cudaStreamWaitEvent(e->copy_a, e->ev_in, 0));
    QS_CUDA(e, cudaStreamWaitEvent(e->copy_b, e->ev_in, 0));
    for (long long t0 = 0, t1 = 0; t0 < tiles; t0 = t1) {
        t1 = t0 + (t0 == 0 ? first : per);
        if (t1 > tiles) t1 = tiles;
        const size_t first_env = (size_t)t0 * qs::kBlock;
        const size_t cnt = (size_t)((t1 * qs::kBlock < e->n ? t1 * qs::kBlock : e->n)) - first_env;
        const float *src = (const float *)act + first_env * 4;
        if (stage) {  // the previous step's uploads completed before that call returned: the staging buffer is free
            // a few host threads per chunk: one core copies ~8 GB/s, the 16 MB of a 2^20-env action array would cost
            // as much as the whole PCIe transfer of the results
            float *dst = e->act_stage + first_env * 4;
            const long long pieces = (long long)((cnt * 16 + 262143) / 262144);  // 256 KB per piece
            const int threads = (int)(pieces < e->host_threads ? pieces : e->host_threads);
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
            for (long long p = 0; p < pieces; ++p) {
                const size_t a = (size_t)p * 16384, b = a + 16384 < cnt ? a + 16384 : cnt;  // envs [a, b) of this chunk
                if (act_dtype == QS_F64) {
                    const double *s64 = (const double *)act + first_env * 4;
                    for (size_t i = a * 4; i < b * 4; ++i) dst[i] = (float)s64[i];
                } else {
                    memcpy(dst + a * 4, src + a * 4, (b - a) * 16);
                }
            }
            src = dst;
        }
        QS_CUDA(e, cudaMemcpyAsync(e->h_act + first_env * 4, src, cnt * 16, cudaMemcpyHostToDevice, e->copy_a));
        if (int r = launch_step(e, t0, t1, t1 == tiles, e->copy_a, false)) return r;
        QS_CUDA(e, cudaEventRecord(e->ev_k, e->copy_a));
        QS_CUDA(e, cudaStreamWaitEvent(e->copy_b, e->ev_k, 0));
        if (want_obs)
            QS_CUDA(e, cudaMemcpyAsync(obs + first_env * e->obs_len, e->h_obs + first_env * e->obs_len, cnt * e->obs_len * 4,
                                       cudaMemcpyDeviceToHost, e->copy_b));
        QS_CUDA(e, cudaMemcpyAsync(rew + first_env, e->h_rew + first_env, cnt * 4, cudaMemcpyDeviceToHost, e->copy_b));
        QS_CUDA(e, cudaMemcpyAsync(done + first_env, e->h_done + first_env, cnt, cudaMemcpyDeviceToHost, e->copy_b));
        if (flags) QS_CUDA(e, cudaMemcpyAsync(flags + first_env, e->h_flags + first_env, cnt, cudaMemcpyDeviceToHost, e->copy_b));
    }
    if (info) {  // the `infos` scan (`:589-594`) as one tiny kernel over the flag bytes, behind the last chunk's step
        if (!e->info_dev) {
            QS_CUDA(e, cudaMalloc(&e->info_dev, 24));
            QS_CUDA(e, cudaHostAlloc((void **)&e->info_host, 24, cudaHostAllocDefault));
            QS_CUDA(e, cudaEventCreateWithFlags(&e->ev_info, cudaEventDisableTiming));
        }
        QS_CUDA(e, cudaMemsetAsync(e->info_dev, 0, 24, e->copy_a));
        qs::step_info_kernel<<<148, 256, 0, e->copy_a>>>(e->h_flags, e->n, e->info_dev);
        e->launches++;
        QS_CUDA(e, cudaMemcpyAsync(e->info_host, e->info_dev, 24, cudaMemcpyDeviceToHost, e->copy_a));
        QS_CUDA(e, cudaEventRecord(e->ev_info, e->copy_a));
        QS_CUDA(e, cudaStreamWaitEvent(e->copy_b, e->ev_info, 0));
    }
    QS_CUDA(e, cudaGetLastError());
    QS_CUDA(e, cudaEventRecord(e->ev_out, e->copy_b));
    QS_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_out, 0));  // later work on the handle's stream is ordered after us
    QS_CUDA(e, cudaEventSynchronize(e->ev_out));
    if (info) {
        const bool paused = mode == QS_MODE_PAUSE;  // env.pause clears the dones (`:570-572`), not the time-limit test
        info->last_done_index = paused ? -1 : (int64_t)e->info_host[0] - 1;
        info->n_done = paused ? 0 : (int64_t)e->info_host[1];
        info->any_truncated = e->info_host[2] ? 1 : 0;
    }
    return QS_OK;
}

int qs_step_host(qs_env *e, const float *act, float *obs, float *rew, uint8_t *done, uint8_t *flags, int mode,
                 int reset_source) {
    return qs_step_host_ex(e, act, QS_F32, obs, rew, done, flags, mode, reset_source, nullptr);
}

static int observe_host(qs_env *e, float *obs, int reset_all) {
    if (!obs) return fail(e, QS_ERR_ARG, "obs_host is NULL");
    if (e->P.obs_packed) return fail(e, QS_ERR_STATE, "host observations: the env's observation format is packed BF16 (qs_set_obs_format)");
    if (int r = ensure_io(e)) return r;
    if (int r = launch_observe(e, e->h_obs, reset_all, reset_all ? "qs_reset_all_host" : "qs_observe_host")) return r;
    QS_CUDA(e, cudaMemcpyAsync(obs, e->h_obs, (size_t)e->n * e->obs_len * 4, cudaMemcpyDeviceToHost, e->stream));
    QS_CUDA(e, cudaStreamSynchronize(e->stream));
    return QS_OK;
}
int qs_reset_all_host(qs_env *e, float *obs) { QS_CHECK_ENV(e); return observe_host(e, obs, 1); }
int qs_observe_host(qs_env *e, float *obs) { QS_CHECK_ENV(e); return observe_host(e, obs, 0); }

void *qs_host_alloc(size_t bytes) {
    void *p = nullptr;
    return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void qs_host_free(void *p) { if (p) cudaFreeHost(p); }


// ================================================================================================ on-device policy
// the trained controller MLP evaluated next to the simulator (quadsim_policy.cuh).
struct qs_policy {
    int device = 0, in_dim = 0, n_hidden = 0, hidden = 0, out_dim = 0, k1 = 0, grid = 0, groups = 4;
    cudaStream_t stream = nullptr;
    std::vector<std::vector<float>> W, b;  // host copies, torch layout [out][in]
    std::vector<bool> have;
    unsigned char *w_dev = nullptr;
    unsigned long long *epoch_dev = nullptr;
    bool dirty = true, pdl = true;
    bool ts = false;     // QS_POLICY_TS=1: qs_policy_forward runs policy_kernel_ts (activations in tensor memory; measured slower: 2 chains)
    int activation = 0;  // 0 ReLU, 1 tanh
    float obs_limit = 0.0f;  // qs_policy_set_obs_limit
    uint64_t seed = 0, launches = 0;
    int64_t env_offset = 0;
    float std[4] = {0, 0, 0, 0};
    size_t smem = 0;
    std::string err;
};

static std::string g_policy_create_error;
#define QS_PCHECK(p) \
    if (!(p)) return QS_ERR_ARG
static int pfail(qs_policy *p, int code, const char *msg) {
    if (p) p->err = msg; else g_policy_create_error = msg;
    return code;
}

static uint16_t bf16_rne(float f) {  // round-to-nearest-even, what cvt.rn.bf16.f32 does
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);  // inf / nan
    return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

// Weights -> BF16 in the canonical K-major UMMA layout: element (n, k) of a matrix with `rows` (padded) output rows
// lives in K-slab k/8 at byte (k/8)*rows*16 + n*16 + (k%8)*2.  Biases are folded in: the column that multiplies the
// constant-1 input (in_dim for layer 1, unit 127 afterwards) holds the bias, and row 127 of every hidden layer
// reproduces the constant.
static void pack_policy_weights(const qs_policy *p, std::vector<unsigned char> &blob) {
    blob.assign(qs::policy_weight_bytes(p->k1, p->n_hidden), 0);
    size_t off = 0;
    for (int l = 0; l <= p->n_hidden; ++l) {
        const bool last = l == p->n_hidden;
        const int rows = last ? qs::kPolOut : qs::kPolHidden, kk = l == 0 ? p->k1 : qs::kPolHidden;
        const int n_out = last ? p->out_dim : p->hidden, n_in = l == 0 ? p->in_dim : p->hidden;
        const int ones_k = l == 0 ? p->in_dim : qs::kPolOnes;
        auto put = [&](int n, int k, float v) {
            const uint16_t h = bf16_rne(v);
            memcpy(&blob[off + (size_t)(k / 8) * rows * 16 + (size_t)n * 16 + (size_t)(k % 8) * 2], &h, 2);
        };
        for (int n = 0; n < n_out; ++n) {
            for (int k = 0; k < n_in; ++k) put(n, k, p->W[l][(size_t)n * n_in + k]);
            put(n, ones_k, p->b[l][n]);
        }
        if (!last) put(qs::kPolOnes, ones_k, 1.0f);
        off += (size_t)(kk / 8) * rows * 16;
    }
}

const char *qs_policy_last_error(const qs_policy *p) { return p ? p->err.c_str() : g_policy_create_error.c_str(); }
uint64_t qs_policy_launch_count(const qs_policy *p) { return p ? p->launches : 0; }

int qs_policy_create(qs_policy **out, int in_dim, int n_hidden, int hidden_dim, int out_dim, int device, void *stream) {
    if (!out) return pfail(nullptr, QS_ERR_ARG, "qs_policy_create: out is NULL");
    *out = nullptr;
    if (in_dim < 1 || in_dim > 63) return pfail(nullptr, QS_ERR_ARG, "qs_policy_create: in_dim must be 1..63");
    if (n_hidden < 1 || n_hidden > 4) return pfail(nullptr, QS_ERR_ARG, "qs_policy_create: 1..4 hidden layers");
    if (hidden_dim < 1 || hidden_dim > qs::kPolOnes) return pfail(nullptr, QS_ERR_ARG, "qs_policy_create: hidden_dim must be 1..127");
    if (out_dim < 1 || out_dim > 4) return pfail(nullptr, QS_ERR_ARG, "qs_policy_create: out_dim must be 1..4");
    qs_policy *p = new (std::nothrow) qs_policy();
    if (!p) return pfail(nullptr, QS_ERR_NOMEM, "qs_policy_create: out of host memory");
    p->device = device; p->in_dim = in_dim; p->n_hidden = n_hidden; p->hidden = hidden_dim; p->out_dim = out_dim;
    p->k1 = (in_dim + 1 + 15) / 16 * 16;
    p->stream = (cudaStream_t)stream;
    p->W.resize(n_hidden + 1); p->b.resize(n_hidden + 1); p->have.assign(n_hidden + 1, false);
    auto bail = [&](cudaError_t c, const char *what) {
        g_policy_create_error = std::string(what) + ": " + cudaGetErrorString(c);
        qs_policy_destroy(p);
        return (int)QS_ERR_CUDA;
    };
    cudaError_t c;
    if ((c = cudaSetDevice(device)) != cudaSuccess) return bail(c, "cudaSetDevice");
    int sms = 0, smem_max = 0;
    if ((c = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail(c, "cudaDeviceGetAttribute");
    if ((c = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return bail(c, "cudaDeviceGetAttribute");
    // one CTA per SM; as many 128-row tile groups (each 32 KB of A operand + 128 TMEM columns) as fit beside the weights
    if (const char *gv = getenv("QS_POLICY_GROUPS")) { int v = atoi(gv); if (v >= 1 && v <= 4) p->groups = v; }
    while (p->groups > 1 && qs::policy_smem_bytes(p->k1, n_hidden, p->groups) > (size_t)smem_max) p->groups--;
    p->smem = qs::policy_smem_bytes(p->k1, n_hidden, p->groups);
    if (p->smem > (size_t)smem_max) { g_policy_create_error = "policy weights do not fit in shared memory"; qs_policy_destroy(p); return QS_ERR_ARG; }
    if ((c = cudaMalloc(&p->w_dev, qs::policy_weight_bytes(p->k1, n_hidden))) != cudaSuccess) return bail(c, "cudaMalloc");
    if ((c = cudaMalloc(&p->epoch_dev, 16)) != cudaSuccess) return bail(c, "cudaMalloc");
    if ((c = cudaMemset(p->epoch_dev, 0, 16)) != cudaSuccess) return bail(c, "cudaMemset");
    if ((c = cudaFuncSetAttribute((const void *)qs::policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem)) != cudaSuccess)
        return bail(c, "cudaFuncSetAttribute(policy smem)");
    // forward-only launches: the kernel that keeps the activations in tensor memory (quadsim_policy.cuh, policy_kernel_ts)
    if (const char *tv = getenv("QS_POLICY_TS")) p->ts = atoi(tv) != 0;
    if (qs::policy_ts_smem_bytes(p->k1, n_hidden) > (size_t)smem_max) p->ts = false;
    if (p->ts && (c = cudaFuncSetAttribute((const void *)qs::policy_kernel_ts, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)qs::policy_ts_smem_bytes(p->k1, n_hidden))) != cudaSuccess)
        return bail(c, "cudaFuncSetAttribute(policy_ts smem)");
    if (const char *pv = getenv("QS_PDL")) p->pdl = atoi(pv) != 0;
    p->grid = sms;
    *out = p;
    return QS_OK;
}

int qs_policy_destroy(qs_policy *p) {
    if (!p) return QS_OK;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream); else cudaDeviceSynchronize();
    cudaFree(p->w_dev); cudaFree(p->epoch_dev);
    delete p;
    return QS_OK;
}

int qs_policy_set_stream(qs_policy *p, void *stream) { QS_PCHECK(p); p->stream = (cudaStream_t)stream; return QS_OK; }
int qs_policy_set_env_offset(qs_policy *p, int64_t off) { QS_PCHECK(p); p->env_offset = off; return QS_OK; }
int qs_policy_seed(qs_policy *p, uint64_t seed) {
    QS_PCHECK(p);
    p->seed = seed;
    if (cudaSetDevice(p->device) != cudaSuccess || cudaMemsetAsync(p->epoch_dev, 0, 16, p->stream) != cudaSuccess)
        return pfail(p, QS_ERR_CUDA, "qs_policy_seed: CUDA error");
    return QS_OK;
}

int qs_policy_set_layer(qs_policy *p, int layer, const float *W, const float *b) {
    QS_PCHECK(p);
    if (layer < 0 || layer > p->n_hidden || !W || !b) return pfail(p, QS_ERR_ARG, "qs_policy_set_layer: bad argument");
    const int n_out = layer == p->n_hidden ? p->out_dim : p->hidden, n_in = layer == 0 ? p->in_dim : p->hidden;
    p->W[layer].assign(W, W + (size_t)n_out * n_in);
    p->b[layer].assign(b, b + n_out);
    p->have[layer] = true;
    p->dirty = true;
    return QS_OK;
}

// hidden activation (SB3 `activation_fn`; the reference's generated C carries both `nn_relu` and `nn_tanh`).  With tanh the
// bias-carrying constant unit cannot be re-emitted through the activation (tanh(1) != 1): it has to sit in the slab the
// kernels never overwrite, i.e. hidden_dim <= 120.
int qs_policy_set_activation(qs_policy *p, int activation) {
    QS_PCHECK(p);
    if (activation != QS_ACT_RELU && activation != QS_ACT_TANH) return pfail(p, QS_ERR_ARG, "qs_policy_set_activation: 0 = relu, 1 = tanh");
    if (activation == QS_ACT_TANH && p->hidden > qs::kPolHidden - 8) return pfail(p, QS_ERR_ARG, "qs_policy_set_activation: tanh needs hidden_dim <= 120");
    p->activation = activation;
    return QS_OK;
}

// PPO's learner sanitises what it reads from the rollout buffers (NaN -> 0, clamp to +-obs_limit: a tumbling quad's roll
// angle winds to ~1e8); a forward over those buffers (values, old log-probs) has to see the same inputs.  0 = off.
int qs_policy_set_obs_limit(qs_policy *p, float limit) {
    QS_PCHECK(p);
    if (!(limit >= 0.0f)) return pfail(p, QS_ERR_ARG, "qs_policy_set_obs_limit: limit must be >= 0 (0 = off)");
    p->obs_limit = limit;
    return QS_OK;
}

int qs_policy_set_std(qs_policy *p, const float *std4) {
    QS_PCHECK(p);
    if (!std4) return pfail(p, QS_ERR_ARG, "qs_policy_set_std: NULL");
    for (int k = 0; k < p->out_dim; ++k) p->std[k] = std4[k];
    return QS_OK;
}

// uploads the weights if they changed and fills the launch-invariant part of the kernel parameters
static int policy_params(qs_policy *p, int64_t n, int deterministic, cudaStream_t stream, qs::PolicyParams &P) {
    for (int l = 0; l <= p->n_hidden; ++l)
        if (!p->have[l]) return pfail(p, QS_ERR_STATE, "qs_policy_forward: a layer's weights are not set (qs_policy_set_layer)");
    if (cudaSetDevice(p->device) != cudaSuccess) return pfail(p, QS_ERR_CUDA, "cudaSetDevice failed");
    if (p->dirty) {
        std::vector<unsigned char> blob;
        pack_policy_weights(p, blob);
        if (cudaStreamSynchronize(stream) != cudaSuccess || cudaMemcpy(p->w_dev, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess)
            return pfail(p, QS_ERR_CUDA, "qs_policy_forward: weight upload failed");
        p->dirty = false;
    }
    P.weights = p->w_dev; P.epoch = p->epoch_dev;
    P.n = n; P.env_offset = p->env_offset; P.seed = p->seed; P.in_dim = p->in_dim; P.k1 = p->k1; P.n_hidden = p->n_hidden; P.hidden = p->hidden;
    P.out_dim = p->out_dim; P.deterministic = deterministic; P.activation = p->activation; P.obs_limit = p->obs_limit;
    P.weight_bytes = qs::policy_weight_bytes(p->k1, p->n_hidden);
    P.tmem_cols = p->groups <= 1 ? 128u : (p->groups == 2 ? 256u : 512u);
    for (int k = 0; k < 4; ++k) P.std[k] = p->std[k];
    return QS_OK;
}

static int policy_launch(qs_policy *p, const float *obs_dev, int64_t n, float *actions_dev, float *mean_dev, float *raw_dev,
                         int deterministic, cudaStream_t stream, bool packed = false) {
    qs::PolicyParams P{};
    if (int r = policy_params(p, n, deterministic, stream, P)) return r;
    P.obs = obs_dev; P.actions = actions_dev; P.mean = mean_dev; P.raw = raw_dev;
    P.obs_packed = packed ? 1 : 0;
    const bool ts = p->ts && !packed && p->obs_limit == 0.0f;  // (the TMEM-activation kernel reads plain float32 rows only)
    const long long tiles = (n + qs::kPolRows - 1) / qs::kPolRows;
    void *args[] = {&P};
    cudaLaunchConfig_t cfg{};
    const int per_cta = ts ? qs::kTsChains : p->groups;  // 128-row tiles in flight per CTA
    const long long ctas = (tiles + per_cta - 1) / per_cta;
    cfg.gridDim = dim3((unsigned)(ctas < p->grid ? ctas : p->grid));
    cfg.blockDim = dim3((unsigned)(ts ? qs::kTsChains * qs::kTsThreads : qs::kPolRows * p->groups));
    cfg.dynamicSmemBytes = ts ? qs::policy_ts_smem_bytes(p->k1, p->n_hidden) : p->smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = p->pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t c = cudaLaunchKernelExC(&cfg, ts ? (const void *)qs::policy_kernel_ts : (const void *)qs::policy_kernel, args);
    if (c != cudaSuccess) { p->err = std::string("policy_kernel launch: ") + cudaGetErrorString(c); return QS_ERR_CUDA; }
    ++g_pdl_serial;
    p->launches++;
    return QS_OK;
}

int qs_policy_forward(qs_policy *p, const float *obs_dev, int64_t n, float *actions_dev, float *mean_dev, float *raw_dev,
                      int deterministic) {
    QS_PCHECK(p);
    if (!obs_dev || !actions_dev || n <= 0) return pfail(p, QS_ERR_ARG, "qs_policy_forward: bad argument");
    if (((uintptr_t)actions_dev & 15) || ((uintptr_t)mean_dev & 15) || ((uintptr_t)raw_dev & 15))
        return pfail(p, QS_ERR_ARG, "qs_policy_forward: outputs must be 16-byte aligned");
    if ((p->in_dim & 3) == 0 && ((uintptr_t)obs_dev & 15)) return pfail(p, QS_ERR_ARG, "qs_policy_forward: obs_dev must be 16-byte aligned");
    return policy_launch(p, obs_dev, n, actions_dev, mean_dev, raw_dev, deterministic, p->stream);
}

int qs_policy_forward_packed(qs_policy *p, const void *packed_obs_dev, int64_t n, float *actions_dev, float *mean_dev,
                             float *raw_dev, int deterministic) {
    QS_PCHECK(p);
    if (!packed_obs_dev || !actions_dev || n <= 0) return pfail(p, QS_ERR_ARG, "qs_policy_forward_packed: bad argument");
    if (p->k1 > qs::kPackK) return pfail(p, QS_ERR_ARG, "qs_policy_forward_packed: packed observations need in_dim <= 31");
    if (((uintptr_t)packed_obs_dev & 15) || ((uintptr_t)actions_dev & 15) || ((uintptr_t)mean_dev & 15) || ((uintptr_t)raw_dev & 15))
        return pfail(p, QS_ERR_ARG, "qs_policy_forward_packed: buffers must be 16-byte aligned");
    return policy_launch(p, (const float *)packed_obs_dev, n, actions_dev, mean_dev, raw_dev, deterministic, p->stream, true);
}

// collect_rollouts on the device (SB3 `OnPolicyAlgorithm.collect_rollouts`, called from `3D quad race.ipynb:820`): for
// t < steps:  actions[t] = policy(obs[t]);  obs[t+1], rewards[t], dones[t] = env.step(actions[t]).  2*steps kernel
// launches enqueued back to back on the env's stream (PDL-chained), no host round trip.  obs[0] must hold the
// current observations (qs_reset_all / the previous rollout's obs[steps]).
int qs_rollout(qs_env *e, qs_policy *p, int steps, float *obs_buf, float *act_buf, float *raw_buf, float *rew_buf,
               uint8_t *done_buf, uint8_t *flags_buf, int deterministic) {
    QS_CHECK_ENV(e);
    if (!p) return fail(e, QS_ERR_ARG, "qs_rollout: policy is NULL");
    if (steps < 1 || !obs_buf || !act_buf || !rew_buf || !done_buf) return fail(e, QS_ERR_ARG, "qs_rollout: bad argument");
    if (p->in_dim != e->obs_len) return fail(e, QS_ERR_ARG, "qs_rollout: policy input width != observation width");
    if (e->P.obs_packed) return fail(e, QS_ERR_STATE, "qs_rollout: the rollout buffers hold float32 rows; the env's observation format is packed BF16");
    if (p->out_dim != 4) return fail(e, QS_ERR_ARG, "qs_rollout: the env takes 4 actions");
    if (p->device != e->device) return fail(e, QS_ERR_ARG, "qs_rollout: env and policy live on different devices");
    const size_t n = (size_t)e->n;
    for (int t = 0; t < steps; ++t) {
        float *obs_t = obs_buf + (size_t)t * n * e->obs_len, *act_t = act_buf + (size_t)t * n * 4;
        float *raw_t = raw_buf ? raw_buf + (size_t)t * n * 4 : nullptr;
        if (int r = policy_launch(p, obs_t, e->n, act_t, nullptr, raw_t, deterministic, e->stream)) { e->err = p->err; return r; }
        if (int r = qs_step(e, act_t, obs_t + n * e->obs_len, rew_buf + (size_t)t * n, done_buf + (size_t)t * n,
                            flags_buf ? flags_buf + (size_t)t * n : nullptr, QS_MODE_NORMAL, QS_RESET_DEVICE)) return r;
    }
    return QS_OK;
}

// The same rollout as ONE launch of the fused closed-loop kernel (quadsim_rollout.cuh): every tile group keeps its 128
// quads in registers for all `steps`, the controller runs on the tensor cores in between, and only the rollout
// buffers are written to HBM.  Same arguments, same results (bit for bit) as qs_rollout.
int qs_rollout_fused_supported(const qs_env *e, const qs_policy *p) {
    if (!e || !p) return 0;
    if (p->in_dim != e->obs_len || p->out_dim != 4 || p->device != e->device) return 0;
    if (!qs::rollout_fused_fits(p->k1, e->obs_len)) return 0;
    int smem_max = 0;
    if (cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device) != cudaSuccess) return 0;
    return qs::rollout_smem_bytes(p->k1, p->n_hidden, p->groups, e->n_gates) + 1024 <= (size_t)smem_max ? 1 : 0;
}

int qs_rollout_fused(qs_env *e, qs_policy *p, int steps, float *obs_buf, float *act_buf, float *raw_buf, float *rew_buf,
                     uint8_t *done_buf, uint8_t *flags_buf, int deterministic) {
    QS_CHECK_ENV(e);
    if (!p) return fail(e, QS_ERR_ARG, "qs_rollout_fused: policy is NULL");
    if (steps < 1 || !obs_buf || !act_buf || !rew_buf || !done_buf) return fail(e, QS_ERR_ARG, "qs_rollout_fused: bad argument");
    if (((uintptr_t)act_buf & 15) || ((uintptr_t)raw_buf & 15)) return fail(e, QS_ERR_ARG, "qs_rollout_fused: action buffers must be 16-byte aligned");
    if (e->P.obs_packed) return fail(e, QS_ERR_STATE, "qs_rollout_fused: the rollout buffers hold float32 rows; the env's observation format is packed BF16");
    if (!qs_rollout_fused_supported(e, p))
        return fail(e, QS_ERR_ARG, "qs_rollout_fused: this env / policy shape does not fit the fused kernel (use qs_rollout)");
    if (int r = prep_launch(e, "qs_rollout_fused")) return r;
    qs::RolloutParams R{};
    if (int r = policy_params(p, e->n, deterministic, e->stream, R.Q)) { e->err = p->err; return r; }
    R.S = e->P;
    R.S.mode = QS_MODE_NORMAL; R.S.reset_source = QS_RESET_DEVICE;
    R.obs_buf = obs_buf; R.act_buf = act_buf; R.raw_buf = raw_buf; R.rew_buf = rew_buf; R.done_buf = done_buf;
    R.flags_buf = flags_buf;
    R.steps = steps;
    const void *fn = e->variant == QS_E2E ? (const void *)qs::rollout_kernel<qs::kE2E> : (const void *)qs::rollout_kernel<qs::kINDI>;
    const size_t smem = qs::rollout_smem_bytes(p->k1, p->n_hidden, p->groups, e->n_gates);
    QS_CUDA(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (e->n + qs::kPolRows - 1) / qs::kPolRows;
    const long long ctas = (tiles + p->groups - 1) / p->groups;
    void *args[] = {&R};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ctas < p->grid ? ctas : p->grid));
    cfg.blockDim = dim3((unsigned)(qs::kPolRows * p->groups));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (e->pdl && p->pdl) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QS_CUDA(e, cudaLaunchKernelExC(&cfg, fn, args));
    ++g_pdl_serial;
    e->launches++;
    p->launches++;
    return QS_OK;
}

// compute_returns_and_advantage on the device (see gae_kernel).  rew/adv/ret (steps, n) f32, val (steps+1, n) f32,
// done (steps, n) u8; asynchronous on `stream`.
int qs_gae(const float *rew_dev, const float *val_dev, const uint8_t *done_dev, float *adv_dev, float *ret_dev, int64_t n,
           int steps, float gamma, float lambda, void *stream) {
    if (!rew_dev || !val_dev || !done_dev || !adv_dev || !ret_dev || n <= 0 || steps <= 0) return QS_ERR_ARG;
    qs::gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rew_dev, val_dev, done_dev, adv_dev, ret_dev, n,
                                                                                 steps, gamma, lambda);
    return cudaGetLastError() == cudaSuccess ? QS_OK : QS_ERR_CUDA;
}


// ================================================================================================ PPO update (row f2)
// SB3's PPO.train() for the reference's configuration as three kernels per minibatch (quadsim_train.cuh).
struct qs_trainer {
    int device = 0, in_dim = 0, k1 = 0, hidden = 0, n_ctas = 0, nf = 0;
    cudaStream_t stream = nullptr;
    float *param = nullptr, *m = nullptr, *v = nullptr, *grad = nullptr;  // (2*nf + 4): policy net, value net, log_std
    unsigned char *w_pi = nullptr, *w_vf = nullptr;                       // BF16 UMMA-layout blobs
    float *partial = nullptr, *stats = nullptr;
    double *mb = nullptr;  // [0..2] minibatch statistics, [3] squared gradient norm
    std::vector<float> h_param;
    bool dirty = true;     // host parameters changed: upload + re-pack before the next minibatch
    uint64_t step = 0, launches = 0;
    size_t smem = 0;
    std::string err;
};
static std::string g_trainer_create_error;
static int tfail(qs_trainer *t, int code, const char *msg) {
    if (t) t->err = msg; else g_trainer_create_error = msg;
    return code;
}
#define QS_TCUDA(t, call)                                                                      \
    do {                                                                                       \
        cudaError_t _c = (call);                                                               \
        if (_c != cudaSuccess) {                                                               \
            (t)->err = std::string(#call) + ": " + cudaGetErrorString(_c);                     \
            return QS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

const char *qs_trainer_last_error(const qs_trainer *t) { return t ? t->err.c_str() : g_trainer_create_error.c_str(); }
uint64_t qs_trainer_launch_count(const qs_trainer *t) { return t ? t->launches : 0; }

// position of layer `layer` (0..3) element (out, in) inside a network's padded parameter block; bias = the column
// that multiplies the constant-1 input (in_dim for layer 0, unit 127 afterwards); the output layer is stored transposed
static size_t tr_index(const qs_trainer *t, int layer, int out, int in) {
    const int w1 = qs::tr_w1_floats(t->k1), wh = qs::tr_wh_floats();
    if (layer == 0) return (size_t)out * t->k1 + in;
    if (layer < 3) return (size_t)w1 + (size_t)(layer - 1) * wh + (size_t)out * qs::kPolHidden + in;
    return (size_t)w1 + 2 * (size_t)wh + (size_t)in * qs::kPolOut + out;
}

int qs_trainer_create(qs_trainer **out, int in_dim, int hidden_dim, int device, void *stream) {
    if (!out) return tfail(nullptr, QS_ERR_ARG, "qs_trainer_create: out is NULL");
    *out = nullptr;
    if (in_dim < 1 || in_dim > 63) return tfail(nullptr, QS_ERR_ARG, "qs_trainer_create: in_dim must be 1..63");
    if (hidden_dim < 1 || hidden_dim > qs::kPolOnes) return tfail(nullptr, QS_ERR_ARG, "qs_trainer_create: hidden_dim must be 1..127");
    qs_trainer *t = new (std::nothrow) qs_trainer();
    if (!t) return tfail(nullptr, QS_ERR_NOMEM, "qs_trainer_create: out of host memory");
    t->device = device; t->in_dim = in_dim; t->hidden = hidden_dim; t->k1 = (in_dim + 1 + 15) / 16 * 16;
    t->stream = (cudaStream_t)stream;
    t->nf = qs::tr_net_floats(t->k1);
    auto bail = [&](cudaError_t c, const char *what) {
        g_trainer_create_error = std::string(what) + ": " + cudaGetErrorString(c);
        qs_trainer_destroy(t);
        return (int)QS_ERR_CUDA;
    };
    cudaError_t c;
    if ((c = cudaSetDevice(device)) != cudaSuccess) return bail(c, "cudaSetDevice");
    int sms = 0, smem_max = 0;
    if ((c = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail(c, "cudaDeviceGetAttribute");
    if ((c = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return bail(c, "cudaDeviceGetAttribute");
    t->smem = qs::train_smem_bytes(t->k1);
    if (t->smem > (size_t)smem_max) { g_trainer_create_error = "training tile does not fit in shared memory"; qs_trainer_destroy(t); return QS_ERR_ARG; }
    t->n_ctas = sms & ~1;  // even: half the CTAs own the policy network, half the value network
    const size_t np = 2 * (size_t)t->nf + 4, wb = qs::policy_weight_bytes(t->k1, 3);
    t->h_param.assign(np, 0.0f);
    for (int net = 0; net < 2; ++net)  // the rows that re-emit the constant 1 of the bias folding
        for (int l = 0; l < 3; ++l) t->h_param[(size_t)net * t->nf + tr_index(t, l, qs::kPolOnes, l == 0 ? in_dim : qs::kPolOnes)] = 1.0f;
    if ((c = cudaMalloc(&t->param, np * 4)) != cudaSuccess || (c = cudaMalloc(&t->m, np * 4)) != cudaSuccess ||
        (c = cudaMalloc(&t->v, np * 4)) != cudaSuccess || (c = cudaMalloc(&t->grad, np * 4)) != cudaSuccess ||
        (c = cudaMalloc(&t->w_pi, wb)) != cudaSuccess || (c = cudaMalloc(&t->w_vf, wb)) != cudaSuccess ||
        (c = cudaMalloc(&t->partial, (size_t)t->n_ctas * (t->nf + qs::kTrStats) * 4)) != cudaSuccess ||
        (c = cudaMalloc(&t->stats, 8 * 4)) != cudaSuccess || (c = cudaMalloc(&t->mb, 4 * 8)) != cudaSuccess)
        return bail(c, "cudaMalloc");
    cudaMemset(t->m, 0, np * 4); cudaMemset(t->v, 0, np * 4); cudaMemset(t->grad, 0, np * 4); cudaMemset(t->stats, 0, 32);
    cudaMemset(t->w_pi, 0, wb); cudaMemset(t->w_vf, 0, wb);
    if ((c = cudaFuncSetAttribute((const void *)qs::ppo_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem)) != cudaSuccess)
        return bail(c, "cudaFuncSetAttribute(train smem)");
    if ((c = cudaDeviceSynchronize()) != cudaSuccess) return bail(c, "cudaDeviceSynchronize");
    *out = t;
    return QS_OK;
}

int qs_trainer_destroy(qs_trainer *t) {
    if (!t) return QS_OK;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream); else cudaDeviceSynchronize();
    cudaFree(t->param); cudaFree(t->m); cudaFree(t->v); cudaFree(t->grad); cudaFree(t->w_pi); cudaFree(t->w_vf);
    cudaFree(t->partial); cudaFree(t->stats); cudaFree(t->mb);
    delete t;
    return QS_OK;
}

int qs_trainer_set_stream(qs_trainer *t, void *stream) { if (!t) return QS_ERR_ARG; t->stream = (cudaStream_t)stream; return QS_OK; }

static int trainer_download(qs_trainer *t) {  // device master parameters -> host mirror
    QS_TCUDA(t, cudaSetDevice(t->device));
    if (t->dirty) return QS_OK;  // host copy is the newer one
    QS_TCUDA(t, cudaMemcpyAsync(t->h_param.data(), t->param, t->h_param.size() * 4, cudaMemcpyDeviceToHost, t->stream));
    QS_TCUDA(t, cudaStreamSynchronize(t->stream));
    return QS_OK;
}

// torch layout in: W (out, in) row-major, b (out).  layer 0..2 hidden, 3 = output (policy: 4 rows, value: 1 row)
int qs_trainer_set_layer(qs_trainer *t, int net, int layer, const float *W, const float *b) {
    if (!t) return QS_ERR_ARG;
    if (net < 0 || net > 1 || layer < 0 || layer > 3 || !W || !b) return tfail(t, QS_ERR_ARG, "qs_trainer_set_layer: bad argument");
    if (int r = trainer_download(t)) return r;
    const int n_out = layer == 3 ? (net == 0 ? 4 : 1) : t->hidden, n_in = layer == 0 ? t->in_dim : t->hidden;
    const int ones = layer == 0 ? t->in_dim : qs::kPolOnes;
    float *base = t->h_param.data() + (size_t)net * t->nf;
    for (int o = 0; o < n_out; ++o) {
        for (int i = 0; i < n_in; ++i) base[tr_index(t, layer, o, i)] = W[(size_t)o * n_in + i];
        base[tr_index(t, layer, o, ones)] = b[o];
    }
    t->dirty = true;
    return QS_OK;
}

int qs_trainer_get_layer(qs_trainer *t, int net, int layer, float *W, float *b) {
    if (!t) return QS_ERR_ARG;
    if (net < 0 || net > 1 || layer < 0 || layer > 3 || !W || !b) return tfail(t, QS_ERR_ARG, "qs_trainer_get_layer: bad argument");
    if (int r = trainer_download(t)) return r;
    const int n_out = layer == 3 ? (net == 0 ? 4 : 1) : t->hidden, n_in = layer == 0 ? t->in_dim : t->hidden;
    const int ones = layer == 0 ? t->in_dim : qs::kPolOnes;
    const float *base = t->h_param.data() + (size_t)net * t->nf;
    for (int o = 0; o < n_out; ++o) {
        for (int i = 0; i < n_in; ++i) W[(size_t)o * n_in + i] = base[tr_index(t, layer, o, i)];
        b[o] = base[tr_index(t, layer, o, ones)];
    }
    return QS_OK;
}

int qs_trainer_set_log_std(qs_trainer *t, const float *ls4) {
    if (!t || !ls4) return QS_ERR_ARG;
    if (int r = trainer_download(t)) return r;
    for (int k = 0; k < 4; ++k) t->h_param[2 * (size_t)t->nf + k] = ls4[k];
    t->dirty = true;
    return QS_OK;
}
int qs_trainer_get_log_std(qs_trainer *t, float *ls4) {
    if (!t || !ls4) return QS_ERR_ARG;
    if (int r = trainer_download(t)) return r;
    for (int k = 0; k < 4; ++k) ls4[k] = t->h_param[2 * (size_t)t->nf + k];
    return QS_OK;
}

int qs_trainer_reset_optimizer(qs_trainer *t) {
    if (!t) return QS_ERR_ARG;
    QS_TCUDA(t, cudaSetDevice(t->device));
    const size_t np = 2 * (size_t)t->nf + 4;
    QS_TCUDA(t, cudaMemsetAsync(t->m, 0, np * 4, t->stream));
    QS_TCUDA(t, cudaMemsetAsync(t->v, 0, np * 4, t->stream));
    t->step = 0;
    return QS_OK;
}

static void pack_blob_from_params(const qs_trainer *t, int net, std::vector<unsigned char> &blob) {
    blob.assign(qs::policy_weight_bytes(t->k1, 3), 0);
    const float *base = t->h_param.data() + (size_t)net * t->nf;
    size_t off = 0;
    for (int l = 0; l < 4; ++l) {
        const int rows = l == 3 ? qs::kPolOut : qs::kPolHidden, kk = l == 0 ? t->k1 : qs::kPolHidden;
        for (int n = 0; n < rows; ++n)
            for (int k = 0; k < kk; ++k) {
                const uint16_t h = bf16_rne(base[tr_index(t, l, n, k)]);
                memcpy(&blob[off + (size_t)(k / 8) * rows * 16 + (size_t)n * 16 + (size_t)(k % 8) * 2], &h, 2);
            }
        off += (size_t)(kk / 8) * rows * 16;
    }
}

static int trainer_upload(qs_trainer *t) {
    if (!t->dirty) return QS_OK;
    QS_TCUDA(t, cudaStreamSynchronize(t->stream));
    QS_TCUDA(t, cudaMemcpy(t->param, t->h_param.data(), t->h_param.size() * 4, cudaMemcpyHostToDevice));
    std::vector<unsigned char> blob;
    for (int net = 0; net < 2; ++net) {
        pack_blob_from_params(t, net, blob);
        QS_TCUDA(t, cudaMemcpy(net == 0 ? t->w_pi : t->w_vf, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    }
    t->dirty = false;
    return QS_OK;
}

// One minibatch: statistics -> gradients of both networks -> reduce -> (apply != 0) clip + Adam + new BF16 weights.
// All pointers are device pointers into the flat (total, .) rollout buffers; idx_dev (rows) int64 or NULL.
int qs_trainer_minibatch(qs_trainer *t, const int64_t *idx_dev, int64_t rows, const float *obs, const float *act,
                         const float *old_logp, const float *adv, const float *ret, const float *weight,
                         const qs_train_hyper *h, int apply) {
    if (!t) return QS_ERR_ARG;
    if (rows <= 0 || !obs || !act || !old_logp || !adv || !ret || !h) return tfail(t, QS_ERR_ARG, "qs_trainer_minibatch: bad argument");
    if ((uintptr_t)act & 15) return tfail(t, QS_ERR_ARG, "qs_trainer_minibatch: act must be 16-byte aligned");
    QS_TCUDA(t, cudaSetDevice(t->device));
    if (int r = trainer_upload(t)) return r;
    QS_TCUDA(t, cudaMemsetAsync(t->mb, 0, 32, t->stream));
    qs::ppo_mbstats_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, t->stream>>>((const long long *)idx_dev, rows, adv, weight, t->mb);
    qs::TrainParams P{};
    P.idx = (const long long *)idx_dev; P.rows = rows; P.obs = obs; P.act = act; P.old_logp = old_logp; P.adv = adv; P.ret = ret;
    P.weight = weight; P.mb = t->mb; P.w_pi = t->w_pi; P.w_vf = t->w_vf; P.log_std = t->param + 2 * (size_t)t->nf;
    P.partial = t->partial; P.in_dim = t->in_dim; P.k1 = t->k1; P.normalize_adv = h->normalize_advantage;
    P.clip_range = h->clip_range; P.vf_coef = h->vf_coef; P.obs_limit = h->obs_limit; P.act_limit = h->act_limit;
    P.weight_bytes = qs::policy_weight_bytes(t->k1, 3);
    const long long tiles = (rows + qs::kPolRows - 1) / qs::kPolRows;
    int ctas = (int)(2 * tiles < t->n_ctas ? 2 * tiles : t->n_ctas);
    qs::ppo_grad_kernel<<<ctas, qs::kPolRows, t->smem, t->stream>>>(P);
    qs::AdamParams A{};
    A.partial = t->partial; A.n_ctas = ctas; A.k1 = t->k1; A.mb = t->mb; A.grad = t->grad; A.norm2 = t->mb + 3;
    A.stats_out = t->stats; A.param = t->param; A.m = t->m; A.v = t->v; A.w_pi = t->w_pi; A.w_vf = t->w_vf;
    A.lr = h->learning_rate; A.beta1 = h->beta1; A.beta2 = h->beta2; A.eps = h->eps; A.max_grad_norm = h->max_grad_norm;
    A.ent_coef = h->ent_coef;
    const int total = 2 * t->nf + 8;
    qs::ppo_reduce_kernel<<<(total + 255) / 256, 256, 0, t->stream>>>(A);
    t->launches += 3;
    if (apply) {
        t->step++;
        A.bc1 = (float)(1.0 - pow((double)h->beta1, (double)t->step));
        A.bc2 = (float)(1.0 - pow((double)h->beta2, (double)t->step));
        qs::ppo_adam_kernel<<<(2 * t->nf + 4 + 255) / 256, 256, 0, t->stream>>>(A);
        t->launches++;
    }
    QS_TCUDA(t, cudaGetLastError());
    return QS_OK;
}

// gradients of the last minibatch in torch layout (tests): W (out, in), b (out)
int qs_trainer_get_grad(qs_trainer *t, int net, int layer, float *W, float *b, float *log_std4) {
    if (!t) return QS_ERR_ARG;
    QS_TCUDA(t, cudaSetDevice(t->device));
    std::vector<float> g(2 * (size_t)t->nf + 4);
    QS_TCUDA(t, cudaMemcpyAsync(g.data(), t->grad, g.size() * 4, cudaMemcpyDeviceToHost, t->stream));
    QS_TCUDA(t, cudaStreamSynchronize(t->stream));
    if (log_std4) for (int k = 0; k < 4; ++k) log_std4[k] = g[2 * (size_t)t->nf + k];
    if (W && b) {
        if (net < 0 || net > 1 || layer < 0 || layer > 3) return tfail(t, QS_ERR_ARG, "qs_trainer_get_grad: bad argument");
        const int n_out = layer == 3 ? (net == 0 ? 4 : 1) : t->hidden, n_in = layer == 0 ? t->in_dim : t->hidden;
        const int ones = layer == 0 ? t->in_dim : qs::kPolOnes;
        const float *base = g.data() + (size_t)net * t->nf;
        for (int o = 0; o < n_out; ++o) {
            for (int i = 0; i < n_in; ++i) W[(size_t)o * n_in + i] = base[tr_index(t, layer, o, i)];
            b[o] = base[tr_index(t, layer, o, ones)];
        }
    }
    return QS_OK;
}

// accumulated loss statistics since the last reset: pg_loss, v_loss, clip_frac, approx_kl (sums of minibatch means),
// grad_norm (sum), 3 unused
int qs_trainer_get_stats(qs_trainer *t, float *out8, int reset) {
    if (!t || !out8) return QS_ERR_ARG;
    QS_TCUDA(t, cudaSetDevice(t->device));
    QS_TCUDA(t, cudaMemcpyAsync(out8, t->stats, 32, cudaMemcpyDeviceToHost, t->stream));
    if (reset) QS_TCUDA(t, cudaMemsetAsync(t->stats, 0, 32, t->stream));
    QS_TCUDA(t, cudaStreamSynchronize(t->stream));
    return QS_OK;
}

// publish the policy network to the device actor without a host round trip: same BF16 blob layout
int qs_trainer_publish(qs_trainer *t, qs_policy *p) {
    if (!t || !p) return QS_ERR_ARG;
    if (p->in_dim != t->in_dim || p->n_hidden != 3 || p->hidden != t->hidden || p->out_dim != 4 || p->device != t->device)
        return tfail(t, QS_ERR_ARG, "qs_trainer_publish: policy shape differs from the trainer's");
    if (int r = trainer_upload(t)) return r;
    QS_TCUDA(t, cudaMemcpyAsync(p->w_dev, t->w_pi, qs::policy_weight_bytes(t->k1, 3), cudaMemcpyDeviceToDevice, t->stream));
    float ls[4];
    QS_TCUDA(t, cudaMemcpyAsync(ls, t->param + 2 * (size_t)t->nf, 16, cudaMemcpyDeviceToHost, t->stream));
    QS_TCUDA(t, cudaStreamSynchronize(t->stream));
    for (int k = 0; k < 4; ++k) p->std[k] = expf(ls[k]);
    p->dirty = false;
    for (int l = 0; l <= p->n_hidden; ++l) p->have[l] = true;
    return QS_OK;
}

}  // extern "C"
