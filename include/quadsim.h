/* quadsim.h -- C ABI of libquadsim.so: the B200 (sm_100a) quadrotor racing simulator.
 *
 * This is the drop-in boundary that sits UNDER the reference's Python `Quadcopter3DGates(VecEnv)` class
 * (`3D quad race.ipynb:287-620` end-to-end Bebop model; `3D quad race INDI inner loop.ipynb:142-410` INDI variant).
 * The reference has no FFI of its own (it is NumPy in notebook cells); the entry points below are the calls its
 * env methods decompose into, and `INTEGRATION.md` shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain C, no exceptions: every call returns QS_OK (0) or a negative qs_status; qs_last_error() explains.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are ordinary host memory.
 *   - the caller owns every buffer it passes; the handle owns the simulator state.
 *   - all kernels are enqueued on the handle's stream (qs_set_stream); calls taking host pointers synchronise
 *     that stream before returning, calls taking only device pointers are asynchronous.
 *   - a handle is not thread-safe.  One process per GPU; a shard of a larger job sets qs_set_env_offset so that
 *     the device RNG is keyed by the GLOBAL env index and results do not depend on the number of GPUs.
 *
 * State lives on the device as 32-env blocks of float4 planes (array of structs of arrays, DESIGN.md "Data layout"); the
 * reference's array-of-structs views (`world_states (N,16|13)`, `disturbances (N,6)`, `target_gates`,
 * `step_counts`) are produced on demand by qs_get_state / consumed by qs_set_state.
 */
#ifndef QUADSIM_H
#define QUADSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qs_env qs_env; /* opaque */

typedef enum {
    QS_OK = 0,
    QS_ERR_ARG = -1,    /* bad argument (NULL, out of range, misaligned) */
    QS_ERR_CUDA = -2,   /* a CUDA runtime call failed; text in qs_last_error */
    QS_ERR_STATE = -3,  /* call not valid for this handle (e.g. disturbances on an INDI env) */
    QS_ERR_NOMEM = -4
} qs_status;

/* which model: f_func of `3D quad race.ipynb:65-152` (+ residual MLPs `:244-262`) or of the INDI notebook `:48-110` */
typedef enum { QS_E2E = 0, QS_INDI = 1 } qs_variant;

/* the three branches of step_wait (`3D quad race.ipynb:568-585`) */
typedef enum {
    QS_MODE_NORMAL = 0,             /* world_states = new_states; reset_(dones)                         `:581-585` */
    QS_MODE_PAUSE_IF_COLLISION = 1, /* advance only envs that are not done, never reset                 `:573-580` */
    QS_MODE_PAUSE = 2               /* env.pause: nothing advances, dones reported False, obs untouched `:570-572` */
} qs_mode;

/* who provides the reset draws of reset_ (`:452-493`) in QS_MODE_NORMAL */
typedef enum {
    QS_RESET_DEVICE = 0, /* fused in the step kernel: Philox4x32-10 keyed by (seed, global env, launch epoch) -- fast path */
    QS_RESET_HOST = 1    /* step leaves done envs un-reset; the host draws (np.random, reference order) and calls
                            qs_apply_reset -- bit-for-bit the reference's reset values */
} qs_reset_source;

/* bits of the optional per-env flag byte written by qs_step */
enum {
    QS_F_DONE = 1,           /* max_steps | ground | gate_collision | out_of_bounds   `:566` */
    QS_F_TRUNCATED = 2,      /* step_counts >= max_steps                              `:553` */
    QS_F_GATE_PASSED = 4,    /*                                                       `:533` */
    QS_F_GATE_COLLISION = 8, /*                                                       `:534` */
    QS_F_GROUND = 16,        /* z_new > 0                                             `:543` */
    QS_F_OUT_OF_BOUNDS = 32  /* |x|,|y| > 10 or |p|,|q|,|r| > 1000                    `:549` */
};

/* running totals accumulated on the device by qs_step when enabled with qs_enable_stats */
typedef struct {
    double reward_sum;
    uint64_t env_steps;
    uint64_t dones;
    uint64_t truncated;
    uint64_t gates_passed;
    uint64_t gate_collisions;
    uint64_t ground_collisions;
    uint64_t out_of_bounds;
} qs_stats;

#define QS_MAX_GATES 255
#define QS_THRUST_WEIGHTS 289 /* Linear(7,32): W[32][7], b[32]; Linear(32,1): W[1][32], b[1]  (row-major [out][in]) */
#define QS_MOMENT_WEIGHTS 451 /* Linear(10,32): W[32][10], b[32]; Linear(32,3): W[3][32], b[3]                     */

/* ---- sizes ------------------------------------------------------------------------------------------------ */
int qs_state_len(int variant);                /* 16 (E2E) or 13 (INDI): width of world_states                     */
int qs_obs_len(int variant, int gates_ahead); /* 20+4*ga (`:330`) or 13+4*ga (INDI `:185`): width of states       */
const char *qs_version(void);

/* ---- life cycle: Quadcopter3DGates.__init__ (`3D quad race.ipynb:288-360`) ------------------------------- */
/* gate_pos (n_gates,3) / gate_yaw (n_gates) / start_pos (3) are host float32 (the reference casts with
 * .astype(np.float32), `:298-300`).  The relative-gate tables of `:309-319` and cos/sin(gate_yaw) are computed
 * here in float32; qs_set_track_tables can override them with the caller's own (NumPy's) values.
 * stream: a cudaStream_t (NULL = default stream).  E2E handles start with the packaged residual weights unset:
 * call qs_set_residual_weights before the first step. */
int qs_create(qs_env **out, int variant, int64_t num_envs, int n_gates, const float *gate_pos, const float *gate_yaw,
              const float *start_pos, int gates_ahead, int device, void *stream);
int qs_destroy(qs_env *env);
const char *qs_last_error(const qs_env *env); /* env may be NULL: error of the last failed qs_create */

/* ---- configuration (the reference's attribute pokes, SURVEY.md section 5 "Config / flags") ---------------- */
int qs_set_stream(qs_env *env, void *stream);
int qs_set_track_tables(qs_env *env, const float *gate_cos, const float *gate_sin, const float *gate_pos_rel,
                        const float *gate_yaw_rel);                /* any pointer may be NULL = keep            */
int qs_get_track_tables(qs_env *env, float *gate_cos, float *gate_sin, float *gate_pos_rel, float *gate_yaw_rel);
int qs_set_max_steps(qs_env *env, int64_t max_steps);               /* env.max_steps (default 1200, `:345`)      */
int qs_set_dt(qs_env *env, float dt);                               /* env.dt (default 0.01f, `:346`)            */
/* env.disturbance_ranges (6,2) rows Mx,My,Mz,Fx,Fy,Fz and env.disturbance_scale (`:355-358`, `:482-489`);
 * ranges_are_f64: dtype of the array the caller assigned (the training cell assigns float64, `:772-781`) */
int qs_set_disturbance_ranges(qs_env *env, const double *ranges12, int ranges_are_f64, double scale);
/* thrust_model.pt / moment_model.pt parameters, row-major [out][in] as torch stores them (`:228-245`) */
int qs_set_residual_weights(qs_env *env, const float *thrust289, const float *moment451);
int qs_seed(qs_env *env, uint64_t seed);                            /* device RNG only; rewinds its launch epoch */
int qs_set_env_offset(qs_env *env, int64_t global_index_of_env0);   /* shard of a multi-GPU job                  */
/* Fused observation all-gather for a sharded job whose policy needs every rank's observations (BASELINE config 4):
 * peer_obs_bases[p] is the base of ANOTHER rank's (total_envs, obs_len) gather buffer, mapped into this process
 * (CUDA IPC / symmetric memory); every qs_step then stores its observation rows into obs_dev AND, tile by tile over
 * NVLink, into each peer buffer at row (row_offset + env).  n_peers = 0 switches it off.  The caller synchronises
 * the ranks (one barrier) before reading a gathered buffer. */
int qs_set_obs_peers(qs_env *env, int n_peers, void *const *peer_obs_bases, int64_t row_offset);
/* Observation format of the DEVICE entry points (qs_step / qs_reset_all / qs_observe) for a consumer that is the
 * on-device policy and nothing else (config 4: the gathered observations only feed `model.predict`, `:803`):
 * QS_OBS_BF16_K32 makes obs_dev (and the peer buffers) a packed buffer of qs_obs_packed_bytes(obs_len, n) bytes: one
 * block per 32 envs, [K chunk 0..chunks-1][row 0..31][8 x BF16] with chunks = ceil(obs_len / 8), i.e. the columns of the
 * first layer's A operand of qs_policy_forward_packed that carry observation values (a trailing chunk that would hold
 * only the constant 1 of the folded bias is synthesised by the policy kernel): 48 B per env over NVLink instead of
 * 96 (E2E, gates_ahead = 1) / 68 (INDI).  Needs obs_len <= 31; with peers, row_offset % 128 == 0.  The host-buffer
 * entry points (qs_step_host ...) and qs_rollout keep the reference's float32 rows and refuse a packed env. */
typedef enum { QS_OBS_F32 = 0, QS_OBS_BF16_K32 = 1 } qs_obs_format;
int qs_set_obs_format(qs_env *env, int format);
int64_t qs_obs_packed_bytes(int obs_len, int64_t num_envs);
int qs_enable_stats(qs_env *env, int on);
int qs_get_stats(qs_env *env, qs_stats *out, int reset_after_read); /* synchronises                              */

/* ---- state import / export in the reference's layout (host pointers, synchronise) ------------------------- */
/* world_states (N, state_len) f32, disturbances (N,6) f32 [E2E only], target_gates (N) i64, step_counts (N) i64.
 * Any pointer may be NULL.  first/count select a slice of envs. */
int qs_set_state(qs_env *env, int64_t first, int64_t count, const float *world_states, const float *disturbances,
                 const int64_t *target_gates, const int64_t *step_counts);
int qs_get_state(qs_env *env, int64_t first, int64_t count, float *world_states, float *disturbances,
                 int64_t *target_gates, int64_t *step_counts);

/* ---- the hot path (device pointers, asynchronous) ----------------------------------------------------------- */
/* update_states_gate (`:365-450`): obs_dev (N, obs_len) f32 row-major, 16-byte aligned */
int qs_observe(qs_env *env, float *obs_dev);
/* reset() with the device RNG (`:495-496`): redraw every env, zero counters, write the observation */
int qs_reset_all(qs_env *env, float *obs_dev);
/* step_async + step_wait (`:498-595`).
 *   actions_dev (N,4) f32 row-major, 16-byte aligned (the policy's output, clipped to [-1,1] by the caller)
 *   obs_dev     (N,obs_len) f32   -- MUST be a different buffer from the one returned by the previous call while
 *                                    the caller still reads it (the reference returns a fresh array every step,
 *                                    SURVEY.md section 8b "Ownership"); not written in QS_MODE_PAUSE
 *   rew_dev     (N) f32
 *   done_dev    (N) u8  0/1
 *   flags_dev   (N) u8  QS_F_* bits, or NULL */
int qs_step(qs_env *env, const float *actions_dev, float *obs_dev, float *rew_dev, uint8_t *done_dev,
            uint8_t *flags_dev, int mode, int reset_source);
/* second half of a QS_RESET_HOST step: the masked stores of reset_ (`:476-489`) for `count` envs.
 *   env_index (count) i32 host, ascending; world_states (count,state_len) f32 host; disturbances (count,6) f32
 *   host or NULL (already multiplied by disturbance_scale).  Zeroes step_counts/target_gates of those envs and
 *   rewrites their rows of obs_dev.  Synchronises (host inputs). */
int qs_apply_reset(qs_env *env, int64_t count, const int32_t *env_index, const float *world_states,
                   const float *disturbances, float *obs_dev);

/* ---- host-buffer convenience: what a NumPy-facing VecEnv calls --------------------------------------------- */
/* One full step with HOST buffers: copies actions in, steps, copies obs/rew/done(/flags) out, synchronises.
 * Buffers from qs_host_alloc are pinned and make the copies asynchronous DMA.  Above 32768 envs the call is pipelined
 * over chunks of envs (upload + kernel of chunk c+1 under the download of chunk c; same bits as one qs_step): four chunks,
 * the first a quarter of an equal share so that the download -- 101 of the 117 bytes per env -- starts early.
 * QS_HOST_CHUNKS / QS_HOST_FIRST_DIV in the environment override both (read at an env's first host-buffer call). */
int qs_step_host(qs_env *env, const float *actions_host, float *obs_host, float *rew_host, uint8_t *done_host,
                 uint8_t *flags_host, int mode, int reset_source);
/* The same step for the arrays a NumPy caller really has: `actions_host` is (num_envs,4) float32 or float64
 * (`step_async` keeps whatever array it is given, `3D quad race.ipynb:498-499`) in pageable or pinned memory, converted
 * chunk by chunk into a pinned staging buffer while the previous chunk is on the bus.  `info` (may be NULL; needs
 * flags_host) returns what the reference's per-env `infos` loop (`:589-594`) computes: the HIGHEST done env index
 * (whose observation row ends up as the one aliased dict's "terminal_observation"; -1 = none) and whether ANY env hit
 * max_steps ("TimeLimit.truncated").  The caller scans nothing. */
typedef enum { QS_F32 = 0, QS_F64 = 1 } qs_dtype;
typedef struct {
    int64_t last_done_index; /* -1 when no env is done (always -1 in QS_MODE_PAUSE, whose dones are cleared `:570-572`) */
    int64_t n_done;
    int32_t any_truncated;
    int32_t reserved;
} qs_step_info;
int qs_step_host_ex(qs_env *env, const void *actions_host, int actions_dtype, float *obs_host, float *rew_host,
                    uint8_t *done_host, uint8_t *flags_host, int mode, int reset_source, qs_step_info *info);
int qs_reset_all_host(qs_env *env, float *obs_host);
int qs_observe_host(qs_env *env, float *obs_host);
void *qs_host_alloc(size_t bytes);
void qs_host_free(void *p);

/* ---- introspection ------------------------------------------------------------------------------------------ */
/* algorithmic HBM bytes one env-step moves: 189+4*(20+4*ga) E2E, 141+4*(13+4*ga) INDI */
int qs_algorithmic_bytes_per_env_step(int variant, int gates_ahead);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
uint64_t qs_launch_count(const qs_env *env);
/* how many of those were CHAINED step launches that skipped the grid-wide wait: consecutive full-range qs_step calls
 * captured into one CUDA graph depend on each other CTA by CTA (CTA b steps the same tiles in every launch), so the next
 * step's loads flow while this step's last tiles drain.  QS_CHAIN=0 in the environment turns it off.  Results are the
 * same bit for bit (tests/test_gpu_chain.py).  No reference counterpart: the reference steps synchronously on the host. */
uint64_t qs_chained_launch_count(const qs_env *env);
/* geometry of the device-resident state, for zero-copy consumers.  Env i lives in block i/32 at lane i%32; block b
 * starts at base + b*block_bytes; inside a block every field is a 32-lane plane at offsets7[k]:
 *   0: (x,y,z,vx) float4   1: (vy,vz,phi,theta) float4   2: (psi,p,q,r) float4   3: (w1..w4) float4 | INDI: T_norm float
 *   4: packed counters u32 (target_gate<<24 | step_count)   5: disturbances (Mx,My,Mz,Fz) float4   6: (Fx,Fy) float2
 * (5, 6 are -1 for INDI).  Any out pointer may be NULL. */
int qs_get_state_layout(qs_env *env, void **base, int *block_bytes, int *offsets7);

/* ---- on-device policy ----------------------------------------------------------- */
/* The trained controller the reference evaluates through SB3 (`model.predict(env.states)`, `3D quad race.ipynb:803`)
 * or through its generated C (`c_code/neural_network.c:397-430`, noise + clip `c_code/nn_controller.c:158-176`):
 *   mean = W_out . relu(W_h ... relu(W_1 . obs + b_1) ...) + b_out ;  action = clip(mean + std * N(0,1), -1, 1).
 * Evaluated on the tensor cores (tcgen05, BF16 operands, FP32 accumulation in TMEM), observations and actions stay
 * on the device.  in_dim 1..63, 1..4 hidden layers of width <= 127, out_dim <= 4. */
typedef struct qs_policy qs_policy; /* opaque */
int qs_policy_create(qs_policy **out, int in_dim, int n_hidden, int hidden_dim, int out_dim, int device, void *stream);
int qs_policy_destroy(qs_policy *policy);
/* hidden activation: SB3's `activation_fn` (`3D quad race.ipynb:784`: torch.nn.ReLU; SB3's default is Tanh); the
 * reference's generated C has both (`nn_relu` / `nn_tanh`, c_code/neural_network.c:407-417).  tanh needs hidden_dim <= 120. */
typedef enum { QS_ACT_RELU = 0, QS_ACT_TANH = 1 } qs_activation;
int qs_policy_set_activation(qs_policy *policy, int activation);
/* > 0: qs_policy_forward sanitises its float32 observations like SB3-side learners do before a forward over the rollout
 * buffer (NaN -> 0, clamp to +-limit); used by the PPO loop's value / old-log-prob pass.  0 (default) = off. */
int qs_policy_set_obs_limit(qs_policy *policy, float limit);
const char *qs_policy_last_error(const qs_policy *policy); /* NULL: error of the last failed qs_policy_create */
int qs_policy_set_stream(qs_policy *policy, void *stream);
/* layer 0..n_hidden (the last is the output layer); W row-major [out][in] and b [out] as torch / the generated C
 * store them (`c_code/neural_network.c:5-395`), host float32 */
int qs_policy_set_layer(qs_policy *policy, int layer, const float *W, const float *b);
int qs_policy_set_std(qs_policy *policy, const float *std);          /* exp(log_std), `c_code/nn_controller.c:7-12` */
int qs_policy_seed(qs_policy *policy, uint64_t seed);                /* exploration noise: Philox keyed by (seed, global env, launch) */
int qs_policy_set_env_offset(qs_policy *policy, int64_t global_index_of_env0);
/* actions_dev (n,4) f32 <- policy(obs_dev (n,in_dim) f32); mean_dev (n,4) f32 receives the pre-noise output and
 * raw_dev (n,4) f32 the sampled action before the clip (the value PPO takes the log-probability of) when not NULL;
 * deterministic != 0 skips the noise (`nn_controller.c:5`).  Asynchronous on the policy's stream. */
int qs_policy_forward(qs_policy *policy, const float *obs_dev, int64_t n, float *actions_dev, float *mean_dev,
                      float *raw_dev, int deterministic);
/* the same forward over a QS_OBS_BF16_K32 buffer (n envs, qs_obs_packed_bytes(in_dim, n) bytes, 16-byte aligned): the packed
 * blocks are TMA-loaded into the first MMA's operand buffer as they are.  Same actions as qs_policy_forward on the
 * float32 rows they were packed from, bit for bit.  Needs in_dim <= 31. */
int qs_policy_forward_packed(qs_policy *policy, const void *packed_obs_dev, int64_t n, float *actions_dev, float *mean_dev,
                             float *raw_dev, int deterministic);
uint64_t qs_policy_launch_count(const qs_policy *policy);
/* collect_rollouts without the host (SB3's loop behind `model.learn`, `3D quad race.ipynb:820`): for t < steps
 *   act_buf[t] = policy(obs_buf[t]);  obs_buf[t+1], rew_buf[t], done_buf[t] = step(act_buf[t])   (fused device reset)
 * obs_buf (steps+1, N, obs_len) f32 with obs_buf[0] = current observations, act_buf (steps, N, 4) f32 (clipped, what
 * the env received), raw_buf (steps, N, 4) f32 or NULL (un-clipped samples, what SB3 keeps in its RolloutBuffer),
 * rew_buf (steps, N) f32, done_buf (steps, N) u8, flags_buf (steps, N) u8 or NULL (the QS_F_* bits of every step:
 * QS_F_TRUNCATED is what the reference reports as infos["TimeLimit.truncated"], `3D quad race.ipynb:592-593`, which SB3's
 * time-limit bootstrap needs) -- all device pointers.  Asynchronous on the env's stream. */
int qs_rollout(qs_env *env, qs_policy *policy, int steps, float *obs_buf, float *act_buf, float *raw_buf, float *rew_buf,
               uint8_t *done_buf, uint8_t *flags_buf, int deterministic);
/* The same rollout as ONE launch of the fused closed-loop kernel: each 128-env tile stays in registers for all
 * `steps`, the controller runs on the tensor cores in between, only the rollout buffers are written to HBM (the
 * simulator state is read and written once per launch instead of once per step).  Same arguments and -- because reset
 * draws and exploration noise are keyed by (seed, env, launch epoch + t) -- the same results, bit for bit, as
 * qs_rollout.  qs_rollout_fused_supported() != 0 when the env / policy shapes fit the kernel (observation tile +
 * layer-1 operand <= 32 KB per tile group, 4 actions, same device). */
int qs_rollout_fused_supported(const qs_env *env, const qs_policy *policy);
int qs_rollout_fused(qs_env *env, qs_policy *policy, int steps, float *obs_buf, float *act_buf, float *raw_buf,
                     float *rew_buf, uint8_t *done_buf, uint8_t *flags_buf, int deterministic);
/* SB3 `RolloutBuffer.compute_returns_and_advantage` on the device: rew/adv/ret (steps, n) f32, val (steps+1, n) f32
 * (last row = bootstrap values), done (steps, n) u8.  A_t = delta_t + gamma*lambda*(1-done_t)*A_{t+1}. */
int qs_gae(const float *rew_dev, const float *val_dev, const uint8_t *done_dev, float *adv_dev, float *ret_dev, int64_t n,
           int steps, float gamma, float lambda, void *stream);

/* ---- PPO update on the tensor cores (row f2) ------------------------------------------------------------------
 * What SB3's `PPO.train()` does for the reference's configuration (`3D quad race.ipynb:784-795`: MlpPolicy with separate
 * pi / vf networks obs -> h -> h -> h -> {4 | 1}, ReLU, free log_std; clipped surrogate + vf_coef * value MSE - ent_coef *
 * entropy; per-minibatch advantage normalisation; clip_grad_norm_; Adam eps 1e-5) as hand-written sm_100a kernels: one
 * fused forward + loss + backward kernel for both networks (tcgen05, weight gradients accumulated in TMEM), a reduction
 * and an Adam kernel per minibatch (csrc/quadsim_train.cuh).  Float32 master parameters and Adam moments live in the
 * handle; layers are exchanged in torch layout: W (out, in) row-major, b (out); layer 0..2 hidden, 3 = output. */
typedef struct qs_trainer qs_trainer;
typedef struct {
    float learning_rate, beta1, beta2, eps;      /* Adam (SB3: 3e-4, 0.9, 0.999, 1e-5) */
    float clip_range, vf_coef, ent_coef, max_grad_norm;
    float obs_limit, act_limit;                  /* sanitising clamps of the learner's inputs (see ppo.py) */
    int32_t normalize_advantage;
    int32_t reserved;
} qs_train_hyper;
int qs_trainer_create(qs_trainer **out, int in_dim, int hidden_dim, int device, void *stream);
int qs_trainer_destroy(qs_trainer *t);
const char *qs_trainer_last_error(const qs_trainer *t);
int qs_trainer_set_stream(qs_trainer *t, void *stream);
int qs_trainer_set_layer(qs_trainer *t, int net /* 0 policy, 1 value */, int layer, const float *W, const float *b);
int qs_trainer_get_layer(qs_trainer *t, int net, int layer, float *W, float *b);
int qs_trainer_set_log_std(qs_trainer *t, const float *log_std4);
int qs_trainer_get_log_std(qs_trainer *t, float *log_std4);
int qs_trainer_reset_optimizer(qs_trainer *t);
/* one minibatch of `rows` samples (idx_dev: int64 indices into the flat rollout buffers, or NULL for 0..rows-1); all
 * pointers are device pointers; apply != 0 also runs the optimizer step.  Asynchronous on the trainer's stream. */
int qs_trainer_minibatch(qs_trainer *t, const int64_t *idx_dev, int64_t rows, const float *obs_dev, const float *raw_act_dev,
                         const float *old_logp_dev, const float *adv_dev, const float *ret_dev, const float *weight_dev,
                         const qs_train_hyper *hyper, int apply);
int qs_trainer_get_grad(qs_trainer *t, int net, int layer, float *W, float *b, float *log_std4);
int qs_trainer_get_stats(qs_trainer *t, float *out8, int reset);
/* hand the updated policy network (BF16 blob + std) to the device actor, device to device */
int qs_trainer_publish(qs_trainer *t, qs_policy *policy);
uint64_t qs_trainer_launch_count(const qs_trainer *t);

#ifdef __cplusplus
}
#endif
#endif /* QUADSIM_H */
