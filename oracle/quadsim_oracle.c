/* TEST INFRASTRUCTURE -- CPU restatement (plain C99, float32) of the reference's quadrotor racing step.
 *
 * NOT part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the CUDA path never does.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks every function here against vectors produced by
 * the unmodified reference cells (oracle/make_golden.py -> tests/golden/ *.npz), including the reference's own
 * known-answer vector K1 (`3D quad race.ipynb:215,219`) and gate tables K3 (`c_code/nn_controller.c:40-60`).
 *
 * What is restated (raw .ipynb JSON line numbers):
 *   qo_track_tables      Quadcopter3DGates.__init__ relative-gate precompute     3D quad race.ipynb:309-319
 *   qo_body_velocity     get_body_velocity                                       3D quad race.ipynb:155
 *   qo_residual_mlp      thrust_moment_model_world_states + the two nn.Sequential 3D quad race.ipynb:244-262
 *   f_e2e                f_func (lambdified sympy expression, same term order)   3D quad race.ipynb:65-152
 *   f_indi               f_func of the INDI notebook                             ...INDI inner loop.ipynb:48-110
 *   qo_step              step_wait up to (not including) the RNG draws           3D quad race.ipynb:501-585
 *   qo_observe           update_states_gate                                      3D quad race.ipynb:365-450
 *
 * Arithmetic: every operation is a separately rounded binary32 operation in the order the reference's NumPy
 * expression evaluates it (compile with -ffp-contract=off).  Float64 literals are rounded to binary32 first,
 * which is what NumPy's weak-scalar promotion does.  sinf/cosf/tanf come from libm instead of NumPy's SIMD
 * loops; that is the one unavoidable last-bit difference.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QO_E2E 0
#define QO_INDI 1

/* step_wait branch selector (`3D quad race.ipynb:568-585`) */
#define QO_MODE_NORMAL 0             /* advance all, reset the done ones                 */
#define QO_MODE_PAUSE_IF_COLLISION 1 /* advance only the not-done ones, never reset      */
#define QO_MODE_PAUSE 2              /* advance nothing, report no dones                 */

/* bits of the per-env flag byte */
#define QO_F_DONE 1
#define QO_F_TRUNC 2 /* step_counts >= max_steps */
#define QO_F_GATE_PASSED 4
#define QO_F_GATE_COLLISION 8
#define QO_F_GROUND 16
#define QO_F_OOB 32

typedef struct {
    int variant;     /* QO_E2E / QO_INDI */
    int n_gates;
    int gates_ahead;
    int ranges_f64;  /* 1: disturbance_ranges is a float64 array (the training notebook assigns one) */
    int64_t max_steps;
    float dt;
    const float *gate_pos;     /* (n_gates,3) */
    const float *gate_yaw;     /* (n_gates)   */
    const float *gate_cos;     /* (n_gates) cos/sin of gate_yaw as the caller's NumPy computed them, or NULL */
    const float *gate_sin;
    const float *gate_pos_rel; /* (n_gates,3) */
    const float *gate_yaw_rel; /* (n_gates)   */
    const double *dist_ranges; /* (6,2) */
    const float *thrust_w;     /* W1(32x7) b1(32) W2(1x32) b2(1)   = 289 */
    const float *moment_w;     /* W1(32x10) b1(32) W2(3x32) b2(3)  = 451 */
} qo_params;

int qo_state_len(int variant) { return variant == QO_E2E ? 16 : 13; }
int qo_obs_len(int variant, int gates_ahead) {
    return variant == QO_E2E ? 20 + 4 * gates_ahead : 13 + 4 * gates_ahead;
}

/* bench.py's CPU legs: use every host thread even when the launcher exported OMP_NUM_THREADS=1 (torchrun does) */
void qo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int qo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------- __init__ */
void qo_track_tables(int ng, const float *gate_pos, const float *gate_yaw, float *pos_rel, float *yaw_rel) {
    for (int i = 0; i < ng; ++i) {
        int j = (i + ng - 1) % ng; /* python's gate[i-1] */
        float dx = gate_pos[3 * i + 0] - gate_pos[3 * j + 0];
        float dy = gate_pos[3 * i + 1] - gate_pos[3 * j + 1];
        float dz = gate_pos[3 * i + 2] - gate_pos[3 * j + 2];
        float c = cosf(gate_yaw[j]), s = sinf(gate_yaw[j]);
        pos_rel[3 * i + 0] = c * dx + s * dy;
        pos_rel[3 * i + 1] = (-s) * dx + c * dy;
        pos_rel[3 * i + 2] = dz;
        yaw_rel[i] = gate_yaw[i] - gate_yaw[j];
    }
}

/* ---------------------------------------------------------------------------------------------- L0 model */
typedef struct { float sph, cph, sth, cth, sps, cps; } trig_t;

static inline trig_t trig_of(const float *s) {
    trig_t t;
    t.sph = sinf(s[6]); t.cph = cosf(s[6]);
    t.sth = sinf(s[7]); t.cth = cosf(s[7]);
    t.sps = sinf(s[8]); t.cps = cosf(s[8]);
    return t;
}

/* v_b = R^T v, term order of the lambdified get_body_velocity */
static inline void body_velocity(const float *s, const trig_t *t, float *vb) {
    float vx = s[3], vy = s[4], vz = s[5];
    vb[0] = vx * t->cps * t->cth + vy * t->sps * t->cth - vz * t->sth;
    vb[1] = vx * (t->sph * t->sth * t->cps - t->sps * t->cph) + vy * (t->sph * t->sps * t->sth + t->cph * t->cps) +
            vz * t->sph * t->cth;
    vb[2] = vx * (t->sph * t->sps + t->sth * t->cph * t->cps) + vy * (-t->sph * t->cps + t->sps * t->sth * t->cph) +
            vz * t->cph * t->cth;
}

void qo_body_velocity(const float *ws, int64_t n, float *vb) {
    for (int64_t i = 0; i < n; ++i) {
        trig_t t = trig_of(ws + 16 * i);
        body_velocity(ws + 16 * i, &t, vb + 3 * i);
    }
}

/* Linear(k,32)-ReLU-Linear(32,m); bias first then sequential accumulate, like c_code/nn_thrust.c:52-60 */
static inline void mlp32(const float *w, int k, int m, const float *x, float *y) {
    const float *w1 = w, *b1 = w + 32 * k, *w2 = b1 + 32, *b2 = w2 + 32 * m;
    float h[32];
    for (int j = 0; j < 32; ++j) {
        float a = b1[j];
        for (int i = 0; i < k; ++i) a += x[i] * w1[j * k + i];
        h[j] = a > 0.0f ? a : 0.0f;
    }
    for (int o = 0; o < m; ++o) {
        float a = b2[o];
        for (int j = 0; j < 32; ++j) a += h[j] * w2[o * 32 + j];
        y[o] = a;
    }
}

static inline void residual(const qo_params *p, const float *s, const trig_t *t, float *thrust, float *moment) {
    float x[10];
    x[0] = s[12]; x[1] = s[13]; x[2] = s[14]; x[3] = s[15];
    body_velocity(s, t, x + 4);
    x[7] = s[9]; x[8] = s[10]; x[9] = s[11];
    mlp32(p->thrust_w, 7, 1, x, thrust);
    mlp32(p->moment_w, 10, 3, x, moment);
}

void qo_residual_mlp(const qo_params *p, const float *ws, int64_t n, float *thrust, float *moment) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        trig_t t = trig_of(ws + 16 * i);
        residual(p, ws + 16 * i, &t, thrust + i, moment + 3 * i);
    }
}

/* Python-literal constants as binary32 (NumPy weak scalars) */
#define KX 1.07933887e-5f
#define KY 9.65250793e-6f
#define KZ 2.7862899e-5f
#define KW 4.36301076e-8f
#define KH 0.0625501332f

/* f_func (E2E): `inspect.getsource(f_func)` of cell 2, transcribed term by term.  d = [M_ext xyz, F_ext xyz]. */
static inline void f_e2e(const float *s, const float *u, const float *d, const trig_t *t, float *f) {
    const float vx = s[3], vy = s[4], vz = s[5], p = s[9], q = s[10], r = s[11];
    const float w1 = s[12], w2 = s[13], w3 = s[14], w4 = s[15];
    const float sph = t->sph, cph = t->cph, sth = t->sth, cth = t->cth, sps = t->sps, cps = t->cps;
    const float Mx = d[0], My = d[1], Mz = d[2], Fx = d[3], Fy = d[4], Fz = d[5];

    const float sumW = 4000 * w1 + 4000 * w2 + 4000 * w3 + 4000 * w4 + 28000;
    const float W1 = 4000 * w1 + 7000, W2 = 4000 * w2 + 7000, W3 = 4000 * w3 + 7000, W4 = 4000 * w4 + 7000;
    /* rotation-matrix entries as the expression spells them */
    const float r01 = sph * sth * cps - sps * cph; /* R[0][1] */
    const float r11 = sph * sps * sth + cph * cps; /* R[1][1] */
    const float r02 = sph * sps + sth * cph * cps; /* R[0][2] */
    const float r12 = -sph * cps + sps * sth * cph; /* R[1][2] */

    const float Dx = Fx + (-KX * vx * cps * cth - KX * vy * sps * cth + KX * vz * sth) * sumW;
    const float Dy = Fy + (-KY * vx * r01 - KY * vy * r11 - KY * vz * sph * cth) * sumW;
    const float vby = vx * r01 + vy * r11 + vz * sph * cth;
    const float vbx = vx * cps * cth + vy * sps * cth - vz * sth;
    const float T = Fz - KW * (W1 * W1) - KW * (W2 * W2) - KW * (W3 * W3) - KW * (W4 * W4) -
                    (KZ * vx * r02 + KZ * vy * r12 + KZ * vz * cph * cth) * sumW - KH * (vby * vby) - KH * (vbx * vbx);

    f[0] = vx; f[1] = vy; f[2] = vz;
    f[3] = Dx * cps * cth + Dy * r01 + r02 * T;
    f[4] = Dx * sps * cth + Dy * r11 + r12 * T;
    f[5] = -Dx * sth + Dy * sph * cth + T * cph * cth + 9.81f;
    {
        const float tth = tanf(s[7]);
        f[6] = p + q * sph * tth + r * cph * tth;
        f[7] = q * cph - r * sph;
        f[8] = q * sph / cth + r * cph / cth;
    }
    f[9] = 1103.7527593819f * Mx - 0.896247240618101f * q * r - 8.79803364238411f * vx * r01 -
           8.79803364238411f * vy * r11 - 8.79803364238411f * vz * sph * cth + 1.55842505518764e-6f * (W1 * W1) -
           1.55842505518764e-6f * (W2 * W2) - 1.55842505518764e-6f * (W3 * W3) + 1.55842505518764e-6f * (W4 * W4);
    f[10] = 805.152979066023f * My + 0.924315619967794f * p * r + 10.4077084541063f * vx * cps * cth +
            10.4077084541063f * vy * sps * cth - 10.4077084541063f * vz * sth + 9.79081191626409e-7f * (W1 * W1) +
            9.79081191626409e-7f * (W2 * W2) - 9.79081191626409e-7f * (W3 * W3) - 9.79081191626409e-7f * (W4 * W4);
    f[11] = 486.854917234664f * Mz - 0.163583252190847f * p * q - 0.395780237098345f * r - 13.3373373580007f * u[0] +
            13.3373373580007f * u[1] - 13.3373373580007f * u[2] + 13.3373373580007f * u[3] + 8.33177659850698f * w1 -
            8.33177659850698f * w2 + 8.33177659850698f * w3 - 8.33177659850698f * w4;
    f[12] = 16.6666666666667f * u[0] - 16.6666666666667f * w1;
    f[13] = 16.6666666666667f * u[1] - 16.6666666666667f * w2;
    f[14] = 16.6666666666667f * u[2] - 16.6666666666667f * w3;
    f[15] = 16.6666666666667f * u[3] - 16.6666666666667f * w4;
}

/* f_func (INDI): lambdified source of `...INDI inner loop.ipynb` cell 2 */
static inline void f_indi(const float *s, const float *u, const trig_t *t, float *f) {
    const float vx = s[3], vy = s[4], vz = s[5], p = s[9], q = s[10], r = s[11], Tn = s[12];
    const float sph = t->sph, cph = t->cph, sth = t->sth, cth = t->cth, sps = t->sps, cps = t->cps;
    const float r01 = sph * sth * cps - sps * cph, r11 = sph * sps * sth + cph * cps;
    const float r02 = sph * sps + sth * cph * cps, r12 = -sph * cps + sps * sth * cph;
    const float KXI = 0.33915248f, KYI = 0.4314916f;
    const float Dx = -KXI * vx * cps * cth - KXI * vy * sps * cth + KXI * vz * sth;
    const float Dy = -KYI * vx * r01 - KYI * vy * r11 - KYI * vz * sph * cth;
    const float mT = -8.0f * Tn - 8.0f;
    f[0] = vx; f[1] = vy; f[2] = vz;
    f[3] = mT * r02 + r01 * Dy + Dx * cps * cth;
    f[4] = mT * r12 + r11 * Dy + Dx * sps * cth;
    f[5] = mT * cph * cth + Dy * sph * cth - Dx * sth + 9.81f;
    {
        const float tth = tanf(s[7]);
        f[6] = p + q * sph * tth + r * cph * tth;
        f[7] = q * cph - r * sph;
        f[8] = q * sph / cth + r * cph / cth;
    }
    f[9] = -33.3333333333333f * p + 100.0f * u[0];
    f[10] = -33.3333333333333f * q + 100.0f * u[1];
    f[11] = -33.3333333333333f * r + 66.6666666666667f * u[2];
    f[12] = 33.3333333333333f * u[3] - 33.3333333333333f * Tn;
}

/* new_states = world_states + dt * f(...)  -- one forward-Euler step (`:503-512`; INDI `:304`) for one env. */
static inline void euler_one(const qo_params *p, const float *s, const float *act, const float *dist, float *out,
                             float *thrust_out, float *moment_out) {
    const int ns = qo_state_len(p->variant);
    float f[16];
    trig_t t = trig_of(s);
    if (p->variant == QO_E2E) {
        float d[6] = {0, 0, 0, 0, 0, 0}, th, mo[3];
        residual(p, s, &t, &th, mo);
        d[0] = mo[0]; d[1] = mo[1]; d[2] = mo[2]; d[5] = th;
        for (int k = 0; k < 6; ++k) d[k] += dist[k];
        if (thrust_out) *thrust_out = th;
        if (moment_out) memcpy(moment_out, mo, sizeof mo);
        f_e2e(s, act, d, &t, f);
    } else {
        f_indi(s, act, &t, f);
    }
    for (int k = 0; k < ns; ++k) out[k] = s[k] + p->dt * f[k];
}

/* Batched form; thrust/moment (optional outputs, E2E only) expose the residual MLP values. */
void qo_euler(const qo_params *p, const float *ws, const float *act, const float *dist, int64_t n, float *out,
              float *thrust_out, float *moment_out) {
    const int ns = qo_state_len(p->variant);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        euler_one(p, ws + ns * i, act + 4 * i, dist ? dist + 6 * i : NULL, out + ns * i,
                  thrust_out ? thrust_out + i : NULL, moment_out ? moment_out + 3 * i : NULL);
}

/* ---------------------------------------------------------------------------------------------- observation */
/* np.remainder for binary32 (python-style sign), then the two wrap branches of `:393-396` */
static inline float wrap_yaw(float yaw) {
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float m = fmodf(yaw, two_pi);
    if (m != 0.0f) { if (m < 0.0f) m += two_pi; } else { m = copysignf(0.0f, two_pi); }
    if (m > pi) m -= two_pi;
    if (m < -pi) m += two_pi;
    return m;
}

static inline void gate_cs(const qo_params *p, int g, float *c, float *s) {
    if (p->gate_cos) { *c = p->gate_cos[g]; *s = p->gate_sin[g]; }
    else { *c = cosf(p->gate_yaw[g]); *s = sinf(p->gate_yaw[g]); }
}

static inline void observe_one(const qo_params *p, const float *s, const float *dist, int64_t tg, float *o) {
    const int ng = p->n_gates, ns = qo_state_len(p->variant);
    const int g = (int)(tg % ng);
    const float *gp = p->gate_pos + 3 * g;
    float c, sn;
    gate_cs(p, g, &c, &sn);
    const float dx = s[0] - gp[0], dy = s[1] - gp[1];
    /* (1x2)@(2x2) with R = [[c,-s],[s,c]] after the transpose((2,1,0)) of `:371-374` */
    o[0] = dx * c + dy * sn;
    o[1] = dx * (-sn) + dy * c;
    o[2] = s[2] - gp[2];
    o[3] = s[3] * c + s[4] * sn;
    o[4] = s[3] * (-sn) + s[4] * c;
    o[5] = s[5];
    o[6] = s[6]; o[7] = s[7];
    o[8] = wrap_yaw(s[8] - p->gate_yaw[g]);
    for (int k = 9; k < ns; ++k) o[k] = s[k];
    for (int i = 0; i < p->gates_ahead; ++i) {
        const int idx = (int)((tg + i + 1) % ng);
        o[ns + 4 * i + 0] = p->gate_pos_rel[3 * idx + 0];
        o[ns + 4 * i + 1] = p->gate_pos_rel[3 * idx + 1];
        o[ns + 4 * i + 2] = p->gate_pos_rel[3 * idx + 2];
        o[ns + 4 * i + 3] = p->gate_yaw_rel[idx];
    }
    if (p->variant == QO_E2E) {
        static const int rows[4] = {0, 1, 2, 5};
        float *tail = o + 16 + 4 * p->gates_ahead;
        for (int k = 0; k < 4; ++k) {
            double lo = p->dist_ranges[2 * rows[k]], hi = p->dist_ranges[2 * rows[k] + 1];
            if (lo == hi) { lo -= 1; hi += 1; }
            if (p->ranges_f64) { /* float32 array (op) float64 scalar -> float64, stored back as float32 */
                tail[k] = (float)(2 * ((double)dist[rows[k]] - lo) / (hi - lo) - 1);
            } else {
                const float lof = (float)lo, hif = (float)hi;
                tail[k] = 2 * (dist[rows[k]] - lof) / (hif - lof) - 1;
            }
        }
    }
}

void qo_observe(const qo_params *p, const float *ws, const float *dist, const int64_t *tg, int64_t n, float *obs) {
    const int ns = qo_state_len(p->variant), nd = qo_obs_len(p->variant, p->gates_ahead);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        observe_one(p, ws + ns * i, dist ? dist + 6 * i : NULL, tg[i], obs + (size_t)nd * i);
}

/* ---------------------------------------------------------------------------------------------- step_wait */
static inline float norm3(float a, float b, float c) { return sqrtf(a * a + b * b + c * c); }

/* One step_wait for n envs, in place on (ws, tg, sc), everything except drawing random numbers.
 *   reset_ws/reset_dist : values for the done envs, row k belongs to the k-th done env in index order --
 *                         what `self.world_states[dones] = np.stack(...)` consumes (`:476`, `:489`).  Only read in
 *                         QO_MODE_NORMAL; may be NULL when the caller knows nothing terminates.
 *   new_raw (optional)  : the un-reset Euler result for every env.
 *   obs                 : observation after the branch logic (not written in QO_MODE_PAUSE, as in the reference,
 *                         which returns the previous self.states untouched).
 * Returns the number of done envs (before the pause override).
 */
int64_t qo_step(const qo_params *p, int mode, int64_t n, float *ws, float *dist, int64_t *tg, int64_t *sc,
                const float *act, const float *reset_ws, const float *reset_dist, float *obs, float *rew,
                uint8_t *done, uint8_t *flags, float *new_raw) {
    const int ns = qo_state_len(p->variant), nd = qo_obs_len(p->variant, p->gates_ahead), ng = p->n_gates;
    int64_t n_done = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_done)
    for (int64_t i = 0; i < n; ++i) {
        float *s = ws + ns * i;
        float nw[16];
        euler_one(p, s, act + 4 * i, dist ? dist + 6 * i : NULL, nw, NULL, NULL);
        if (new_raw) memcpy(new_raw + ns * i, nw, ns * sizeof(float));
        sc[i] += 1;

        const int g = (int)(tg[i] % ng);
        const float *gp = p->gate_pos + 3 * g;
        float c, sn;
        gate_cs(p, g, &c, &sn);
        const float ox = s[0] - gp[0], oy = s[1] - gp[1], oz = s[2] - gp[2];
        const float nx = nw[0] - gp[0], ny = nw[1] - gp[1], nz = nw[2] - gp[2];
        const float d_old = norm3(ox, oy, oz), d_new = norm3(nx, ny, nz);
        float r = d_old - d_new - 0.0f * norm3(nw[9], nw[10], nw[11]); /* rat_penalty = 0*0.01*|omega| */

        const float proj_old = ox * c + oy * sn, proj_new = nx * c + ny * sn;
        const int plane = (proj_old < 0) && (proj_new > 0);
        const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
        const int passed = plane && (ax < 0.5f && ay < 0.5f && az < 0.5f);
        const int collided = plane && (ax > 0.5f || ay > 0.5f || az > 0.5f);
        if (passed) r = 10 - 10 * d_new;
        if (collided) r = -10;
        const int ground = nw[2] > 0;
        if (ground) r = -10;
        const int oob = (fabsf(nw[0]) > 10 || fabsf(nw[1]) > 10) ||
                        (fabsf(nw[9]) > 1000 || fabsf(nw[10]) > 1000 || fabsf(nw[11]) > 1000);
        if (oob) r = -10;
        const int trunc = sc[i] >= p->max_steps;
        if (passed) tg[i] = (tg[i] + 1) % ng;
        const int dn = trunc || ground || collided || oob;
        rew[i] = r;
        if (flags)
            flags[i] = (uint8_t)((dn ? QO_F_DONE : 0) | (trunc ? QO_F_TRUNC : 0) | (passed ? QO_F_GATE_PASSED : 0) |
                                 (collided ? QO_F_GATE_COLLISION : 0) | (ground ? QO_F_GROUND : 0) | (oob ? QO_F_OOB : 0));
        n_done += dn;
        if (mode == QO_MODE_PAUSE) {
            done[i] = 0;
        } else {
            done[i] = (uint8_t)dn;
            if (mode == QO_MODE_NORMAL || !dn) memcpy(s, nw, ns * sizeof(float));
        }
    }
    if (mode == QO_MODE_PAUSE) return n_done;
    if (mode == QO_MODE_NORMAL) { /* reset_(dones): masked stores in env-index order */
        int64_t k = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (!done[i]) continue;
            if (reset_ws) memcpy(ws + ns * i, reset_ws + ns * k, ns * sizeof(float));
            if (reset_dist && dist) memcpy(dist + 6 * i, reset_dist + 6 * k, 6 * sizeof(float));
            sc[i] = 0;
            tg[i] = 0;
            ++k;
        }
    }
    if (obs) qo_observe(p, ws, dist, tg, n, obs);
    (void)nd;
    return n_done;
}

/* The masked stores of reset_ on their own (`:476-489`): lets the host learn `done`, draw from np.random in the
 * reference's order, and then finish the step with qo_apply_reset + qo_observe. */
void qo_apply_reset(const qo_params *p, int64_t n, float *ws, float *dist, int64_t *tg, int64_t *sc,
                    const uint8_t *done, const float *reset_ws, const float *reset_dist) {
    const int ns = qo_state_len(p->variant);
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!done[i]) continue;
        memcpy(ws + ns * i, reset_ws + ns * k, ns * sizeof(float));
        if (reset_dist && dist) memcpy(dist + 6 * i, reset_dist + 6 * k, 6 * sizeof(float));
        sc[i] = 0;
        tg[i] = 0;
        ++k;
    }
}

/* ------------------------------------------------------------------------------------------------ policy forward
 * TEST INFRASTRUCTURE.  Restates the reference's generated controller network, `c_code/neural_network.c`:
 *   nn_linear (`:397-405`): neuron = bias; for j: neuron += input[j] * weights[i*in + j]   (float32, that order)
 *   nn_relu   (`:407-411`), nn_forward (`:419-430`): Linear-ReLU x n_hidden, then Linear.
 * W[l] is row-major [out][in].  dims = {in, hidden, ..., hidden, out} (n_layers + 1 entries).
 * Pinned by tests/test_policy_oracle.py against oracle/_ref/libnn_policy_ref.so (the reference's own C, compiled
 * where it lies) through tests/golden/policy_k4.npz. */
static int qo_policy_activation = 0; /* 0 = ReLU (`nn_relu`), 1 = tanh (`nn_tanh`, c_code/neural_network.c:413-417) */
void qo_set_policy_activation(int act) { qo_policy_activation = act; }
static inline float qo_act(float a) { return qo_policy_activation ? tanhf(a) : fmaxf(0.0f, a); }

void qo_policy_forward(int n_layers, const int *dims, const float *const *W, const float *const *b, const float *obs,
                       int64_t n, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float cur[256], nxt[256];
        for (int k = 0; k < dims[0]; ++k) cur[k] = obs[i * dims[0] + k];
        for (int l = 0; l < n_layers; ++l) {
            const int in = dims[l], on = dims[l + 1];
            for (int o = 0; o < on; ++o) {
                float acc = b[l][o];
                for (int k = 0; k < in; ++k) acc += cur[k] * W[l][(size_t)o * in + k];
                nxt[o] = (l + 1 < n_layers) ? qo_act(acc) : acc;
            }
            for (int o = 0; o < on; ++o) cur[o] = nxt[o];
        }
        for (int o = 0; o < dims[n_layers]; ++o) out[i * dims[n_layers] + o] = cur[o];
    }
}

static float qo_bf16(float f) { /* round-to-nearest-even to bfloat16, returned as float */
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u += 0x7FFFu + ((u >> 16) & 1u);
    u &= 0xFFFF0000u;
    memcpy(&f, &u, 4);
    return f;
}

/* The same network with the arithmetic of the B200 tensor-core path (quadsim_policy.cuh): inputs, weights, biases and
 * hidden activations rounded to BF16, products exact, accumulation in (at least) float32 -- done here in double
 * and rounded once, so that the only difference left against the GPU is the tensor core's own summation order. */
void qo_policy_forward_bf16(int n_layers, const int *dims, const float *const *W, const float *const *b,
                            const float *obs, int64_t n, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float cur[256], nxt[256];
        for (int k = 0; k < dims[0]; ++k) cur[k] = qo_bf16(obs[i * dims[0] + k]);
        for (int l = 0; l < n_layers; ++l) {
            const int in = dims[l], on = dims[l + 1];
            for (int o = 0; o < on; ++o) {
                double acc = (double)qo_bf16(b[l][o]);
                for (int k = 0; k < in; ++k) acc += (double)cur[k] * (double)qo_bf16(W[l][(size_t)o * in + k]);
                const float a = (float)acc;
                nxt[o] = (l + 1 < n_layers) ? qo_bf16(qo_act(a)) : a;
            }
            for (int o = 0; o < on; ++o) cur[o] = nxt[o];
        }
        for (int o = 0; o < dims[n_layers]; ++o) out[i * dims[n_layers] + o] = cur[o];
    }
}
