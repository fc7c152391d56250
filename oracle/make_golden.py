"""TEST INFRASTRUCTURE — freezes outputs of the *reference itself* into ``tests/golden/``.

Run in the build container (needs ``/root/reference``):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz and the packaged residual weights

Every array written here comes from the unmodified reference cells exec'd by ``oracle/reference_exec.py``
(``Quadcopter3DGates``, ``f_func``, ``thrust_moment_model_world_states``).  The only code of ours on the path is
input generation (seeded NumPy) and book-keeping.  Library versions are recorded in ``tests/golden/MANIFEST.json``
because the reference's last-bit behaviour depends on sympy's printed expression and NumPy's f32 ufuncs.

Files
-----
kat.npz                     K1 residual-MLP known answer (`3D quad race.ipynb:248-266`), K3 gate tables of both tracks
{e2e,indi}_single_step.npz  teacher-forced single steps over random + adversarial (near gate plane / ground /
                            bounds / time-limit) states, in all three step_wait branches (`:568-585`)
{e2e,indi}_traj_n1.npz      config C1: N=1, np.random.seed(0), 1000 steps, actions default_rng(1)
{e2e,indi}_traj_n16.npz     N=16 seeded rollout with resets (RNG-order parity of reset_)
{e2e,indi}_obs_ga{0,2}.npz  observation layout for other gates_ahead values
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_exec as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NS = {"e2e": 16, "indi": 13}


# ------------------------------------------------------------------------------------------------ inputs
def random_states(variant, m, rng, track):
    """SURVEY §8d teacher-forced sampling box."""
    ns = NS[variant]
    ws = np.zeros((m, ns), np.float32)
    ws[:, 0:2] = rng.uniform(-4, 4, (m, 2))
    ws[:, 2] = rng.uniform(-3, 0, m)
    ws[:, 3:6] = rng.uniform(-8, 8, (m, 3))
    ws[:, 6:8] = rng.uniform(-1.2, 1.2, (m, 2))
    ws[:, 8] = rng.uniform(-np.pi, np.pi, m)
    ws[:, 9:12] = rng.uniform(-6, 6, (m, 3))
    ws[:, 12:] = rng.uniform(-1, 1, (m, ns - 12))
    return ws


def gate_crossing_states(variant, m, rng, track):
    """States a few cm in front of their target gate flying through it; lateral/vertical offsets straddle the
    +-0.5 m gate half-width so gate_passed / gate_collision / neither all occur (`3D quad race.ipynb:528-534`)."""
    gp, gy, _ = track
    gp = gp.astype(np.float32)
    gy = gy.astype(np.float32)
    ng = len(gy)
    tg = rng.integers(0, ng, m)
    ws = random_states(variant, m, rng, track)
    n = np.stack([np.cos(gy[tg]), np.sin(gy[tg])], 1)
    t = np.stack([-n[:, 1], n[:, 0]], 1)
    s = rng.uniform(0.0, 0.06, m)[:, None]
    lat = rng.choice([0.0, 0.3, 0.49, 0.499, 0.5, 0.501, 0.52, 0.7], m) * rng.choice([-1, 1], m)
    lat = lat + rng.normal(0, 1e-3, m) * (rng.random(m) < 0.5)
    hz = rng.choice([0.0, 0.2, 0.49, 0.5, 0.51, 0.8], m) * rng.choice([-1, 1], m)
    speed = rng.uniform(1.0, 12.0, m)[:, None]
    ws[:, 0:2] = gp[tg, 0:2] - s * n + lat[:, None] * t
    ws[:, 2] = gp[tg, 2] + hz
    ws[:, 3:5] = speed * n + rng.normal(0, 0.5, (m, 2))
    ws[:, 5] = rng.normal(0, 1.0, m)
    ws[:, 6:8] = rng.uniform(-0.6, 0.6, (m, 2))
    return ws.astype(np.float32), tg


def edge_states(variant, m, rng, track):
    """Ground (z>0), |x|,|y|>10, |rates|>1000 and exact-threshold cases (`:543-550`)."""
    ws = random_states(variant, m, rng, track)
    k = m // 4
    ws[:k, 2] = rng.choice([-1e-3, -1e-4, 0.0, 1e-4, 1e-3, -0.02, 0.02], k)
    ws[:k, 5] = rng.uniform(-3, 3, k)
    ws[k:2 * k, 0] = rng.choice([9.9, 9.99, 10.0, 10.01, -9.99, -10.0, -10.02], k)
    ws[k:2 * k, 3] = rng.uniform(-4, 4, k)
    ws[2 * k:3 * k, 1] = rng.choice([9.95, 10.0, 10.05, -9.95, -10.0, -10.05], k)
    ws[3 * k:, 9 + rng.integers(0, 3, m - 3 * k)] = rng.choice([999.0, 1000.0, 1001.0, -1000.5, 1500.0], m - 3 * k)
    ws[3 * k:, 6:8] = rng.uniform(-0.3, 0.3, (m - 3 * k, 2))
    return ws.astype(np.float32)


def single_step_inputs(variant, track, seed):
    rng = np.random.default_rng(seed)
    ng = len(track[1])
    a, b, c = 4096, 3072, 1024
    ws_a = random_states(variant, a, rng, track)
    ws_b, tg_b = gate_crossing_states(variant, b, rng, track)
    ws_c = edge_states(variant, c, rng, track)
    ws = np.concatenate([ws_a, ws_b, ws_c]).astype(np.float32)
    m = len(ws)
    tg = np.concatenate([rng.integers(0, ng, a), tg_b, rng.integers(0, ng, c)]).astype(np.int64)
    sc = rng.integers(0, 1198, m).astype(np.int64)
    sc[rng.random(m) < 0.03] = 1198  # -> 1199 after the step: one below the limit
    sc[rng.random(m) < 0.03] = 1199  # -> 1200: time-limit reached exactly
    sc[rng.random(m) < 0.01] = 1500
    act = rng.uniform(-1, 1, (m, 4)).astype(np.float32)
    dist = None
    if variant == "e2e":
        dr = R.training_disturbance_ranges()
        dist = rng.uniform(dr[:, 0], dr[:, 1], (m, 6)).astype(np.float32)
    return ws, tg, sc, act, dist


def teacher_set_inputs(variant, track, seed, n=1 << 20):
    """SURVEY section 8(d)'s teacher-forced parity set at full size: n states -- half from the uniform sampling box,
    3/8 flying through / past their gate, 1/8 on the ground / bounds / rate thresholds -- with counters around the
    time limit.  Pure NumPy (PCG64): the same arrays on every machine, so a digest of the reference's answers can be
    checked on the GPU box where the reference itself is absent."""
    rng = np.random.default_rng(seed)
    ng = len(track[1])
    a, b = n // 2, (3 * n) // 8
    c = n - a - b
    ws_a = random_states(variant, a, rng, track)
    ws_b, tg_b = gate_crossing_states(variant, b, rng, track)
    ws_c = edge_states(variant, c, rng, track)
    ws = np.concatenate([ws_a, ws_b, ws_c]).astype(np.float32)
    tg = np.concatenate([rng.integers(0, ng, a), tg_b, rng.integers(0, ng, c)]).astype(np.int64)
    sc = rng.integers(0, 1198, n).astype(np.int64)
    sc[rng.random(n) < 0.03] = 1198
    sc[rng.random(n) < 0.03] = 1199
    sc[rng.random(n) < 0.01] = 1500
    act = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    dist = None
    if variant == "e2e":
        dr = R.training_disturbance_ranges()
        dist = rng.uniform(dr[:, 0], dr[:, 1], (n, 6)).astype(np.float32)
    perm = rng.permutation(n)  # mix the three kinds across warps / tiles
    return ws[perm], tg[perm], sc[perm], act[perm], (dist[perm] if dist is not None else None)


TEACHER_SEED = {"e2e": 2020, "indi": 2021}
TEACHER_STRIDE = 64  # every 64th env of the reference's float outputs is kept in the digest


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def teacher_reference_step(variant, ws, tg, sc, act, dist, track):
    """One pause_if_collision step of the UNMODIFIED reference env over the teacher set (no RNG involved)."""
    n = len(ws)
    env = R.make_reference_env(variant, n, gates_ahead=1, pause_if_collision=True, track=track)
    force_state(env, ws, tg, sc, dist)
    obs0 = env.states.copy()
    obs, rew, done, _ = env.step(act)
    return dict(obs0=obs0, obs=obs.copy(), rew=rew.copy(), done=done.copy(), ws=env.world_states.copy(),
                tg=env.target_gates.copy(), sc=env.step_counts.copy())


def make_teacher_digest(variant, n=1 << 20):
    """2**20-state teacher-forced digest: SHA-256 of the reference's done / target_gate / step_count arrays (bit-exact
    quantities) plus every 64th row of its float outputs.  ~3 MB per variant instead of ~250 MB."""
    track = R.zigzag_track() if variant == "e2e" else R.rectangle_track()
    seed = TEACHER_SEED[variant]
    ws, tg, sc, act, dist = teacher_set_inputs(variant, track, seed, n)
    r = teacher_reference_step(variant, ws, tg, sc, act, dist, track)
    k = TEACHER_STRIDE
    out = dict(n=np.int64(n), seed=np.int64(seed), stride=np.int64(k),
               sha_inputs=sha(np.concatenate([ws.ravel(), act.ravel()])), sha_done=sha(r["done"].astype(np.uint8)),
               sha_tg=sha(r["tg"].astype(np.int64)), sha_sc=sha(r["sc"].astype(np.int64)),
               n_done=np.int64(r["done"].sum()), n_gate_passed=np.int64((r["tg"] != tg).sum()),
               obs0=r["obs0"][::k], obs=r["obs"][::k], rew=r["rew"][::k], ws=r["ws"][::k],
               rew_sum=np.float64(r["rew"].astype(np.float64).sum()))
    np.savez_compressed(os.path.join(GOLD, f"{variant}_teacher_2p20_digest.npz"), **out)
    return {kk: (str(v) if isinstance(v, str) else (int(v) if np.ndim(v) == 0 and np.issubdtype(np.asarray(v).dtype, np.integer)
                                                     else (float(v) if np.ndim(v) == 0 else list(np.shape(v)))))
            for kk, v in out.items()}


# ------------------------------------------------------------------------------------------------ helpers
def force_state(env, ws, tg, sc, dist):
    env.world_states = ws.copy()
    env.target_gates = tg.copy()
    env.step_counts = sc.copy()
    if dist is not None:
        env.disturbances = dist.copy()
    env.update_states()


def raw_euler_step(variant, G, env, ws, act, dist):
    """new_states exactly as step_wait forms them (`3D quad race.ipynb:503-512`, INDI `:304`)."""
    out = {}
    if variant == "e2e":
        d = np.zeros((len(ws), 6), np.float32)
        thrust, moment = G["thrust_moment_model_world_states"](ws)
        d[:, 0:3] = moment
        d[:, 5:6] = thrust
        d += dist
        out["mlp_thrust"], out["mlp_moment"] = thrust, moment
        out["body_velocity"] = G["get_body_velocity"](ws.T).T
        new = ws + env.dt * G["f_func"](ws.T, act.T, d.T).T
    else:
        new = ws + env.dt * G["f_func"](ws.T, act.T).T
    assert new.dtype == np.float32
    out["new_states_raw"] = new
    return out


def snapshot(env, variant):
    s = dict(ws=env.world_states.copy(), tg=env.target_gates.copy(), sc=env.step_counts.copy(),
             obs=env.states.copy())
    if variant == "e2e":
        s["dist"] = env.disturbances.copy()
    return s


# ------------------------------------------------------------------------------------------------ fixtures
def make_single_step(variant, seed):
    G = R.load_reference(variant)
    track = R.zigzag_track() if variant == "e2e" else R.rectangle_track()
    ws, tg, sc, act, dist = single_step_inputs(variant, track, seed)
    m = len(ws)
    out = dict(in_ws=ws, in_tg=tg, in_sc=sc, in_act=act)
    if dist is not None:
        out["in_dist"] = dist
        out["disturbance_ranges"] = R.training_disturbance_ranges()

    env = R.make_reference_env(variant, m, gates_ahead=1, track=track)
    out.update(raw_euler_step(variant, G, env, ws, act, dist))
    force_state(env, ws, tg, sc, dist)
    out["in_obs"] = env.states.copy()

    # branch 2: pause_if_collision (advance only non-done envs, no reset)
    env = R.make_reference_env(variant, m, gates_ahead=1, pause_if_collision=True, track=track)
    force_state(env, ws, tg, sc, dist)
    obs, rew, done, _ = env.step(act)
    snap = snapshot(env, variant)
    out.update(pic_obs=obs.copy(), pic_rew=rew.copy(), pic_done=done.copy(), pic_ws=snap["ws"], pic_tg=snap["tg"],
               pic_sc=snap["sc"])

    # branch 3: normal (advance everything, reset the done ones from the global NumPy stream)
    env = R.make_reference_env(variant, m, gates_ahead=1, track=track)
    force_state(env, ws, tg, sc, dist)
    np.random.seed(seed + 100)
    obs, rew, done, infos = env.step(act)
    snap = snapshot(env, variant)
    out.update(nrm_seed=np.int64(seed + 100), nrm_obs=obs.copy(), nrm_rew=rew.copy(), nrm_done=done.copy(),
               nrm_ws=snap["ws"], nrm_tg=snap["tg"], nrm_sc=snap["sc"])
    if variant == "e2e":
        out["nrm_dist"] = snap["dist"]
    out["nrm_info_truncated"] = np.bool_(infos[0].get("TimeLimit.truncated", False))
    out["nrm_info_terminal_obs"] = np.asarray(infos[0].get("terminal_observation", np.zeros(0, np.float32)))

    # branch 1: pause (nothing advances, dones cleared, counters and gate index still move)
    env = R.make_reference_env(variant, m, gates_ahead=1, track=track)
    force_state(env, ws, tg, sc, dist)
    env.pause = True
    obs, rew, done, _ = env.step(act)
    snap = snapshot(env, variant)
    out.update(pau_obs=obs.copy(), pau_rew=rew.copy(), pau_done=done.copy(), pau_ws=snap["ws"], pau_tg=snap["tg"],
               pau_sc=snap["sc"])
    # flag statistics so the test can assert coverage
    out["n_done"] = np.int64(out["pic_done"].sum())
    out["n_gate_passed"] = np.int64((out["pic_tg"] != tg).sum())
    np.savez_compressed(os.path.join(GOLD, f"{variant}_single_step.npz"), **out)
    return {k: (int(v) if np.ndim(v) == 0 else list(np.shape(v))) for k, v in out.items()}


def make_traj(variant, n, steps, np_seed, act_seed, name, max_steps=None):
    env = R.make_reference_env(variant, n, gates_ahead=1)
    if max_steps is not None:
        env.max_steps = max_steps
    acts = np.random.default_rng(act_seed).uniform(-1, 1, (steps, n, 4)).astype(np.float32)
    np.random.seed(np_seed)
    obs0 = env.reset()
    rec = {k: [v] for k, v in snapshot(env, variant).items()}
    rews, dones, trunc, term = [], [], [], []
    assert obs0 is env.states
    for t in range(steps):
        obs, rew, done, infos = env.step(acts[t])
        for k, v in snapshot(env, variant).items():
            rec[k].append(v)
        rews.append(rew.copy())
        dones.append(done.copy())
        trunc.append(bool(infos[0].get("TimeLimit.truncated", False)))
        term.append(np.asarray(infos[0]["terminal_observation"]).copy() if "terminal_observation" in infos[0]
                    else np.full(obs.shape[1], np.nan, np.float32))
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(actions=acts, rew=np.stack(rews), done=np.stack(dones), info_truncated=np.array(trunc),
               info_terminal_obs=np.stack(term), np_seed=np.int64(np_seed), act_seed=np.int64(act_seed),
               max_steps=np.int64(env.max_steps))
    if variant == "e2e":
        out["disturbance_ranges"] = R.training_disturbance_ranges()
    np.savez_compressed(os.path.join(GOLD, f"{variant}_{name}.npz"), **out)
    return dict(steps=steps, n=n, n_done=int(out["done"].sum()))


def make_obs_ga(variant, ga, seed):
    track = R.zigzag_track() if variant == "e2e" else R.rectangle_track()
    ws, tg, sc, act, dist = single_step_inputs(variant, track, seed)
    ws, tg, sc, act = ws[:512], tg[:512], sc[:512], act[:512]
    dist = dist[:512] if dist is not None else None
    out = dict(in_ws=ws, in_tg=tg, in_sc=sc, in_act=act)
    # float32 ranges here (the ctor default dtype) + one degenerate row -> exercises the lo==hi widening branch
    env = R.make_reference_env(variant, len(ws), gates_ahead=ga, track=track, pause_if_collision=True,
                               disturbance_ranges=None)
    if variant == "e2e":
        env.disturbance_ranges = R.training_disturbance_ranges().astype(np.float32)
        out["in_dist"] = dist
        out["disturbance_ranges"] = env.disturbance_ranges.copy()
    force_state(env, ws, tg, sc, dist)
    out["in_obs"] = env.states.copy()
    obs, rew, done, _ = env.step(act)
    out.update(pic_obs=obs.copy(), pic_rew=rew.copy(), pic_done=done.copy(), pic_ws=env.world_states.copy(),
               pic_tg=env.target_gates.copy())
    np.savez_compressed(os.path.join(GOLD, f"{variant}_obs_ga{ga}.npz"), **out)


def make_obs_offcentre(seed=77):
    """Disturbance observation 2*(d-lo)/(hi-lo)-1 (`3D quad race.ipynb:414-448`) for NARROW ranges FAR from zero, as a
    float64 and as a float32 ranges array: a folded d*scale+offset form loses ~1e-3 here."""
    track = R.zigzag_track()
    rng = np.random.default_rng(seed)
    m = 512
    ws, tg, sc, act, _ = single_step_inputs("e2e", track, seed)
    ws, tg, sc = ws[:m], tg[:m], sc[:m]
    ranges = np.array([[10.0, 10.001], [-5.0, -4.99], [0.2, 0.2003], [0, 0], [0, 0], [100.0, 100.01]])
    dist = rng.uniform(ranges[:, 0], ranges[:, 1], (m, 6)).astype(np.float32)
    out = dict(in_ws=ws, in_tg=tg, in_sc=sc, in_dist=dist, disturbance_ranges=ranges)
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        env = R.make_reference_env("e2e", m, gates_ahead=1, track=track, disturbance_ranges=None)
        env.disturbance_ranges = ranges.astype(dt)
        force_state(env, ws, tg, sc, dist)
        out[f"obs_{tag}"] = env.states.copy()
    np.savez_compressed(os.path.join(GOLD, "e2e_obs_offcentre.npz"), **out)
    return {k: list(np.shape(v)) for k, v in out.items()}


def make_kat():
    G = R.load_reference("e2e")
    st = np.zeros((4, 16), np.float32)
    st[0] = [0, 1, 2, 3, 4, 5, 0, 0, 0, 9, 10, 11, 12, 13, 14, 15]
    st[1:] = np.random.default_rng(7).standard_normal((3, 16)).astype(np.float32)
    thrust, moment = G["thrust_moment_model_world_states"](st)
    out = dict(k1_states=st, k1_thrust=thrust, k1_moment=moment, k1_vb=G["get_body_velocity"](st.T).T)
    for variant, track in (("e2e", R.zigzag_track()), ("indi", R.rectangle_track())):
        env = R.make_reference_env(variant, 1, gates_ahead=1, track=track)
        out[f"{variant}_gate_pos"] = env.gate_pos
        out[f"{variant}_gate_yaw"] = env.gate_yaw
        out[f"{variant}_gate_pos_rel"] = env.gate_pos_rel
        out[f"{variant}_gate_yaw_rel"] = env.gate_yaw_rel
        out[f"{variant}_start_pos"] = env.start_pos
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **out)


def export_residual_weights():
    """thrust_model.pt / moment_model.pt -> packaged f32 arrays (weights are data; the product needs them)."""
    G = R.load_reference("e2e")
    t, m = G["thrust_model"], G["moment_model"]
    arrs = dict(thrust_w1=t[0].weight, thrust_b1=t[0].bias, thrust_w2=t[2].weight, thrust_b2=t[2].bias,
                moment_w1=m[0].weight, moment_b1=m[0].bias, moment_w2=m[2].weight, moment_b2=m[2].bias)
    arrs = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in arrs.items()}
    dst = os.path.join(ROOT, "optimal_quad_control_rl_b200", "data", "residual_mlp.npz")
    np.savez(dst, **arrs)
    return {k: list(v.shape) for k, v in arrs.items()}


def main():
    import sympy
    import torch

    os.makedirs(GOLD, exist_ok=True)
    if "--teacher-only" in sys.argv:  # add / refresh the 2**20-state digests without touching the other fixtures
        with open(os.path.join(GOLD, "MANIFEST.json")) as f:
            manifest = json.load(f)
        for variant in ("e2e", "indi"):
            manifest["files"][f"{variant}_teacher_2p20_digest"] = make_teacher_digest(variant)
        manifest["files"]["e2e_obs_offcentre"] = make_obs_offcentre()
        with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
        return
    manifest = dict(generator="oracle/make_golden.py", reference_commit="7aaa3f689e0991c74132274c9f0bb8c72c2ad3a1",
                    numpy=np.__version__, sympy=sympy.__version__, torch=torch.__version__, files={})
    manifest["residual_weights"] = export_residual_weights()
    make_kat()
    for variant, seed in (("e2e", 11), ("indi", 12)):
        manifest["files"][f"{variant}_single_step"] = make_single_step(variant, seed)
        manifest["files"][f"{variant}_traj_n1"] = make_traj(variant, 1, 1000, 0, 1, "traj_n1")
        manifest["files"][f"{variant}_traj_n16"] = make_traj(variant, 16, 400, 2, 3, "traj_n16", max_steps=150)
        for ga in (0, 2):
            make_obs_ga(variant, ga, seed + 20 + ga)
        manifest["files"][f"{variant}_teacher_2p20_digest"] = make_teacher_digest(variant)
    manifest["files"]["e2e_obs_offcentre"] = make_obs_offcentre()
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(GOLD)):
        print(f"{os.path.getsize(os.path.join(GOLD, fn)):>9d}  {fn}")


if __name__ == "__main__":
    main()
