"""Pins the CPU oracle (oracle/quadsim_oracle.c + oracle/c_oracle.py) against vectors frozen from the
unmodified reference cells (oracle/make_golden.py).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_close, golden, scaled_err
from oracle import c_oracle as O

VARIANTS = ("e2e", "indi")


def make_env(variant, n, tracks, ga=1, pic=False, ranges=None):
    gp, gy, sp = tracks[variant]
    env = O.OracleEnv(variant, n, gp, gy, sp, gates_ahead=ga, pause_if_collision=pic)
    if ranges is not None:
        env.disturbance_ranges = ranges
    return env


def test_k1_residual_mlp_known_answer(tracks):
    """K1: the vector the reference prints in cell 4 (`3D quad race.ipynb:215,219`)."""
    k = golden("kat")
    env = make_env("e2e", 4, tracks)
    _, th, mo = O.euler(env, k["k1_states"], np.zeros((4, 4), np.float32), np.zeros((4, 6), np.float32))
    np.testing.assert_allclose(th[0, 0], 36.098232, rtol=1e-6)          # literal from the notebook's stored output
    np.testing.assert_allclose(mo[0], [0.2847767, -0.22512697, -0.05896095], rtol=2e-6)
    np.testing.assert_allclose(th, k["k1_thrust"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(mo, k["k1_moment"], rtol=2e-6, atol=1e-6)


def test_k2_reference_c_mlp_agrees(tracks):
    """oracle/_ref = the reference's generated nn_thrust.c / nn_moment.c; must match our MLP restatement."""
    L = O.ref_mlp_lib()
    if L is None:
        pytest.skip("oracle/_ref not built (reference not mounted at build time)")
    rng = np.random.default_rng(0)
    tw, mw = O.load_residual_weights()
    env = make_env("e2e", 1, tracks)
    for _ in range(64):
        ws = np.zeros((1, 16), np.float32)
        ws[0, 3:6] = rng.uniform(-8, 8, 3)       # attitude zero -> body velocity == world velocity
        ws[0, 9:16] = rng.uniform(-2, 2, 7)
        x = np.concatenate([ws[0, 12:16], ws[0, 3:6], ws[0, 9:12]]).astype(np.float32)
        yt = np.zeros(1, np.float32)
        ym = np.zeros(3, np.float32)
        L.nn_thrust_forward(x.ctypes.data_as(O._fp), yt.ctypes.data_as(O._fp))
        L.nn_moment_forward(x.ctypes.data_as(O._fp), ym.ctypes.data_as(O._fp))
        _, th, mo = O.euler(env, ws, np.zeros((1, 4), np.float32), np.zeros((1, 6), np.float32))
        assert th[0, 0] == yt[0] and (mo[0] == ym).all()   # same summation order -> bit-identical


@pytest.mark.parametrize("variant", VARIANTS)
def test_k3_gate_tables(variant, tracks):
    k = golden("kat")
    pr, yr = O.track_tables(k[f"{variant}_gate_pos"], k[f"{variant}_gate_yaw"])
    np.testing.assert_allclose(pr, k[f"{variant}_gate_pos_rel"], rtol=0, atol=3e-7)
    np.testing.assert_array_equal(yr, k[f"{variant}_gate_yaw_rel"])
    if variant == "indi":  # literals of c_code/nn_controller.c:40-60
        np.testing.assert_allclose(pr[0], [2.8284265995025635, 2.82842755317688, 0.0], atol=3e-7)
        assert yr[0] == np.float32(-4.71238899230957)


@pytest.mark.parametrize("variant", VARIANTS)
def test_single_step_euler(variant, tracks):
    g = golden(f"{variant}_single_step")
    env = make_env(variant, len(g["in_ws"]), tracks)
    new, th, mo = O.euler(env, g["in_ws"], g["in_act"], g.get("in_dist"))
    assert_close(new, g["new_states_raw"], "new_states_raw")
    # positions are pos + dt*vel: two rounded f32 ops -> must be bit-identical
    np.testing.assert_array_equal(new[:, 0:3], g["new_states_raw"][:, 0:3])
    if variant == "e2e":
        assert_close(th, g["mlp_thrust"], "thrust")
        assert_close(mo, g["mlp_moment"], "moment")
    print(variant, "max scaled err", scaled_err(new, g["new_states_raw"]).max(axis=0))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("branch", ["pic", "nrm", "pau"])
def test_single_step_branches(variant, branch, tracks):
    g = golden(f"{variant}_single_step")
    n = len(g["in_ws"])
    env = make_env(variant, n, tracks, pic=(branch == "pic"), ranges=g.get("disturbance_ranges"))
    env.force(g["in_ws"], g["in_tg"], g["in_sc"], g.get("in_dist"))
    assert_close(env.states, g["in_obs"], "obs before")
    prev_obs = env.states
    env.pause = branch == "pau"
    if branch == "nrm":
        np.random.seed(int(g["nrm_seed"]))
    obs, rew, done, infos = env.step(g["in_act"])
    np.testing.assert_array_equal(done, g[f"{branch}_done"])
    np.testing.assert_array_equal(env.target_gates, g[f"{branch}_tg"])
    np.testing.assert_array_equal(env.step_counts, g[f"{branch}_sc"])
    assert_close(rew, g[f"{branch}_rew"], "reward")
    assert_close(env.world_states, g[f"{branch}_ws"], "world_states")
    assert_close(obs, g[f"{branch}_obs"], "obs")
    if branch == "pau":
        assert obs is prev_obs and not done.any()
    if branch == "nrm":
        d = g["nrm_done"]
        # reset values are pure RNG + cast: bit-identical, and prove the draw order
        np.testing.assert_array_equal(env.world_states[d], g["nrm_ws"][d])
        if variant == "e2e":
            np.testing.assert_array_equal(env.disturbances, g["nrm_dist"])
        assert bool(infos[0].get("TimeLimit.truncated", False)) == bool(g["nrm_info_truncated"])
        assert_close(infos[0]["terminal_observation"], g["nrm_info_terminal_obs"], "terminal_observation")
        assert all(i is infos[0] for i in infos)
    assert g["n_gate_passed"] > 300 and g["n_done"] > 1000  # the fixture really exercises the flag logic


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("ga", [0, 2])
def test_obs_layout_other_gates_ahead(variant, ga, tracks):
    g = golden(f"{variant}_obs_ga{ga}")
    env = make_env(variant, len(g["in_ws"]), tracks, ga=ga, pic=True, ranges=g.get("disturbance_ranges"))
    env.force(g["in_ws"], g["in_tg"], g["in_sc"], g.get("in_dist"))
    assert env.states.shape == g["in_obs"].shape
    assert_close(env.states, g["in_obs"], "obs before")
    obs, rew, done, _ = env.step(g["in_act"])
    np.testing.assert_array_equal(done, g["pic_done"])
    np.testing.assert_array_equal(env.target_gates, g["pic_tg"])
    assert_close(obs, g["pic_obs"], "obs")
    assert_close(rew, g["pic_rew"], "reward")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", ["traj_n1", "traj_n16"])
def test_trajectory_teacher_forced(variant, name, tracks):
    """Config C1 (N=1, seed 0, 1000 steps) and an N=16 rollout: every step re-synced to the reference state."""
    g = golden(f"{variant}_{name}")
    steps, n = g["actions"].shape[:2]
    env = make_env(variant, n, tracks, ranges=g.get("disturbance_ranges"))
    env.max_steps = int(g["max_steps"])
    np.random.seed(int(g["np_seed"]))
    obs = env.reset()
    np.testing.assert_array_equal(env.world_states, g["ws"][0])   # reset draw order
    assert_close(obs, g["obs"][0], "reset obs")
    for t in range(steps):
        env.force(g["ws"][t], g["tg"][t], g["sc"][t], g["dist"][t] if "dist" in g else None)
        obs, rew, done, infos = env.step(g["actions"][t])
        np.testing.assert_array_equal(done, g["done"][t], err_msg=f"done @ {t}")
        np.testing.assert_array_equal(env.target_gates, g["tg"][t + 1])
        np.testing.assert_array_equal(env.step_counts, g["sc"][t + 1])
        assert_close(env.world_states, g["ws"][t + 1], f"ws @ {t}")
        assert_close(obs, g["obs"][t + 1], f"obs @ {t}")
        assert_close(rew, g["rew"][t], f"rew @ {t}")
        assert bool(infos[0].get("TimeLimit.truncated", False)) == bool(g["info_truncated"][t])
        if done.any():
            assert_close(infos[0]["terminal_observation"], g["info_terminal_obs"][t], "terminal_observation")
            np.testing.assert_array_equal(env.world_states[done], g["ws"][t + 1][done])


@pytest.mark.parametrize("variant", VARIANTS)
def test_trajectory_free_running_report(variant, tracks):
    """Free-running C1 is chaotic in the tail: report drift, gate only the first 20 steps."""
    g = golden(f"{variant}_traj_n1")
    env = make_env(variant, 1, tracks, ranges=g.get("disturbance_ranges"))
    np.random.seed(int(g["np_seed"]))
    env.reset()
    first_bad = None
    for t in range(g["actions"].shape[0]):
        env.step(g["actions"][t])
        e = scaled_err(env.world_states, g["ws"][t + 1]).max()
        if e > 1e-5 and first_bad is None:
            first_bad = t
    print(f"{variant}: free-running first step over 1e-5: {first_bad}")
    assert first_bad is None or first_bad >= 20


@pytest.mark.parametrize("variant", VARIANTS)
def test_teacher_set_2p20_digest(variant, tracks):
    """The full-size teacher-forced set of SURVEY section 8(d) (2**20 states incl. the adversarial generators): the
    oracle's done / target_gate / step_count arrays hash to the reference's, and every 64th row of its float outputs
    is within the gate of the reference's (digest frozen by oracle/make_golden.py --teacher-only)."""
    from oracle import make_golden as MG
    g = golden(f"{variant}_teacher_2p20_digest")
    n, k = int(g["n"]), int(g["stride"])
    track = tracks[variant]
    ws, tg, sc, act, dist = MG.teacher_set_inputs(variant, track, int(g["seed"]), n)
    assert MG.sha(np.concatenate([ws.ravel(), act.ravel()])) == str(g["sha_inputs"]), "input generator drifted"
    env = make_env(variant, n, tracks, pic=True, ranges=MG.R.training_disturbance_ranges() if variant == "e2e" else None)
    env.force(ws, tg, sc, dist)
    assert_close(env.states[::k], g["obs0"], "obs before")
    obs, rew, done, _ = env.step(act)
    assert MG.sha(done.astype(np.uint8)) == str(g["sha_done"])
    assert MG.sha(env.target_gates.astype(np.int64)) == str(g["sha_tg"])
    assert MG.sha(env.step_counts.astype(np.int64)) == str(g["sha_sc"])
    np.testing.assert_array_equal(env.world_states[::k, 0:3], g["ws"][:, 0:3])
    assert_close(env.world_states[::k], g["ws"], "world_states")
    assert_close(obs[::k], g["obs"], "obs")
    assert_close(rew[::k], g["rew"], "reward")
    assert int(done.sum()) == int(g["n_done"]) > n // 20


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_disturbance_obs_narrow_offcentre_ranges(dtype, tracks):
    """2*(d-lo)/(hi-lo)-1 for ranges like [10, 10.001]: subtract first, in the dtype of the ranges array."""
    g = golden("e2e_obs_offcentre")
    dt = np.float64 if dtype == "f64" else np.float32
    env = make_env("e2e", len(g["in_ws"]), tracks, ranges=g["disturbance_ranges"].astype(dt))
    env.force(g["in_ws"], g["in_tg"], g["in_sc"], g["in_dist"])
    assert_close(env.states, g[f"obs_{dtype}"], "obs")
