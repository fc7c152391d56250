"""Row f4 (CPU): the trajectory log keeps the reference's ``np.savez`` schema
(`3D quad race INDI inner loop.ipynb:647-695`); driven here over the oracle env."""
import numpy as np


def test_log_schema_and_values(tmp_path, tracks):
    from optimal_quad_control_rl_b200.trajectory import TrajectoryLog
    from oracle import c_oracle as O
    gp, gy, sp = tracks["indi"]
    env = O.OracleEnv("indi", 3, gp, gy, sp, gates_ahead=1)
    env.max_steps = 10000
    np.random.seed(0)
    env.reset()
    rng = np.random.default_rng(0)
    log = TrajectoryLog(env, index=1)
    states, acts, times = [], [], []
    for _ in range(50):
        a = rng.uniform(-1, 1, (3, 4)).astype(np.float32)
        env.step(a)
        log.record(a)
        states.append(env.world_states[1].copy()); acts.append(a[1]); times.append(env.step_counts[1] * env.dt)
    path = log.save("INDI_NET_sim", folder=str(tmp_path))
    z = np.load(path)
    assert list(z.keys()) == ["t", "x", "y", "z", "vx", "vy", "vz", "V", "phi", "theta", "psi", "u1", "u2", "u3", "u4", "u"]
    s, a = np.array(states), np.array(acts)
    np.testing.assert_array_equal(z["t"], np.array(times))
    for k, name in enumerate(["x", "y", "z", "vx", "vy", "vz", "phi", "theta", "psi"]):
        np.testing.assert_array_equal(z[name], s[:, k])
    np.testing.assert_array_equal(z["V"], np.sqrt(s[:, 3] ** 2 + s[:, 4] ** 2 + s[:, 5] ** 2))
    np.testing.assert_array_equal(z["u"], (a + 1) / 2)
    assert z["u"].shape == (50, 4) and len(log) == 50


def test_log_policy_run_uses_the_notebook_loop(tracks):
    from optimal_quad_control_rl_b200.trajectory import log_policy_run
    from oracle import c_oracle as O
    gp, gy, sp = tracks["e2e"]
    env = O.OracleEnv("e2e", 2, gp, gy, sp, gates_ahead=1)

    class Model:
        calls = 0

        def predict(self, obs, deterministic=False):
            Model.calls += 1
            assert obs.shape == (2, 24)
            return np.zeros((2, 4), np.float32), None

    np.random.seed(1)
    log = log_policy_run(Model(), env, 20, deterministic=True)
    d = log.as_dict()
    assert Model.calls == 20 and len(d["t"]) == 20 and np.all(d["u1"] == 0.5)
