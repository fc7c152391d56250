"""Row f2: GAE on the device against SB3's formula, and a short PPO run over the device rollout."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gae_reference(rew, val, done, gamma, lam):
    """stable_baselines3.common.buffers.RolloutBuffer.compute_returns_and_advantage, restated:
    next_non_terminal comes from the done flag of the SAME step (SB3 stores episode_starts of the next step)."""
    T, n = rew.shape
    adv = np.zeros((T, n), np.float64)
    last = np.zeros(n, np.float64)
    for t in reversed(range(T)):
        nnt = 1.0 - done[t].astype(np.float64)
        delta = rew[t] + gamma * val[t + 1] * nnt - val[t]
        last = delta + gamma * lam * nnt * last
        adv[t] = last
    return adv, adv + val[:T]


@pytest.mark.parametrize("T,n", [(1, 1), (37, 1000), (256, 4096)])
def test_gae_kernel_matches_sb3_formula(T, n):
    import torch
    from optimal_quad_control_rl_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(T)
    rew = rng.normal(0, 1, (T, n)).astype(np.float32)
    val = rng.normal(0, 3, (T + 1, n)).astype(np.float32)
    done = (rng.uniform(size=(T, n)) < 0.05).astype(np.uint8)
    d = lambda a: torch.from_numpy(a).cuda()
    r, v, dn = d(rew), d(val), d(done)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    st = lib.qs_gae(L._vp(r.data_ptr()), L._vp(v.data_ptr()), L._vp(dn.data_ptr()), L._vp(adv.data_ptr()),
                    L._vp(ret.data_ptr()), n, T, 0.999, 0.95, L._vp(torch.cuda.current_stream().cuda_stream))
    assert st == 0
    a_ref, r_ref = gae_reference(rew, val, done, 0.999, 0.95)
    np.testing.assert_allclose(adv.cpu().numpy(), a_ref, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(ret.cpu().numpy(), r_ref, rtol=2e-4, atol=2e-4)


def test_ppo_short_run_improves_reward(tracks):
    """The reference's hyper-parameters (`3D quad race.ipynb:784-795`) on the INDI env: the reward per step after a
    handful of iterations must beat the untrained policy's, and the importance ratio starts at 1."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["indi"]
    env = Q.Quadcopter3DGatesINDI(4096, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=0)
    ppo = Q.PPO(env, n_steps=128, batch_size=16384, n_epochs=4, seed=0)
    b = ppo.collect_rollouts()
    with torch.no_grad():  # before any update: new log-prob == stored log-prob
        lp = ppo._log_prob(b["obs"][:4].reshape(-1, env.state_len), b["raw_actions"][:4].reshape(-1, 4))
    assert torch.allclose(lp, b["log_probs"][:4].reshape(-1), atol=1e-5)
    assert torch.isfinite(b["advantages"]).all() and torch.isfinite(b["returns"]).all()
    # the BF16 actor sampled around (almost) the float32 mean: |raw - mean_f32| / std is a standard normal
    with torch.no_grad():
        z = (b["raw_actions"][0] - ppo.pi(b["obs"][0])) / ppo.log_std.exp()
    assert abs(z.mean().item()) < 0.05 and abs(z.std().item() - 1) < 0.05
    ppo.learn(iterations=12)
    h = ppo.history
    print([round(r["reward_per_step"], 4) for r in h])
    assert all(np.isfinite(r["pg_loss"]) and np.isfinite(r["v_loss"]) for r in h)
    assert np.mean([r["reward_per_step"] for r in h[-3:]]) > np.mean([r["reward_per_step"] for r in h[:2]]) + 0.005
    a, _ = ppo.predict(env.states if env.states.any() else np.zeros((4096, env.state_len), np.float32), deterministic=True)
    assert a.shape == (4096, 4) and np.isfinite(a).all()


def test_ppo_update_survives_degenerate_samples(tracks):
    """A tumbling quad's roll angle can wind to ~1e8 (tan(theta) near pi/2) while its episode goes on; such rows, and
    non-finite ones, must get zero weight instead of turning the networks into NaN."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["e2e"]
    env = Q.Quadcopter3DGates(2048, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=1)
    env.disturbance_ranges = Q.training_disturbance_ranges()
    ppo = Q.PPO(env, n_steps=32, batch_size=8192, n_epochs=2, seed=1)
    orig = env.rollout

    def poisoned(actor, steps, buffers=None, **kw):
        b = orig(actor, steps, buffers=buffers, **kw)
        b["obs"][3, 5, 6] = 2.3e8
        b["obs"][4, 6, 0] = float("inf")
        b["obs"][5, 7, 2] = float("nan")
        b["raw_actions"][6, 8, 1] = float("nan")
        b["raw_actions"][7, 9] = 1e30
        return b

    env.rollout = poisoned
    b = ppo.collect_rollouts()
    w = b["weights"]
    assert w[3, 5] == 0 and w[4, 6] == 0 and w[5, 7] == 0 and w[6, 8] == 0 and w[7, 9] == 0 and w.mean() > 0.99
    for k in ("values", "log_probs", "advantages", "returns"):
        assert torch.isfinite(b[k]).all(), k
    tr = ppo.train()
    assert not tr["rolled_back"] and np.isfinite([tr["pg_loss"], tr["v_loss"], tr["approx_kl"]]).all()
    assert all(torch.isfinite(p).all() for p in [*ppo.pi.parameters(), *ppo.vf.parameters(), ppo.log_std])
    assert tr["approx_kl"] < 1.0 and tr["valid_frac"] < 1.0


def test_ppo_checkpoint_round_trip_and_controller_export(tmp_path, tracks):
    """model.save / PPO.load (`3D quad race.ipynb:823`, `:3985`) and the C export of the trained actor (row f3): the
    generated C network must command what the float32 actor commands."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["indi"]
    env = Q.Quadcopter3DGatesINDI(1024, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=2)
    ppo = Q.PPO(env, n_steps=16, batch_size=4096, n_epochs=1, seed=2)
    ppo.learn(iterations=2)
    path = ppo.save(str(tmp_path / "models" / "indi" / str(ppo.num_timesteps)))
    env2 = Q.Quadcopter3DGatesINDI(1024, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=2)
    again = Q.PPO.load(path, env2)
    assert again.num_timesteps == ppo.num_timesteps == 2 * 16 * 1024 and len(again.history) == 2
    for a, b in zip(ppo.pi.parameters(), again.pi.parameters()):
        assert torch.equal(a, b)
    obs = env.reset()
    a1, _ = ppo.predict(obs, deterministic=True)
    a2, _ = again.predict(obs, deterministic=True)
    np.testing.assert_array_equal(a1, a2)
    again.learn(iterations=1)  # optimizer state restored: training continues
    assert again.num_timesteps == 3 * 16 * 1024
    # ---- C export of the actor
    files = Q.export_controller(ppo, env, str(tmp_path / "c_code"))
    assert len(files) == 4
    ctl = Q.CController(Q.build_controller(str(tmp_path / "c_code")), "indi")
    ctl.deterministic = True
    w, b = ppo._pi_arrays()
    x = obs[:32].astype(np.float64)
    for l, (wl, bl) in enumerate(zip(w, b)):  # float64 restatement of the actor
        x = x @ wl.astype(np.float64).T + bl
        if l < len(w) - 1:
            x = np.maximum(x, 0)
    got = np.stack([ctl.forward(o) for o in obs[:32]])
    np.testing.assert_allclose(got, np.clip(x, -1, 1), atol=1e-5)
    env.close(); env2.close()


def test_trajectory_log_over_the_gpu_env(tmp_path, tracks):
    """Row f4: the reference's logging loop (`3D quad race INDI inner loop.ipynb:647-695`) against the GPU env."""
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["indi"]
    env = Q.Quadcopter3DGatesINDI(64, gp, gy, sp, gates_ahead=1, pause_if_collision=True)
    env.max_steps = 10000
    pol = type("M", (), {"predict": lambda self, obs, deterministic=False: (np.full((64, 4), 0.1, np.float32), None)})()
    np.random.seed(0)
    log = Q.log_policy_run(pol, env, 30, index=5)
    d = log.as_dict()
    ws = env.world_states
    assert len(log) == 30 and d["x"][-1] == ws[5, 0] and d["psi"][-1] == ws[5, 8]
    np.testing.assert_allclose(d["t"], np.arange(1, 31) * np.float32(0.01), rtol=1e-6)
    assert np.allclose(d["u"], 0.55)
    z = np.load(log.save("gpu_run", folder=str(tmp_path)))
    assert z["u"].shape == (30, 4)
    r = env.render()
    assert set(r) == {"x", "y", "z", "vx", "vy", "vz", "phi", "theta", "psi", "p", "q", "r", "T", "u1", "u2", "u3", "u4"}
    env.close()
