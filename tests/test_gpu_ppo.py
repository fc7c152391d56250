"""Row f2: GAE on the device against SB3's formula, and a short PPO run over the device rollout."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def PK():
    """the reference's policy_kwargs (`3D quad race.ipynb:784`)"""
    import torch
    return dict(activation_fn=torch.nn.ReLU, net_arch=[dict(pi=[120, 120, 120], vf=[120, 120, 120])], log_std_init=0)


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
    ppo = Q.PPO("MlpPolicy", env, policy_kwargs=PK(), n_steps=128, batch_size=16384, n_epochs=4, gamma=0.999, seed=0)
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


def test_buffer_evaluation_on_the_forward_kernel_matches_torch(tracks):
    """`evaluate="device"` (opt-in): values and old log-probs of the collected buffer come from two
    launches of the tcgen05 forward kernel instead of torch GEMMs.  Against the float32 torch pass on the same buffer: BF16
    operand rounding only; against the BF16 actor that sampled the actions: the stored log-prob is the log-density of the
    noise it drew, i.e. the PPO ratio of an unchanged policy is 1."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["e2e"]
    env = Q.Quadcopter3DGates(4096, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=2)
    env.disturbance_ranges = Q.training_disturbance_ranges()
    ppo = Q.PPO("MlpPolicy", env, policy_kwargs=PK(), n_steps=64, batch_size=32768, n_epochs=2, gamma=0.999, seed=2,
                evaluate="device")
    assert ppo.evaluate == "device" and ppo.critic is not None
    ppo.learn(iterations=3)  # move the weights off their initialisation (values of an untrained critic are ~0)
    b = ppo.collect_rollouts()
    T, n, d = 64, 4096, env.state_len
    dev = {k: b[k].clone() for k in ("values", "log_probs", "weights")}
    ppo.evaluate = "torch"
    ppo._evaluate_buffer(b, T, n, d)
    ppo.evaluate = "device"
    assert torch.equal(dev["weights"], b["weights"])
    scale = max(1.0, b["values"].abs().max().item())
    ev = (dev["values"] - b["values"]).abs().max().item() / scale
    em = (dev["values"] - b["values"]).abs().mean().item() / scale
    el = (dev["log_probs"] - b["log_probs"]).abs().max().item()
    print(f"device vs torch evaluation: values max {ev:.2e} mean {em:.2e} (scaled by {scale:.1f}), log-probs max {el:.2e}")
    assert ev < 6e-2 and em < 6e-3 and el < 0.15  # BF16 operands through four layers (measured 2.6e-2 / 2.7e-2 at the maximum)
    # the stored log-prob against the very mean the BF16 actor sampled around
    mean = torch.empty((n, 4), device="cuda")
    ppo.actor.forward(b["obs"][5], deterministic=True, mean_out=mean)
    lp = ppo.policy.log_prob(mean, b["raw_actions"][5])
    assert torch.allclose(lp, dev["log_probs"][5], atol=1e-5)
    tr = ppo.train()
    assert np.isfinite([tr["pg_loss"], tr["v_loss"], tr["approx_kl"]]).all() and tr["approx_kl"] < 0.1
    env.close()


def test_ppo_update_survives_degenerate_samples(tracks):
    """A tumbling quad's roll angle can wind to ~1e8 (tan(theta) near pi/2) while its episode goes on; such rows, and
    non-finite ones, must get zero weight instead of turning the networks into NaN."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["e2e"]
    env = Q.Quadcopter3DGates(2048, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=1)
    env.disturbance_ranges = Q.training_disturbance_ranges()
    ppo = Q.PPO("MlpPolicy", env, policy_kwargs=PK(), n_steps=32, batch_size=8192, n_epochs=2, gamma=0.999, seed=1)
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
    ppo = Q.PPO("MlpPolicy", env, policy_kwargs=PK(), n_steps=16, batch_size=4096, n_epochs=1, gamma=0.999, seed=2)
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
    again.learn(iterations=1, reset_num_timesteps=False)  # optimizer state restored: training continues (`:820`)
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


# ------------------------------------------------------------------------------------------------ SB3 drop-in (row f2)
def test_reference_training_cell_runs_with_only_the_import_changed(tmp_path, tracks):
    """The reference's training cell (`3D quad race.ipynb:765-795`) and its train() loop (`:816-823`), statement for
    statement, with `from stable_baselines3 import PPO` / `VecMonitor` replaced by this package's.  num_envs / n_steps are
    the reference's (100 x 1000 would take minutes through the host loop: shortened to 100 x 20, batch 500)."""
    import torch
    from optimal_quad_control_rl_b200 import PPO, VecMonitor, Quadcopter3DGates
    gate_pos, gate_yaw, start_pos = tracks["e2e"]
    models_dir, log_dir = str(tmp_path / "models" / "E2E"), str(tmp_path / "logs" / "E2E")
    np.random.seed(0)
    env = Quadcopter3DGates(num_envs=100, gates_pos=gate_pos, gate_yaw=gate_yaw, start_pos=start_pos, gates_ahead=1)
    test_env = Quadcopter3DGates(num_envs=10, gates_pos=gate_pos, gate_yaw=gate_yaw, start_pos=start_pos, gates_ahead=1,
                                 pause_if_collision=True)
    env = VecMonitor(env)
    disturbance_ranges = np.array([[-0.03, 0.03], [-0.03, 0.03], [-0.01, 0.01], [0, 0], [0, 0], [-0.5, 0.5]])
    env.venv.disturbance_ranges = disturbance_ranges
    test_env.disturbance_ranges = disturbance_ranges
    policy_kwargs = dict(activation_fn=torch.nn.ReLU, net_arch=[dict(pi=[120, 120, 120], vf=[120, 120, 120])], log_std_init=0)
    model = PPO("MlpPolicy", env, policy_kwargs=policy_kwargs, verbose=0, tensorboard_log=log_dir, n_steps=20, batch_size=500,
                n_epochs=10, gamma=0.999)
    print(model.policy)
    assert model.num_timesteps == 0 and model.rollout == "host" and model.bootstrap == "sb3_a8"
    test_env.reset()
    actions, _ = model.predict(test_env.states, deterministic=False)      # animate_policy's body (`:800-806`)
    states, rewards, dones, infos = test_env.step(actions)
    assert states.shape == (10, 24) and set(test_env.render()) >= {"x", "psi", "w4", "u1"}
    log_name = "zigzag"
    TIMESTEPS = model.n_steps * env.num_envs * 2                            # train() (`:816-823`), 2 rollouts per save
    for i in range(2):
        model.learn(total_timesteps=TIMESTEPS, reset_num_timesteps=False, tb_log_name=log_name)
        time_steps = model.num_timesteps
        model.save(models_dir + '/' + log_name + '/' + str(time_steps))
    assert model.num_timesteps == 2 * TIMESTEPS == 8000
    import os
    assert os.path.isfile(models_dir + "/zigzag/8000.zip")
    assert os.path.isfile(os.path.join(log_dir, "zigzag_1", "progress.jsonl"))
    loaded = PPO.load(models_dir + "/zigzag/8000.zip")                     # `:3985-3995`: no env
    network = list(loaded.policy.mlp_extractor.policy_net) + [loaded.policy.action_net]
    assert [m.out_features for m in network if hasattr(m, "out_features")] == [120, 120, 120, 4]
    network_std = loaded.policy.log_std.exp().cpu().detach().numpy()
    assert network_std.shape == (4,) and str(loaded.policy.action_dist).startswith("DiagGaussian")
    a1, _ = model.predict(states, deterministic=True)
    a2, _ = loaded.predict(states, deterministic=True)
    np.testing.assert_array_equal(a1, a2)
    h = model.history
    assert len(h) == 4 and all(np.isfinite(r["pg_loss"]) and np.isfinite(r["ep_rew_mean"]) for r in h)


def _fake_sb3_collect(env, policy_fn, value_fn, n_steps, gamma):
    """SB3 2.1 ``OnPolicyAlgorithm.collect_rollouts`` restated with NumPy buffers (test double of the real caller):
    obs are read from ``_last_obs`` AFTER ``env.step``; infos are used the way VecMonitor + SB3 use them."""
    n = env.num_envs
    last_obs = env.reset()
    buf = {"obs": [], "actions": [], "rewards": [], "dones": []}
    for t in range(n_steps):
        actions = policy_fn(last_obs)
        clipped = np.clip(actions, -1, 1)
        new_obs, rewards, dones, infos = env.step(clipped)
        assert new_obs.dtype == np.float32 and rewards.dtype == np.float32 and dones.dtype == np.bool_ and len(infos) == n
        rewards = rewards.copy()
        for idx, done in enumerate(dones):
            if done and infos[idx].get("terminal_observation") is not None and infos[idx].get("TimeLimit.truncated", False):
                rewards[idx] += gamma * value_fn(infos[idx]["terminal_observation"][None])[0]
        buf["obs"].append(last_obs.copy())        # rollout_buffer.add(self._last_obs, ...): AFTER the step
        buf["actions"].append(actions); buf["rewards"].append(rewards); buf["dones"].append(dones.copy())
        last_obs = new_obs
    return {k: np.stack(v) for k, v in buf.items()}, last_obs


@pytest.mark.parametrize("reset_rng", ["numpy", "device"])
def test_fake_sb3_collect_rollouts_over_the_numpy_facing_env(reset_rng, tracks):
    """SB3-shaped driver over the NumPy-facing GPU env: the observation handed out by step t must still be intact when
    the caller stores it after step t+1 (obs_buffers=2 pinned ring), infos behave like the reference's aliased dict,
    VecMonitor can copy them, float64 actions are accepted; in numpy-RNG mode the whole rollout matches the CPU oracle
    env driven by the same loop."""
    import optimal_quad_control_rl_b200 as Q
    from oracle import c_oracle as O
    gp, gy, sp = tracks["e2e"]
    n, T = 256, 40
    rng = np.random.default_rng(3)
    Wp = rng.normal(0, 0.3, (24, 4))
    policy_fn = lambda obs: (np.tanh(obs @ Wp) + rng.normal(0, 0.5, (len(obs), 4)))      # float64 actions on purpose
    value_fn = lambda obs: obs[:, :3].sum(1).astype(np.float32)
    env = Q.Quadcopter3DGates(n, gp, gy, sp, gates_ahead=1, reset_rng=reset_rng, seed=5, obs_buffers=2)
    env.disturbance_ranges = Q.training_disturbance_ranges()
    env.max_steps = 15                                                                    # time-outs happen
    np.random.seed(11)
    mon = Q.VecMonitor(env)
    seen = []
    inner_step = mon.step

    def spy_step(a):                                                                      # keep what the env handed out
        out = inner_step(a)
        seen.append((out[0], out[0].copy()))
        return out
    mon.step = spy_step
    buf, last = _fake_sb3_collect(mon, policy_fn, value_fn, T, 0.999)
    # (1) every observation array survived until it was stored one step later
    for t in range(len(seen) - 1):
        np.testing.assert_array_equal(buf["obs"][t + 1], seen[t][1])
    assert buf["dones"].sum() > n and mon.episode_count == buf["dones"].sum()
    if reset_rng == "numpy":  # (2) the same loop over the CPU oracle env, same np.random stream
        rng = np.random.default_rng(3)
        Wp2 = rng.normal(0, 0.3, (24, 4))
        policy2 = lambda obs: (np.tanh(obs @ Wp2) + rng.normal(0, 0.5, (len(obs), 4)))
        ora = O.OracleEnv("e2e", n, gp, gy, sp, gates_ahead=1)
        ora.disturbance_ranges = Q.training_disturbance_ranges()
        ora.max_steps = 15
        np.random.seed(11)
        ref, _ = _fake_sb3_collect(Q.VecMonitor(ora), policy2, value_fn, T, 0.999)
        np.testing.assert_array_equal(buf["dones"], ref["dones"])
        np.testing.assert_allclose(buf["obs"], ref["obs"], rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(buf["rewards"], ref["rewards"], rtol=2e-4, atol=2e-3)
    env.close()


def test_device_rollout_a8_bootstrap_and_host_rollout_agree_on_semantics(tracks):
    """bootstrap='sb3_a8' on the device path: rewards of the buffer == rewards of bootstrap='none' + the NumPy
    restatement of SB3's loop over the aliased infos, on the very same rollout."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    from optimal_quad_control_rl_b200.ppo import a8_bootstrap_
    gp, gy, sp = tracks["indi"]
    env = Q.Quadcopter3DGatesINDI(512, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=9)
    env.max_steps = 20
    ppo = Q.PPO("MlpPolicy", env, policy_kwargs=PK(), n_steps=48, batch_size=4096, n_epochs=1, gamma=0.999, seed=3,
                bootstrap="sb3_a8")
    assert ppo.rollout == "device"
    b = ppo.collect_rollouts()
    fl, dn = b["flags"].cpu().numpy(), b["dones"].cpu().numpy().astype(bool)
    assert ((fl & 2) != 0).any() and (((fl & 1) != 0) == dn).all()
    # recompute from the raw pieces: the device rewards minus the bootstrap must be what the env returned
    obs_next = b["obs"][1:].cpu().numpy()
    def V(o):
        with torch.no_grad():
            return ppo._values(torch.as_tensor(o, device=ppo.device)).cpu().numpy()
    add = np.zeros_like(fl, dtype=np.float32)
    for t in range(48):                                   # SB3's loop over the reference's aliased infos
        info = {}
        idx = np.flatnonzero(dn[t])
        if idx.size:
            info["terminal_observation"] = obs_next[t, idx[-1]]
        if ((fl[t] & 2) != 0).any():
            info["TimeLimit.truncated"] = True
        for i in idx:
            if info.get("terminal_observation") is not None and info.get("TimeLimit.truncated", False):
                add[t, i] = 0.999 * V(info["terminal_observation"][None])[0]
    env2 = Q.Quadcopter3DGatesINDI(512, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=9)
    env2.max_steps = 20
    plain = Q.PPO("MlpPolicy", env2, policy_kwargs=PK(), n_steps=48, batch_size=4096, n_epochs=1, gamma=0.999, seed=3,
                  bootstrap="none")
    b2 = plain.collect_rollouts()
    assert torch.equal(b2["dones"], b["dones"]) and torch.equal(b2["obs"], b["obs"])
    # (the values are TF32 GEMMs whose last bits depend on the batch shape: (T, D) there, (1, D) here)
    np.testing.assert_allclose(b["rewards"].cpu().numpy(), b2["rewards"].cpu().numpy() + add, rtol=1e-2, atol=2e-3)
    assert np.array_equal(b["rewards"].cpu().numpy() != b2["rewards"].cpu().numpy(), add != 0)
    assert np.abs(add).sum() > 0
    env.close(); env2.close()
