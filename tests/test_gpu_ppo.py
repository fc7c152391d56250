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
