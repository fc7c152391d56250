"""Row f2, host-side logic that needs no GPU: the a8 bootstrap as a pure-torch function against a NumPy restatement of
SB3's loop over the reference's aliased infos, VecMonitor, and the SB3-shaped policy object."""
import numpy as np
import pytest
import torch


def sb3_loop_bootstrap(rewards, dones, flags, obs_next, value_np, gamma):
    """SB3 2.1 collect_rollouts bootstrap applied to infos built like `3D quad race.ipynb:589-594` builds them."""
    out = rewards.copy()
    T, n = rewards.shape
    for t in range(T):
        infos = [{}] * n                                # ONE dict, aliased
        for i in range(n):
            if dones[t, i]:
                infos[i]["terminal_observation"] = obs_next[t, i]
            if flags[t, i] & 2:
                infos[i]["TimeLimit.truncated"] = True
        for idx, done in enumerate(dones[t]):
            if done and infos[idx].get("terminal_observation") is not None and infos[idx].get("TimeLimit.truncated", False):
                out[t, idx] += gamma * value_np(infos[idx]["terminal_observation"][None])[0]
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_a8_bootstrap_matches_sb3_loop_over_aliased_infos(seed):
    from optimal_quad_control_rl_b200.ppo import a8_bootstrap_
    rng = np.random.default_rng(seed)
    T, n, d = 30, 37, 5
    dones = rng.uniform(size=(T, n)) < 0.08
    trunc = dones & (rng.uniform(size=(T, n)) < 0.3)
    dones[3] = False; trunc[3] = False                   # a step with no done env
    dones[4] = False; dones[4, 7] = True; trunc[4] = False; trunc[4, 7] = True   # the only done env is the truncated one
    flags = (dones.astype(np.uint8) * 1) | (trunc.astype(np.uint8) * 2)
    rewards = rng.normal(0, 1, (T, n)).astype(np.float32)
    obs_next = rng.normal(0, 1, (T, n, d)).astype(np.float32)
    w = rng.normal(0, 1, d).astype(np.float32)
    value_np = lambda o: (o @ w).astype(np.float32)
    want = sb3_loop_bootstrap(rewards, dones, flags, obs_next, value_np, 0.999)
    got = a8_bootstrap_(torch.from_numpy(rewards.copy()), torch.from_numpy(dones.astype(np.uint8)), torch.from_numpy(flags),
                        torch.from_numpy(obs_next), lambda o: o @ torch.from_numpy(w), 0.999).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    assert np.abs(want - rewards).sum() > 0
    # a step where some env is done but none truncated is untouched; so is every not-done env
    no_tr = ~trunc.any(1)
    np.testing.assert_array_equal(got[no_tr], rewards[no_tr])
    np.testing.assert_array_equal(got[~dones], rewards[~dones])


class _FakeVecEnv:
    """4 envs; env i ends its episode every (i+2) steps; infos = ONE aliased dict like the reference."""

    def __init__(self):
        self.num_envs, self.t, self.poke = 4, np.zeros(4, int), None

    def reset(self):
        self.t[:] = 0
        return np.zeros((4, 3), np.float32)

    def step_async(self, a):
        self.a = a

    def step_wait(self):
        self.t += 1
        dones = self.t % (np.arange(4) + 2) == 0
        info = {}
        if dones.any():
            info["terminal_observation"] = np.full(3, float(np.flatnonzero(dones)[-1]), np.float32)
        return np.zeros((4, 3), np.float32), np.ones(4, np.float32), dones, [info] * 4


def test_vecmonitor_copies_infos_and_forwards_attributes():
    from optimal_quad_control_rl_b200.ppo import VecMonitor
    raw = _FakeVecEnv()
    env = VecMonitor(raw)
    env.venv.poke = 5                                     # `env.venv.disturbance_ranges = ...` (`:780`)
    assert raw.poke == 5 and env.poke == 5 and env.num_envs == 4
    env.reset()
    lens = []
    for _ in range(12):
        obs, rew, dones, infos = env.step(np.zeros((4, 4)))
        for i in np.flatnonzero(dones):
            assert infos[i]["episode"]["l"] == i + 2 and infos[i]["episode"]["r"] == float(i + 2)
            assert "terminal_observation" in infos[i]
            lens.append(infos[i]["episode"]["l"])
        for i in np.flatnonzero(~dones):
            assert "episode" not in infos[i]
    assert env.episode_count == len(lens) == 6 + 4 + 3 + 2
    with pytest.raises(AttributeError):
        env.no_such_attribute


def test_policy_object_has_sb3_layout():
    from optimal_quad_control_rl_b200.ppo import ActorCriticPolicy, _parse_net_arch
    assert _parse_net_arch([dict(pi=[120, 120, 120], vf=[64])]) == ([120, 120, 120], [64])   # the reference's spelling
    assert _parse_net_arch(dict(pi=[8], vf=[9, 9])) == ([8], [9, 9])
    assert _parse_net_arch([32, 32]) == ([32, 32], [32, 32])
    assert _parse_net_arch(None) == ([64, 64], [64, 64])
    with pytest.raises(NotImplementedError):
        _parse_net_arch([64, dict(pi=[8], vf=[8])])
    torch.manual_seed(0)
    pol = ActorCriticPolicy(24, 4, net_arch=[dict(pi=[120, 120, 120], vf=[120, 120, 120])], activation_fn=torch.nn.ReLU,
                            log_std_init=0)
    network = torch.nn.Sequential(*(list(pol.mlp_extractor.policy_net) + [pol.action_net]))   # `:3988-3990`
    obs = torch.randn(7, 24)
    a, v, lp = pol(obs, deterministic=True)
    assert torch.allclose(a, network(obs)) and v.shape == (7, 1) and lp.shape == (7,)
    assert torch.equal(pol.log_std.exp(), torch.ones(4))
    v2, lp2, ent = pol.evaluate_actions(obs, a)
    assert torch.allclose(lp, lp2) and torch.allclose(v, v2) and ent.shape == (7,)
    # SB3's orthogonal init: action net gain 0.01 -> tiny initial means
    assert a.abs().max() < 0.2


def test_aliased_infos_behaves_like_the_reference_list():
    """From 2**16 envs on `step_wait` returns `AliasedInfos` instead of `[info] * N` (the N pointers alone cost ~2 ms per step
    at N = 2**20): everything SB3 / VecMonitor do with `infos` must give what the reference's aliased list gives."""
    from optimal_quad_control_rl_b200.envs import AliasedInfos
    n = 1000
    info = {"terminal_observation": np.arange(3), "TimeLimit.truncated": True}
    ref, got = [info] * n, AliasedInfos(info, n)
    assert len(got) == n and got[0] is info and got[n - 1] is info and got[-1] is info
    assert all(g is r for g, r in zip(got, ref)) and sum(1 for _ in got) == n
    assert list(got[:]) == ref and got[10:20] == ref[10:20] and got[::250] == ref[::250] and got == ref
    with pytest.raises(IndexError):
        got[n]
    new_infos = list(got[:])                        # VecMonitor.step_wait: `new_infos = list(infos[:])`
    new_infos[3] = got[3].copy()
    new_infos[3]["episode"] = {"r": 1.0}
    assert "episode" not in info and new_infos[4] is info
    for idx, done in enumerate([True] * n):         # SB3 collect_rollouts: bootstrap read
        assert got[idx].get("terminal_observation") is not None and got[idx].get("TimeLimit.truncated", False)
    got[5]["x"] = 1                                 # writes alias, like the reference's one dict
    assert got[900]["x"] == 1
