"""Rows f1 + f2: the fused closed-loop rollout kernel (policy on the tensor cores + env step, quads resident in
registers for all steps, ONE launch) against the unfused path (2 * steps launches of policy_kernel / step_kernel).
Reset draws and exploration noise are keyed by (seed, env, launch epoch + t), so the two must agree bit for bit --
buffers, final simulator state, device totals, and whatever is collected afterwards."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("obs", "actions", "raw_actions", "rewards", "dones")


def make(variant, n, tracks, seed=2, max_steps=12, ga=1):
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["indi" if variant == "indi" else "e2e"]
    if variant == "e2e":
        env = Q.Quadcopter3DGates(n, gp, gy, sp, gates_ahead=ga, reset_rng="device", seed=seed)
        env.disturbance_ranges = Q.training_disturbance_ranges()
    else:
        env = Q.Quadcopter3DGatesINDI(n, gp, gy, sp, gates_ahead=ga, reset_rng="device", seed=seed)
    env.max_steps = max_steps  # time-limit resets inside the rollout
    rng = np.random.default_rng(7)
    dims = [env.state_len, 120, 120, 120, 4]
    w = [rng.normal(0, 1.0 / np.sqrt(i), (o, i)).astype(np.float32) for i, o in zip(dims[:-1], dims[1:])]
    b = [rng.normal(0, 0.1, o).astype(np.float32) for o in dims[1:]]
    pol = Q.MlpPolicy(w, b, std=[0.6, 0.5, 0.7, 0.4], seed=8)
    return env, pol


def state_of(env):
    s = [env.world_states, env.target_gates, env.step_counts]
    if env._VARIANT == "e2e":
        s.append(env.disturbances)
    return s


@pytest.mark.parametrize("variant,n,steps,ga", [("e2e", 3000, 25, 1), ("e2e", 128, 1, 1), ("e2e", 1, 9, 1), ("e2e", 4133, 40, 2),
                                                ("e2e", 2048, 30, 0), ("indi", 4096, 25, 1), ("indi", 1001, 30, 1),
                                                ("indi", 77, 5, 2), ("e2e", 160000, 6, 1)])
@pytest.mark.parametrize("deterministic", [False, True])
def test_fused_rollout_equals_unfused(variant, n, steps, ga, deterministic, tracks):
    import torch
    outs = []
    for fused in (True, False):
        env, pol = make(variant, n, tracks, ga=ga)
        assert env._lib.qs_rollout_fused_supported(env._h, pol._h) == 1
        env.enable_stats(True)
        obs0 = env.reset_tensor().clone()
        l0 = env.launch_count
        r = env.rollout(pol, steps, deterministic=deterministic, fused=fused)
        torch.cuda.synchronize()
        assert env.launch_count - l0 == (1 if fused else steps)
        assert torch.equal(r["obs"][0], obs0)
        first = {k: r[k].clone() for k in KEYS}
        st = env.stats()
        mid = state_of(env)
        r2 = env.rollout(pol, 3, deterministic=deterministic, fused=fused)  # epochs advanced by `steps` on both paths
        torch.cuda.synchronize()
        assert torch.equal(r2["obs"][0], first["obs"][steps])
        a_next = pol.forward(r2["obs"][3].contiguous(), deterministic=deterministic).clone()  # policy epoch too
        outs.append((first, st, mid, {k: r2[k].clone() for k in KEYS}, state_of(env), a_next))
        env.close(); pol.close()
    (fa, sa, ma, ra, ea, na), (fb, sb, mb, rb, eb, nb) = outs
    for k in KEYS:
        assert torch.equal(fa[k], fb[k]), (k, (fa[k] != fb[k]).sum().item(), fa[k].numel())
        assert torch.equal(ra[k], rb[k]), ("second rollout", k)
    for x, y in zip(ma + ea, mb + eb):
        np.testing.assert_array_equal(x, y)
    assert torch.equal(na, nb)
    for k in ("env_steps", "dones", "truncated", "gates_passed", "gate_collisions", "ground_collisions", "out_of_bounds"):
        assert sa[k] == sb[k], k
    np.testing.assert_allclose(sa["reward_sum"], sb["reward_sum"], rtol=1e-6)  # f32 partial sums in a different order
    assert sa["env_steps"] == n * steps
    if steps >= 12:
        assert sa["dones"] >= n  # every env timed out at least once: the fused reset path ran inside the kernel


def test_fused_rollout_matches_oracle_step_for_step(tracks):
    """Sanity check against the CPU oracle, FREE-RUNNING (the kernel never exposes intermediate world states): the
    oracle replays the rollout's actions from the same initial state; envs are followed until their first reset
    (device RNG afterwards).  Free-running float32 trajectories drift, so the gate here is loose
    (1e-3 over 12 steps; measured ~1e-6) -- the parity gate proper is bit-equality with the unfused path above, which
    tests/test_gpu_parity.py pins to the oracle and the golden vectors at 1e-5."""
    import torch
    from oracle import c_oracle as O
    import optimal_quad_control_rl_b200 as Q
    n, steps = 512, 12
    env, pol = make("e2e", n, tracks, max_steps=1200)
    gp, gy, sp = tracks["e2e"]
    ora = O.OracleEnv("e2e", n, gp, gy, sp, gates_ahead=1)
    ora.disturbance_ranges = Q.training_disturbance_ranges()
    env.reset_tensor()
    ws, tg, sc, dist = state_of(env)
    r = env.rollout(pol, steps, fused=True)
    torch.cuda.synchronize()
    obs, act = r["obs"].cpu().numpy(), r["actions"].cpu().numpy()
    rew, done = r["rewards"].cpu().numpy(), r["dones"].cpu().numpy().astype(bool)
    ora.force(ws, tg, sc, dist)
    ora.update_states()
    alive = np.ones(n, bool)
    worst = 0.0
    for t in range(steps):
        np.random.seed(t)
        o, rw, dn, _ = ora.step(act[t])
        assert (dn[alive] != done[t][alive]).mean() <= 0.01  # a flag may flip on a 1-ulp difference when free-running
        alive &= dn == done[t]
        keep = alive & ~dn
        err = lambda a, b: float((np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1)).max()) if a.size else 0.0
        worst = max(worst, err(rew[t][alive], rw[alive]), err(obs[t + 1][keep], o[keep]))
        alive = keep  # after a reset the two RNGs differ: stop following that env
    print("fused rollout vs oracle, free-running 12 steps: worst scaled error", worst)
    assert alive.sum() > n // 2
    assert worst <= 1e-3, worst
