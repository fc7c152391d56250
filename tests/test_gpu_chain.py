"""Chained step launches (StepParams::chain): consecutive steps captured into one CUDA graph depend on each other CTA
by CTA instead of grid by grid.  Whatever the overlap, the results must be the ones of the classic launch order, bit
for bit: same states, observations, rewards, flags, device RNG draws and reward statistics."""
import numpy as np
import pytest

from test_gpu_parity import VARIANTS, make_env

pytestmark = pytest.mark.gpu


def _run(variant, n, tracks, chain, graph_steps, replays, eager_between, monkeypatch):
    import torch
    monkeypatch.setenv("QS_CHAIN", "1" if chain else "0")
    env = make_env(variant, n, tracks, reset_rng="device", seed=5, obs_buffers=2)
    if variant == "e2e":
        import optimal_quad_control_rl_b200 as Q
        env.disturbance_ranges = Q.training_disturbance_ranges()
    env.max_steps = 7  # time-outs in every replay: the fused reset draws (keyed by the launch epoch) run all the time
    env.enable_stats()
    env.reset_tensor()
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = [torch.rand((n, 4), device="cuda", generator=gen) * 2 - 1 for _ in range(graph_steps)]
    outs = []
    if graph_steps:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            env.step_tensor(acts[0])  # warm-up launch outside the capture
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for a in acts:
                last = env.step_tensor(a)
        for r in range(replays):
            g.replay()
            got = [t.clone() for t in last]  # before the eager step recycles the ring slot
            if eager_between:
                env.step_tensor(acts[r % graph_steps])  # an eager launch between two replays
            torch.cuda.synchronize()
            outs.append(got + [torch.from_numpy(env.world_states.copy())])
    torch.cuda.synchronize()
    # every captured step but the first of the graph skipped the grid-wide wait -- or none did with QS_CHAIN=0
    assert env.chained_launch_count == ((graph_steps - 1) if chain else 0)
    st = env.stats()
    final = (env.world_states.copy(), env.target_gates.copy(), env.step_counts.copy())
    env.close()
    return outs, final, st


def _eager_reference(variant, n, tracks, graph_steps, replays, eager_between, monkeypatch):
    """The same call sequence, one classic launch per step (no graph, chaining compiled out of the decision)."""
    import torch
    monkeypatch.setenv("QS_CHAIN", "0")
    env = make_env(variant, n, tracks, reset_rng="device", seed=5, obs_buffers=2)
    if variant == "e2e":
        import optimal_quad_control_rl_b200 as Q
        env.disturbance_ranges = Q.training_disturbance_ranges()
    env.max_steps = 7
    env.enable_stats()
    env.reset_tensor()
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = [torch.rand((n, 4), device="cuda", generator=gen) * 2 - 1 for _ in range(graph_steps)]
    outs = []
    env.step_tensor(acts[0])
    for r in range(replays):
        for a in acts:
            last = env.step_tensor(a)
        last = [t.clone() for t in last]
        if eager_between:
            env.step_tensor(acts[r % graph_steps])
        torch.cuda.synchronize()
        outs.append(last + [torch.from_numpy(env.world_states.copy())])
    final = (env.world_states.copy(), env.target_gates.copy(), env.step_counts.copy())
    assert env.chained_launch_count == 0
    st = env.stats()
    env.close()
    return outs, final, st


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n,eager_between", [(300_000, False), (4096, False), (1000, True), (1 << 20, False)])
def test_chained_graph_replays_equal_classic_launches(variant, n, eager_between, tracks, monkeypatch):
    steps, replays = 6, 4
    ref_outs, ref_final, ref_st = _eager_reference(variant, n, tracks, steps, replays, eager_between, monkeypatch)
    for chain in (True, False):  # the graph with chained launches, and the same graph with classic ones
        outs, final, st = _run(variant, n, tracks, chain, steps, replays, eager_between, monkeypatch)
        for r, (a, b) in enumerate(zip(outs, ref_outs)):
            for k, (x, y) in enumerate(zip(a, b)):
                assert np.array_equal(x.cpu().numpy(), y.cpu().numpy()), (chain, "replay", r, "tensor", k)
        for x, y in zip(final, ref_final):
            np.testing.assert_array_equal(x, y)
        assert st == ref_st


@pytest.mark.parametrize("ga,pic", [(1, False), (0, False), (1, True)])
def test_two_envs_per_thread_kernel_equals_one_env_kernel(ga, pic, tracks, monkeypatch):
    """step_kernel_x2 (E2E, a thread steps two envs and shares the residual nets' weight fetches between them) against
    step_kernel<e2e>: every output of every step, the state and the statistics, bit for bit; ragged N, fused resets."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    n, steps = 300_007, 12
    res = []
    for x2 in ("1", "0"):
        monkeypatch.setenv("QS_STEP_X2", x2)
        env = make_env("e2e", n, tracks, ga=ga, pic=pic, reset_rng="device", seed=9)
        env.disturbance_ranges = Q.training_disturbance_ranges()
        env.max_steps = 5
        env.enable_stats()
        env.reset_tensor()
        gen = torch.Generator(device="cuda").manual_seed(4)
        outs = []
        for _ in range(steps):
            a = torch.rand((n, 4), device="cuda", generator=gen) * 2 - 1
            outs.append([t.clone() for t in env.step_tensor(a)])
        torch.cuda.synchronize()
        res.append((outs, env.world_states.copy(), env.target_gates.copy(), env.step_counts.copy(), env.disturbances.copy(),
                    env.stats()))
        env.close()
    (oa, wa, ta, sa, da, sta), (ob, wb, tb, sb, db, stb) = res
    for t, (x, y) in enumerate(zip(oa, ob)):
        for k, (p, q) in enumerate(zip(x, y)):
            assert torch.equal(p, q), ("step", t, "tensor", k)
    np.testing.assert_array_equal(wa, wb)
    np.testing.assert_array_equal(ta, tb)
    np.testing.assert_array_equal(sa, sb)
    np.testing.assert_array_equal(da, db)
    assert sta["dones"] > 0 and {k: v for k, v in sta.items() if k != "reward_sum"} == {k: v for k, v in stb.items() if k != "reward_sum"}
    assert abs(sta["reward_sum"] - stb["reward_sum"]) <= 1e-6 * max(1.0, abs(stb["reward_sum"]))  # other partition of the float sums


def test_a_foreign_kernel_between_two_captured_steps_breaks_the_chain(tracks, monkeypatch):
    """Chaining skips the grid-wide wait, so it is only allowed when the step depends on NOTHING but the previous
    chained step: any other node captured in between (here the kernel that produces the next actions) is waited for."""
    import torch
    monkeypatch.setenv("QS_CHAIN", "1")
    n = 8192
    env = make_env("e2e", n, tracks, reset_rng="device", seed=1)
    env.reset_tensor()
    a = torch.zeros((n, 4), device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.step_tensor(a)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.step_tensor(a)
        env.step_tensor(a)          # chained: depends on the previous step only
        a.add_(0.125)               # a foreign kernel writes the actions
        env.step_tensor(a)          # not chained
        env.step_tensor(a)          # chained again
    assert env.chained_launch_count == 2
    g.replay()
    torch.cuda.synchronize()
    ref = make_env("e2e", n, tracks, reset_rng="device", seed=1)
    ref.reset_tensor()
    b = torch.zeros((n, 4), device="cuda")
    for k in range(5):
        if k == 3:
            b.add_(0.125)
        ref.step_tensor(b)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(env.world_states, ref.world_states)


@pytest.mark.parametrize("variant", VARIANTS)
def test_step_graph_helper_equals_step_by_step(variant, tracks):
    """env.step_graph(actions (T,N,4)): one chained CUDA graph; every step's outputs equal T step_tensor calls."""
    import torch
    n, T = 50_000, 9
    gen = torch.Generator(device="cuda").manual_seed(8)
    acts = torch.rand((T, n, 4), device="cuda", generator=gen) * 2 - 1
    a = make_env(variant, n, tracks, reset_rng="device", seed=2)
    b = make_env(variant, n, tracks, reset_rng="device", seed=2)
    for env in (a, b):
        env.max_steps = 4
        env.reset_tensor()
    graph, out = a.step_graph(acts)
    assert a.chained_launch_count == T - 1
    for rep in range(2):
        graph.replay()
        torch.cuda.synchronize()
        for t in range(T):
            o, r, d, f = b.step_tensor(acts[t])
            assert torch.equal(out["obs"][t], o) and torch.equal(out["rewards"][t], r), (rep, t)
            assert torch.equal(out["dones"][t], d) and torch.equal(out["flags"][t], f), (rep, t)
    np.testing.assert_array_equal(a.world_states, b.world_states)
    np.testing.assert_array_equal(a.current_obs_tensor().cpu().numpy(), b.current_obs_tensor().cpu().numpy())
