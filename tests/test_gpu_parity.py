"""Parity of the CUDA path (through the C ABI / the VecEnv mirror) against
  (1) golden vectors frozen from the unmodified reference cells, and
  (2) the CPU oracle on seeded inputs at sizes up to BASELINE.json's 2**20.
Floats: |a-b| <= 1e-5*max(|b|,1) (north_star's 1e-5 float32 gate).  done / gate index / step
counters / positions / reset values: bit-exact."""
import numpy as np
import pytest

from conftest import assert_close, golden, scaled_err

pytestmark = pytest.mark.gpu

VARIANTS = ("e2e", "indi")


def make_env(variant, n, tracks, ga=1, pic=False, ranges=None, **kw):
    import optimal_quad_control_rl_b200 as Q
    cls = Q.Quadcopter3DGates if variant == "e2e" else Q.Quadcopter3DGatesINDI
    gp, gy, sp = tracks[variant]
    env = cls(n, gp, gy, sp, gates_ahead=ga, pause_if_collision=pic, **kw)
    if ranges is not None:
        env.disturbance_ranges = ranges
    return env


def force(env, ws, tg, sc, dist=None):
    env.world_states = ws
    env.target_gates = tg
    env.step_counts = sc
    if dist is not None:
        env.disturbances = dist
    return env.update_states()


def test_library_is_the_cuda_one():
    import optimal_quad_control_rl_b200._lib as L
    lib = L.load()
    assert b"sm_100a" in lib.qs_version()


@pytest.mark.parametrize("variant", VARIANTS)
def test_state_roundtrip_is_exact(variant, tracks):
    g = golden(f"{variant}_single_step")
    env = make_env(variant, len(g["in_ws"]), tracks)
    force(env, g["in_ws"], g["in_tg"], g["in_sc"], g.get("in_dist"))
    np.testing.assert_array_equal(env.world_states, g["in_ws"])
    np.testing.assert_array_equal(env.target_gates, g["in_tg"])
    np.testing.assert_array_equal(env.step_counts, g["in_sc"])
    if variant == "e2e":
        np.testing.assert_array_equal(env.disturbances, g["in_dist"])


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("first,count", [(5, 1), (31, 3), (0, 7), (100, 33)])
def test_state_subrange_get_set_all_fields_at_once(variant, first, count, tracks):
    """qs_get_state / qs_set_state on an odd slice with every field requested in ONE call (13 floats x an odd count
    once misaligned the int64 staging arrays)."""
    import ctypes as C
    from optimal_quad_control_rl_b200 import _lib as L
    g = golden(f"{variant}_single_step")
    n = 160
    env = make_env(variant, n, tracks)
    force(env, g["in_ws"][:n], g["in_tg"][:n], g["in_sc"][:n], g["in_dist"][:n] if variant == "e2e" else None)
    ns = g["in_ws"].shape[1]
    ws, tg, sc = np.empty((count, ns), np.float32), np.empty(count, np.int64), np.empty(count, np.int64)
    dist = np.empty((count, 6), np.float32) if variant == "e2e" else None
    fp = lambda a: a.ctypes.data_as(L._fp) if a is not None else None
    ip = lambda a: a.ctypes.data_as(L._i64p)
    env._call("qs_get_state", first, count, fp(ws), fp(dist), ip(tg), ip(sc))
    sl = slice(first, first + count)
    np.testing.assert_array_equal(ws, g["in_ws"][sl])
    np.testing.assert_array_equal(tg, g["in_tg"][sl] % env.num_gates)
    np.testing.assert_array_equal(sc, g["in_sc"][sl])
    if dist is not None:
        np.testing.assert_array_equal(dist, g["in_dist"][sl])
    ws2, sc2 = env._get_rows(first, count)
    np.testing.assert_array_equal(ws2, ws); np.testing.assert_array_equal(sc2, sc)
    # write the slice back shifted by one env and read the whole state: only that slice changed
    before = env.world_states
    env._call("qs_set_state", first, count, fp(np.ascontiguousarray(ws[::-1])), fp(dist), ip(tg), ip(sc))
    after = env.world_states
    np.testing.assert_array_equal(after[sl], ws[::-1])
    mask = np.ones(n, bool); mask[sl] = False
    np.testing.assert_array_equal(after[mask], before[mask])
    with pytest.raises(IndexError):
        env._get_rows(n - 1, 2)


@pytest.mark.parametrize("variant", VARIANTS)
def test_gate_tables_match_reference(variant, tracks):
    k = golden("kat")
    env = make_env(variant, 4, tracks)
    np.testing.assert_array_equal(env.gate_pos_rel, k[f"{variant}_gate_pos_rel"])
    np.testing.assert_array_equal(env.gate_yaw_rel, k[f"{variant}_gate_yaw_rel"])


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("branch", ["pic", "nrm", "pau"])
def test_single_step_vs_reference_golden(variant, branch, tracks):
    """8192 teacher-forced states (random box + gate-plane + ground/bounds/time-limit edge cases), all three
    step_wait branches, against the reference's own outputs."""
    g = golden(f"{variant}_single_step")
    n = len(g["in_ws"])
    env = make_env(variant, n, tracks, pic=(branch == "pic"), ranges=g.get("disturbance_ranges"))
    obs0 = force(env, g["in_ws"], g["in_tg"], g["in_sc"], g.get("in_dist"))
    assert_close(obs0, g["in_obs"], "obs before")
    env.pause = branch == "pau"
    if branch == "nrm":
        np.random.seed(int(g["nrm_seed"]))
    prev = env.states
    obs, rew, done, infos = env.step(g["in_act"])
    assert obs.dtype == np.float32 and rew.dtype == np.float32 and done.dtype == bool
    np.testing.assert_array_equal(done, g[f"{branch}_done"])
    np.testing.assert_array_equal(env.target_gates, g[f"{branch}_tg"])
    np.testing.assert_array_equal(env.step_counts, g[f"{branch}_sc"])
    assert_close(rew, g[f"{branch}_rew"], "reward")
    ws = env.world_states
    assert_close(ws, g[f"{branch}_ws"], "world_states")
    assert_close(obs, g[f"{branch}_obs"], "obs")
    if branch == "pic":  # positions of advanced envs: pos + dt*vel, bit-exact
        adv = ~done
        np.testing.assert_array_equal(ws[adv, 0:3], g["pic_ws"][adv, 0:3])
    if branch == "pau":
        assert obs is prev and not done.any()
    if branch == "nrm":
        d = g["nrm_done"]
        np.testing.assert_array_equal(ws[d], g["nrm_ws"][d])  # host-replayed np.random draws: bit-identical
        if variant == "e2e":
            np.testing.assert_array_equal(env.disturbances, g["nrm_dist"])
        assert bool(infos[0].get("TimeLimit.truncated", False)) == bool(g["nrm_info_truncated"])
        assert_close(infos[0]["terminal_observation"], g["nrm_info_terminal_obs"], "terminal_observation")
        assert all(i is infos[0] for i in infos)
    e = scaled_err(ws, g[f"{branch}_ws"]).max(axis=0)
    print(f"{variant}/{branch}: worst scaled err per state column {np.array2string(e, precision=2)}")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("ga", [0, 2])
def test_obs_layout_other_gates_ahead(variant, ga, tracks):
    g = golden(f"{variant}_obs_ga{ga}")
    env = make_env(variant, len(g["in_ws"]), tracks, ga=ga, pic=True, ranges=g.get("disturbance_ranges"))
    obs0 = force(env, g["in_ws"], g["in_tg"], g["in_sc"], g.get("in_dist"))
    assert obs0.shape == g["in_obs"].shape
    assert_close(obs0, g["in_obs"], "obs before")
    obs, rew, done, _ = env.step(g["in_act"])
    np.testing.assert_array_equal(done, g["pic_done"])
    np.testing.assert_array_equal(env.target_gates, g["pic_tg"])
    assert_close(obs, g["pic_obs"], "obs")
    assert_close(rew, g["pic_rew"], "reward")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", ["traj_n1", "traj_n16"])
def test_trajectory_teacher_forced(variant, name, tracks):
    """BASELINE config 1 (N=1, np.random.seed(0), 1000 steps) + an N=16 rollout, re-synced to the reference's
    state every step; resets replay the reference's np.random order."""
    g = golden(f"{variant}_{name}")
    steps, n = g["actions"].shape[:2]
    env = make_env(variant, n, tracks, ranges=g.get("disturbance_ranges"))
    env.max_steps = int(g["max_steps"])
    np.random.seed(int(g["np_seed"]))
    obs = env.reset()
    np.testing.assert_array_equal(env.world_states, g["ws"][0])
    assert_close(obs, g["obs"][0], "reset obs")
    worst = 0.0
    for t in range(steps):
        force(env, g["ws"][t], g["tg"][t], g["sc"][t], g["dist"][t] if "dist" in g else None)
        obs, rew, done, infos = env.step(g["actions"][t])
        np.testing.assert_array_equal(done, g["done"][t], err_msg=f"done @ {t}")
        np.testing.assert_array_equal(env.target_gates, g["tg"][t + 1])
        np.testing.assert_array_equal(env.step_counts, g["sc"][t + 1])
        ws = env.world_states
        assert_close(ws, g["ws"][t + 1], f"ws @ {t}")
        assert_close(obs, g["obs"][t + 1], f"obs @ {t}")
        assert_close(rew, g["rew"][t], f"rew @ {t}")
        worst = max(worst, scaled_err(ws, g["ws"][t + 1]).max())
        assert bool(infos[0].get("TimeLimit.truncated", False)) == bool(g["info_truncated"][t])
        if done.any():
            assert_close(infos[0]["terminal_observation"], g["info_terminal_obs"][t], "terminal_observation")
            np.testing.assert_array_equal(ws[done], g["ws"][t + 1][done])
    print(f"{variant}/{name}: worst scaled state error over {steps} teacher-forced steps {worst:.2e}")


@pytest.mark.parametrize("variant", VARIANTS)
def test_trajectory_free_running_report(variant, tracks):
    g = golden(f"{variant}_traj_n1")
    env = make_env(variant, 1, tracks, ranges=g.get("disturbance_ranges"))
    np.random.seed(int(g["np_seed"]))
    env.reset()
    first_bad = None
    for t in range(g["actions"].shape[0]):
        env.step(g["actions"][t])
        if first_bad is None and scaled_err(env.world_states, g["ws"][t + 1]).max() > 1e-5:
            first_bad = t
    print(f"{variant}: free-running first step over 1e-5: {first_bad}")
    assert first_bad is None or first_bad >= 20


# ------------------------------------------------------------------------------------------------ vs the CPU oracle
def oracle_inputs(variant, n, tracks, seed):
    rng = np.random.default_rng(seed)
    ns = 16 if variant == "e2e" else 13
    ng = len(tracks[variant][1])
    ws = np.zeros((n, ns), np.float32)
    ws[:, 0:2] = rng.uniform(-4, 4, (n, 2)); ws[:, 2] = rng.uniform(-3, 0.02, n)
    ws[:, 3:6] = rng.uniform(-8, 8, (n, 3)); ws[:, 6:8] = rng.uniform(-1.2, 1.2, (n, 2))
    ws[:, 8] = rng.uniform(-3 * np.pi, 3 * np.pi, n); ws[:, 9:12] = rng.uniform(-6, 6, (n, 3))
    ws[:, 12:] = rng.uniform(-1, 1, (n, ns - 12))
    # a quarter of the envs sit just in front of their target gate, flying through it
    k = n // 4
    gp, gy, _ = tracks[variant]
    tg = rng.integers(0, ng, n).astype(np.int64)
    nrm = np.stack([np.cos(gy[tg[:k]]), np.sin(gy[tg[:k]])], 1)
    tan = np.stack([-nrm[:, 1], nrm[:, 0]], 1)
    ws[:k, 0:2] = gp[tg[:k], 0:2] - rng.uniform(0, 0.06, (k, 1)) * nrm + rng.uniform(-0.7, 0.7, (k, 1)) * tan
    ws[:k, 2] = gp[tg[:k], 2] + rng.uniform(-0.7, 0.7, k)
    ws[:k, 3:5] = rng.uniform(1, 12, (k, 1)) * nrm
    sc = rng.integers(0, 1203, n).astype(np.int64)
    act = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    dist = rng.uniform([-.03, -.03, -.01, -.1, -.1, -.5], [.03, .03, .01, .1, .1, .5], (n, 6)).astype(np.float32)
    return ws, tg, sc, act, (dist if variant == "e2e" else None)


@pytest.mark.parametrize("variant,n", [("e2e", 1), ("e2e", 127), ("e2e", 4096), ("indi", 3), ("indi", 262144),
                                       ("e2e", 1 << 20), ("indi", (1 << 20) + 77)])
def test_step_vs_oracle(variant, n, tracks):
    """Ragged sizes (tail tile, unaligned bulk store), BASELINE's N=4096 / 262144 / 2**20."""
    from oracle import c_oracle as O
    ws, tg, sc, act, dist = oracle_inputs(variant, n, tracks, seed=n)
    ranges = np.array([[-.03, .03], [-.03, .03], [-.01, .01], [-.1, .1], [-.1, .1], [-.5, .5]]) if variant == "e2e" else None
    gp, gy, sp = tracks[variant]
    ora = O.OracleEnv(variant, n, gp, gy, sp, gates_ahead=1, pause_if_collision=True)
    env = make_env(variant, n, tracks, pic=True, ranges=ranges)
    if ranges is not None:
        ora.disturbance_ranges = ranges
    ora.force(ws, tg, sc, dist)
    obs0 = force(env, ws, tg, sc, dist)
    assert_close(obs0, ora.states, "obs before")
    o_obs, o_rew, o_done, _ = ora.step(act)
    obs, rew, done, _ = env.step(act)
    np.testing.assert_array_equal(done, o_done)
    np.testing.assert_array_equal(env.last_flags, ora.last_flags)
    np.testing.assert_array_equal(env.target_gates, ora.target_gates)
    np.testing.assert_array_equal(env.step_counts, ora.step_counts)
    ws1 = env.world_states
    np.testing.assert_array_equal(ws1[:, 0:3], ora.world_states[:, 0:3])
    assert_close(ws1, ora.world_states, "world_states")
    assert_close(obs, o_obs, "obs")
    assert_close(rew, o_rew, "reward")
    np.testing.assert_array_equal(rew, o_rew)  # distances and differences use the same rounded operations
    assert done.sum() > 0 or n < 100


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_disturbance_obs_narrow_offcentre_ranges(dtype, tracks):
    """2*(d-lo)/(hi-lo)-1 for ranges like [10, 10.001] (`3D quad race.ipynb:414-448`): the kernel subtracts first, like the
    reference, in the precision of the ranges array (a folded d*scale+offset form is off by ~1e-3 here)."""
    g = golden("e2e_obs_offcentre")
    dt = np.float64 if dtype == "f64" else np.float32
    env = make_env("e2e", len(g["in_ws"]), tracks, ranges=g["disturbance_ranges"].astype(dt))
    obs = force(env, g["in_ws"], g["in_tg"], g["in_sc"], g["in_dist"])
    assert_close(obs, g[f"obs_{dtype}"], "obs")


@pytest.mark.parametrize("variant", VARIANTS)
def test_teacher_set_2p20_digest(variant, tracks):
    """SURVEY section 8(d)'s full-size teacher-forced set through the CUDA path: 2**20 states incl. the near-threshold
    generators; done / target_gate / step_count hash to the REFERENCE's own arrays (digest frozen from the unmodified
    notebook cells by oracle/make_golden.py --teacher-only), every 64th float row within the 1e-5 gate of the reference's."""
    from oracle import make_golden as MG
    g = golden(f"{variant}_teacher_2p20_digest")
    n, k = int(g["n"]), int(g["stride"])
    ws, tg, sc, act, dist = MG.teacher_set_inputs(variant, tracks[variant], int(g["seed"]), n)
    assert MG.sha(np.concatenate([ws.ravel(), act.ravel()])) == str(g["sha_inputs"]), "input generator drifted"
    env = make_env(variant, n, tracks, pic=True, ranges=MG.R.training_disturbance_ranges() if variant == "e2e" else None)
    obs0 = force(env, ws, tg, sc, dist)
    assert_close(obs0[::k], g["obs0"], "obs before")
    obs, rew, done, _ = env.step(act)
    assert MG.sha(done.astype(np.uint8)) == str(g["sha_done"])
    assert MG.sha(env.target_gates.astype(np.int64)) == str(g["sha_tg"])
    assert MG.sha(env.step_counts.astype(np.int64)) == str(g["sha_sc"])
    w1 = env.world_states
    np.testing.assert_array_equal(w1[::k, 0:3], g["ws"][:, 0:3])
    assert_close(w1[::k], g["ws"], "world_states")
    assert_close(obs[::k], g["obs"], "obs")
    assert_close(rew[::k], g["rew"], "reward")
    print(variant, "max scaled err CUDA vs reference (every 64th of 2**20):", float(scaled_err(w1[::k], g["ws"]).max()))


# ------------------------------------------------------------------------------------------------ size-independent properties
@pytest.mark.parametrize("variant", VARIANTS)
def test_properties_full_size(variant, tracks):
    import torch
    n = 1 << 20
    env = make_env(variant, n, tracks, reset_rng="device", seed=5)
    env.disturbance_ranges = np.array([[-.03, .03], [-.03, .03], [-.01, .01], [0, 0], [0, 0], [-.5, .5]])
    obs = env.reset_tensor().clone()
    ws0 = env.world_states
    sp = env.start_pos
    # reset distribution (`:455-474`): box bounds and means
    assert (np.abs(ws0[:, 0:3] - sp) <= 0.5 + 1e-6).all() and (np.abs(ws0[:, 3:6]) <= 0.5).all()
    assert (np.abs(ws0[:, 6:8]) <= np.pi / 9 + 1e-6).all() and (np.abs(ws0[:, 8]) <= np.pi + 1e-6).all()
    assert np.abs(ws0.mean(0) - np.r_[sp, np.zeros(ws0.shape[1] - 3)]).max() < 0.01
    assert np.unique(ws0[:, 0]).size > n // 20
    # observe is a pure function of the state (idempotent)
    o1 = torch.from_numpy(env.update_states()).cuda()
    assert torch.equal(o1, obs)
    # pause: state, observation untouched; counters advance
    env.pause = True
    a = torch.rand((n, 4), device="cuda") * 2 - 1
    _, rew, done, _ = env.step_tensor(a)
    assert not done.any().item()
    np.testing.assert_array_equal(env.world_states, ws0)
    assert (env.step_counts == 1).all()
    env.pause = False
    # a normal step resets exactly the done envs: their counters are 0, everyone else's are 2
    env.enable_stats(True)
    _, rew, done, flags = env.step_tensor(a)
    sc = env.step_counts
    d = done.cpu().numpy().astype(bool)
    assert (sc[d] == 0).all() and (sc[~d] == 2).all()
    st = env.stats()
    assert st["env_steps"] == n and st["dones"] == int(d.sum())
    np.testing.assert_allclose(st["reward_sum"], rew.double().sum().item(), rtol=1e-9)


@pytest.mark.parametrize("variant", VARIANTS)
def test_device_rng_is_shard_invariant(variant, tracks):
    """Philox is keyed by the GLOBAL env index: 1 handle of 4096 == 2 handles of 2048 with offsets."""
    n = 4096
    full = make_env(variant, n, tracks, reset_rng="device", seed=9)
    a = make_env(variant, n // 2, tracks, reset_rng="device", seed=9, env_offset=0)
    b = make_env(variant, n // 2, tracks, reset_rng="device", seed=9, env_offset=n // 2)
    of = full.reset()
    oa, ob = a.reset(), b.reset()
    np.testing.assert_array_equal(of, np.concatenate([oa, ob]))
    act = np.random.default_rng(0).uniform(-1, 1, (n, 4)).astype(np.float32)
    for env in (full, a, b):
        env.max_steps = 3  # force time-limit resets through the fused path
    for _ in range(4):
        of = full.step(act)[0]
        oa, ob = a.step(act[:n // 2])[0], b.step(act[n // 2:])[0]
        np.testing.assert_array_equal(of, np.concatenate([oa, ob]))


@pytest.mark.parametrize("variant", VARIANTS)
def test_device_rng_advances_every_launch_also_under_graph_replay(variant, tracks):
    """The RNG's launch epoch lives in device memory: every launch (also the SAME captured launch replayed by a
    CUDA graph) draws fresh reset values, and qs_seed rewinds the stream."""
    import torch
    n = 1000  # ragged tail tile, several warps with > 32/ND resets (per-lane path) since every env times out
    env = make_env(variant, n, tracks, reset_rng="device", seed=11)
    env.max_steps = 1
    env.reset_tensor()
    a = torch.zeros((n, 4), device="cuda")
    seen = []
    for _ in range(3):
        env.step_tensor(a)
        seen.append(env.world_states.copy())
    assert not np.array_equal(seen[0], seen[1]) and not np.array_equal(seen[1], seen[2])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.step_tensor(a)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.step_tensor(a)
    reps = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        reps.append(env.world_states.copy())
    assert not np.array_equal(reps[0], reps[1]) and not np.array_equal(reps[1], reps[2])
    sp = env.start_pos
    for ws in reps:  # still the reset distribution
        assert (np.abs(ws[:, 0:3] - sp) <= 0.5 + 1e-6).all() and np.unique(ws[:, 0]).size > n // 2
    # same seed, same call sequence => same values
    env._call("qs_seed", 11)
    env.reset_tensor()
    env.step_tensor(a)
    np.testing.assert_array_equal(env.world_states, seen[0])


@pytest.mark.parametrize("variant", VARIANTS)
def test_few_resets_per_warp_use_cooperative_draw_and_match_per_lane_draw(variant, tracks):
    """Warp-cooperative reset (<= 32/ND terminating lanes per warp) and the per-lane fallback (mass termination)
    are the same function of (seed, env, epoch): force both on the same envs and compare."""
    n = 4096
    ws, tg, sc, act, dist = oracle_inputs(variant, n, tracks, seed=3)
    outs = []
    for dense in (False, True):
        env = make_env(variant, n, tracks, reset_rng="device", seed=21)
        env.max_steps = 50
        force(env, ws, tg, np.full(n, 10), dist)
        sc2 = np.full(n, 10)
        pick = np.arange(n) % 32 == 5 if not dense else np.ones(n, bool)   # 1 lane per warp | every lane
        sc2[pick] = 49                                                       # those time out in this step
        env.step_counts = sc2
        env.update_states()
        import torch
        _, _, done, _ = env.step_tensor(torch.from_numpy(act).cuda())
        d = done.cpu().numpy().astype(bool)
        assert d[pick].all()
        outs.append((env.world_states, d))
    sel = np.arange(n) % 32 == 5
    np.testing.assert_array_equal(outs[0][0][sel], outs[1][0][sel])


@pytest.mark.parametrize("n,first_div", [(5000, None), (100003, None), (600001, "1"), (600001, "4"), (600001, "2")])
def test_c_abi_host_step_matches_device_step(n, first_div, tracks, monkeypatch):
    """qs_step_host (HOST buffers in/out; chunk-pipelined above 32768 envs, equal chunks or a shorter first one) ==
    qs_step on device buffers, bit for bit, fused resets included."""
    import ctypes as C
    import optimal_quad_control_rl_b200._lib as L
    if first_div is not None:  # read by the library at an env's first host-buffer call
        monkeypatch.setenv("QS_HOST_FIRST_DIV", first_div)
    e1 = make_env("e2e", n, tracks, reset_rng="device", seed=3)
    e2 = make_env("e2e", n, tracks, reset_rng="device", seed=3)
    o1 = e1.reset()
    lib = e2._lib
    o2 = np.empty_like(o1)
    e2._call("qs_reset_all_host", o2.ctypes.data_as(L._vp))
    np.testing.assert_array_equal(o1, o2)
    act = np.random.default_rng(1).uniform(-1, 1, (n, 4)).astype(np.float32)
    obs, rew, done, _ = e1.step(act)
    rew2 = np.empty(n, np.float32); done2 = np.empty(n, np.uint8); fl2 = np.empty(n, np.uint8)
    e2._call("qs_step_host", act.ctypes.data_as(L._vp), o2.ctypes.data_as(L._vp), rew2.ctypes.data_as(L._vp),
             done2.ctypes.data_as(L._vp), fl2.ctypes.data_as(L._vp), L.MODE_NORMAL, L.RESET_DEVICE)
    np.testing.assert_array_equal(obs, o2)
    np.testing.assert_array_equal(rew, rew2)
    np.testing.assert_array_equal(done, done2.astype(bool))
    np.testing.assert_array_equal(e1.last_flags, fl2)
    np.testing.assert_array_equal(e1.world_states, e2.world_states)
    # a second step: the RNG epoch advanced exactly once in both
    for e in (e1, e2):
        e.max_steps = 2
    e2._push_config()
    obs, rew, done, _ = e1.step(act)
    e2._call("qs_step_host", act.ctypes.data_as(L._vp), o2.ctypes.data_as(L._vp), rew2.ctypes.data_as(L._vp),
             done2.ctypes.data_as(L._vp), fl2.ctypes.data_as(L._vp), L.MODE_NORMAL, L.RESET_DEVICE)
    assert done.all()
    np.testing.assert_array_equal(obs, o2)
    np.testing.assert_array_equal(e1.world_states, e2.world_states)


def test_errors_are_reported_not_raised_across_abi(tracks):
    import ctypes as C
    import optimal_quad_control_rl_b200._lib as L
    lib = L.load()
    h = L._vp()
    gp = np.zeros((2, 3), np.float32); gy = np.zeros(2, np.float32); sp = np.zeros(3, np.float32)
    fp = lambda a: a.ctypes.data_as(L._fp)
    assert lib.qs_create(C.byref(h), 7, 16, 2, fp(gp), fp(gy), fp(sp), 1, 0, None) == -1
    assert b"variant" in lib.qs_last_error(None)
    assert lib.qs_create(C.byref(h), L.E2E, 16, 2, fp(gp), fp(gy), fp(sp), 1, 0, None) == 0
    # stepping an E2E env before its residual weights are set is a state error, not a crash
    assert lib.qs_step(h, L._vp(16), L._vp(16), L._vp(16), L._vp(16), None, 0, 0) == -3
    assert b"weights" in lib.qs_last_error(h)
    assert lib.qs_set_state(h, 10, 100, None, None, None, None) == -1
    lib.qs_destroy(h)


@pytest.mark.parametrize("variant,n", [("e2e", 5000), ("indi", 40000)])
def test_numpy_step_uses_pinned_host_path_and_matches_tensor_path(variant, n, tracks):
    """reset_rng="device": env.step(NumPy) is ONE qs_step_host call into a ring of pinned arrays; it must return what
    the zero-copy tensor path returns for the same seed, keep earlier outputs intact for obs_buffers steps, and hand a
    consistent observation tensor to a rollout that follows."""
    import torch
    ranges = None
    if variant == "e2e":
        import optimal_quad_control_rl_b200 as Q
        ranges = Q.training_disturbance_ranges()
    a = make_env(variant, n, tracks, ranges=ranges, reset_rng="device", seed=5, obs_buffers=3)
    b = make_env(variant, n, tracks, ranges=ranges, reset_rng="device", seed=5, obs_buffers=3)
    a.max_steps = b.max_steps = 7
    oa = a.reset()
    ob = b.reset_tensor().cpu().numpy()
    np.testing.assert_array_equal(oa, ob)
    rng = np.random.default_rng(0)
    prev = None
    for t in range(12):
        act = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        l0 = a.launch_count
        obs, rew, done, infos = a.step(act if t % 2 else act.astype(np.float64))  # float64 actions are cast
        tobs, trew, tdone, tfl = b.step_tensor(torch.from_numpy(act).cuda())
        np.testing.assert_array_equal(obs, tobs.cpu().numpy())
        np.testing.assert_array_equal(rew, trew.cpu().numpy())
        np.testing.assert_array_equal(done, tdone.cpu().numpy().astype(bool))
        assert done.dtype == np.bool_ and obs.dtype == np.float32 and a.states is obs
        if prev is not None:  # the previous step's arrays are still what they were (SB3 reads them after env.step)
            np.testing.assert_array_equal(prev[0], prev[1])
        prev = (obs, obs.copy())
        if done.any():
            assert "terminal_observation" in infos[0]
        if t == 6:
            assert done.all() and infos[0].get("TimeLimit.truncated")
    np.testing.assert_array_equal(a.current_obs_tensor().cpu().numpy(), obs)
    np.testing.assert_array_equal(a.world_states, b.world_states)
    a.close(); b.close()
