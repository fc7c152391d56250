"""Row f3 (CPU): the flight-controller C export.  The generated sources are compiled with gcc and checked
  * bit for bit against the reference's own shipped ``c_code`` (compiled where it lies into oracle/_ref) for the same
    weights (K4) and track -- network, controller, gate switching, tables;
  * against the oracle env's observation fed through the oracle's restated network (both model variants);
  * against the reference's known-answer vector K1 for the residual models (`3D quad race.ipynb:215,219`)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden

REF_FLAGS = ("-O2", "-ffp-contract=off")  # what oracle/Makefile builds the reference's c_code with


def k4():
    z = golden("policy_k4")
    n = len(z["dims"]) - 1
    return [z[f"W{l}"] for l in range(n)], [z[f"b{l}"] for l in range(n)], z["std"], z


TRAIN_RANGES = np.array([[-0.03, 0.03], [-0.03, 0.03], [-0.01, 0.01], [0, 0], [0, 0], [-0.5, 0.5]])


@pytest.fixture(scope="module")
def e2e_controller(tmp_path_factory, tracks):
    """The shipped controller's configuration: K4 policy, 8-gate rectangle track, E2E model, gates_ahead=1."""
    from optimal_quad_control_rl_b200 import codegen as G
    w, b, std, _ = k4()
    gp, gy, sp = tracks["indi"]
    spec = G.track_spec(gp, gy, sp, 1, "e2e", TRAIN_RANGES)
    d = str(tmp_path_factory.mktemp("c_code_e2e"))
    files = G.export_controller((w, b, std), spec, d, residual_weights=True)
    assert sorted(os.path.basename(f) for f in files) == [
        "neural_network.c", "neural_network.h", "nn_controller.c", "nn_controller.h", "nn_moment.c", "nn_moment.h",
        "nn_thrust.c", "nn_thrust.h"]
    return G.CController(G.build_controller(d, flags=REF_FLAGS), "e2e"), spec, d


def test_network_forward_matches_oracle_and_kat(e2e_controller):
    from oracle import c_oracle as O
    ctl, _, _ = e2e_controller
    w, b, _, z = k4()
    rng = np.random.default_rng(0)
    obs = np.concatenate([z["kat_in"].reshape(-1, 24), rng.normal(0, 2, (64, 24)).astype(np.float32)])
    ours = np.stack([ctl.forward(o) for o in obs])
    want = np.clip(O.policy_forward(w, b, obs), -1, 1)
    np.testing.assert_array_equal(ours, want)  # same accumulation order, float32: identical bits
    np.testing.assert_allclose(ours[:len(z["kat_out"].reshape(-1, 4))], np.clip(z["kat_out"].reshape(-1, 4), -1, 1),
                               rtol=0, atol=1e-6)


def test_generated_controller_is_bit_identical_to_the_reference_c_code(e2e_controller):
    from oracle import c_oracle as O
    ref = O.ref_policy_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libnn_policy_ref.so not built (reference not mounted)")
    ctl, spec, _ = e2e_controller
    fp = C.POINTER(C.c_float)
    # exported tables
    for name, shape in (("gate_pos", (8, 3)), ("gate_yaw", (8,)), ("start_pos", (3,)), ("output_std", (4,))):
        n = int(np.prod(shape))
        a = np.ctypeslib.as_array((C.c_float * n).in_dll(ctl.lib, name)).reshape(shape)
        r = np.ctypeslib.as_array((C.c_float * n).in_dll(ref, name)).reshape(shape)
        np.testing.assert_array_equal(a, r, err_msg=name)
    C.c_bool.in_dll(ref, "deterministic").value = True
    ctl.deterministic = True
    assert ctl.deterministic
    rng = np.random.default_rng(1)
    ref.nn_reset()
    ctl.reset()
    # a random walk around the track so that gate switches happen; both controllers see the same inputs
    ws = np.zeros(16, np.float32)
    switches = 0
    for t in range(4000):
        g = spec.gate_pos[ctl.target_gate_index]
        ws[0:3] = g + rng.normal(0, 0.8, 3)
        ws[3:6] = rng.normal(0, 3, 3)
        ws[6:8] = rng.uniform(-1, 1, 2)
        ws[8] = rng.uniform(-12, 12)  # heading is never wrapped in world_states: exercises the while-loops
        ws[9:12] = rng.normal(0, 2, 3)
        ws[12:16] = rng.uniform(3000, 11000, 4)
        d = rng.uniform(-1, 1, 4).astype(np.float32) * np.array([0.03, 0.03, 0.01, 0.5], np.float32)
        a, r = np.zeros(4, np.float32), np.zeros(4, np.float32)
        before = ctl.target_gate_index
        ctl.lib.nn_control(ws.ctypes.data_as(fp), d.ctypes.data_as(fp), a.ctypes.data_as(fp))
        ref.nn_control(ws.ctypes.data_as(fp), d.ctypes.data_as(fp), r.ctypes.data_as(fp))
        assert ctl.target_gate_index == C.c_uint8.in_dll(ref, "target_gate_index").value
        switches += ctl.target_gate_index != before
        assert np.array_equal(a, r), (t, a, r)
    assert switches > 100
    # with noise: both draw from libc rand() in the same order
    C.c_bool.in_dll(ref, "deterministic").value = False
    ctl.deterministic = False
    libc = C.CDLL(None)
    out = []
    for lib_ in (ctl.lib, ref):
        libc.srand(7)
        lib_.nn_reset()
        a = np.zeros(4, np.float32)
        lib_.nn_control(ws.ctypes.data_as(fp), d.ctypes.data_as(fp), a.ctypes.data_as(fp))
        out.append(a)
    assert np.array_equal(out[0], out[1])
    C.c_bool.in_dll(ref, "deterministic").value = True


@pytest.mark.parametrize("variant,ga", [("e2e", 1), ("e2e", 0), ("indi", 1), ("indi", 2)])
def test_controller_closed_loop_agrees_with_oracle_env(tmp_path, tracks, variant, ga):
    """The controller's own observation transform + network, driven with the oracle env's world state, must command
    what the env's observation fed through the restated network commands (the notebook's closed loop, `:4494-4519`)."""
    from optimal_quad_control_rl_b200 import codegen as G
    from oracle import c_oracle as O
    gp, gy, sp = tracks[variant]
    env = O.OracleEnv(variant, 1, gp, gy, sp, gates_ahead=ga)
    if variant == "e2e":
        env.disturbance_ranges = TRAIN_RANGES
    rng = np.random.default_rng(5)
    dims = [env.state_len, 48, 48, 4]
    w = [rng.normal(0, 1 / np.sqrt(i), (o, i)).astype(np.float32) for i, o in zip(dims[:-1], dims[1:])]
    b = [rng.normal(0, 0.1, o).astype(np.float32) for o in dims[1:]]
    G.export_controller((w, b, np.full(4, 0.5, np.float32)), env, str(tmp_path))
    ctl = G.CController(G.build_controller(str(tmp_path), flags=REF_FLAGS), variant)
    ctl.deterministic = True
    np.random.seed(3)
    obs = env.reset()
    ctl.reset()
    worst, compared = 0.0, 0
    for t in range(300):
        dist = env.disturbances[0, [0, 1, 2, 5]] if variant == "e2e" else None
        a_c = ctl.control(env.world_states[0], dist)
        a_o = np.clip(O.policy_forward(w, b, obs), -1, 1)[0]
        if ctl.target_gate_index == env.target_gates[0]:  # the controller switches on the plane test alone
            worst = max(worst, float(np.abs(a_c - a_o).max()))
            compared += 1
        obs, _, done, _ = env.step(a_o[None])
        if done[0]:
            ctl.reset()
    assert compared >= 200
    assert worst < 5e-5, worst  # physical-unit round trip (rpm / rad/s) in float32 + libm vs NumPy trig of the gate yaw


def test_residual_models_known_answer(e2e_controller):
    """K1: state [0,1,2,3,4,5,0,0,0,9..15] -> thrust 36.098232, moment [0.2847767, -0.22512697, -0.05896095]."""
    ctl, _, _ = e2e_controller
    k = golden("kat")
    fp = C.POINTER(C.c_float)
    xt = np.array([12, 13, 14, 15, 3, 4, 5], np.float32)
    xm = np.array([12, 13, 14, 15, 3, 4, 5, 9, 10, 11], np.float32)
    t, m = np.zeros(1, np.float32), np.zeros(3, np.float32)
    ctl.lib.nn_thrust_forward(xt.ctypes.data_as(fp), t.ctypes.data_as(fp))
    ctl.lib.nn_moment_forward(xm.ctypes.data_as(fp), m.ctypes.data_as(fp))
    np.testing.assert_allclose(t, k["k1_thrust"][0], rtol=1e-6)
    np.testing.assert_allclose(m, k["k1_moment"][0], rtol=1e-6, atol=1e-7)
    from oracle import c_oracle as O
    ref = O.ref_mlp_lib()
    if ref is not None:  # and bit-identical to the reference's generated nn_thrust.c / nn_moment.c
        t2, m2 = np.zeros(1, np.float32), np.zeros(3, np.float32)
        ref.nn_thrust_forward(xt.ctypes.data_as(fp), t2.ctypes.data_as(fp))
        ref.nn_moment_forward(xm.ctypes.data_as(fp), m2.ctypes.data_as(fp))
        assert np.array_equal(t, t2) and np.array_equal(m, m2)


def test_generator_rejects_inconsistent_inputs(tracks):
    from optimal_quad_control_rl_b200 import codegen as G
    gp, gy, sp = tracks["e2e"]
    spec = G.track_spec(gp, gy, sp, 1, "e2e", TRAIN_RANGES)
    w, b, std, _ = k4()
    with pytest.raises(ValueError):
        G.network_sources(w, b[:-1])
    with pytest.raises(ValueError):
        G.network_sources([w[0], w[0]], [b[0], b[0]])  # 120 -> 24 does not chain
    with pytest.raises(ValueError):
        G.export_controller((w, b, std), G.track_spec(gp, gy, sp, 2, "e2e"), "/tmp/_never_written")  # 24 != 28 inputs
    with pytest.raises(ValueError):
        G.controller_sources(spec, std, variant="indi", disturbance_input=True)
    src, hdr = G.controller_sources(G.track_spec(gp, gy, sp, 1, "e2e"), std)  # empty ranges are widened, not divided by
    assert "d_min[4] = {-1.0, -1.0, -1.0, -1.0}" in src and "extern uint8_t target_gate_index;" in hdr
    src, hdr = G.network_sources(w, b, activation="tanh")
    assert "nn_tanh(y" in src and "void nn_forward(const float* input, float* output);" in hdr


@pytest.mark.parametrize("variant", ["e2e", "indi"])
def test_track_spec_reproduces_the_reference_gate_tables(variant, tracks):
    """track_spec (the host-side __init__ precompute, `3D quad race.ipynb:309-319`) against the frozen reference tables
    (K3: for the rectangle track they are the constants baked into `c_code/nn_controller.c:40-60`)."""
    from optimal_quad_control_rl_b200 import codegen as G
    k = golden("kat")
    gp, gy, sp = tracks[variant]
    spec = G.track_spec(gp, gy, sp, 1, variant)
    np.testing.assert_array_equal(spec.gate_pos_rel, k[f"{variant}_gate_pos_rel"])
    np.testing.assert_array_equal(spec.gate_yaw_rel, k[f"{variant}_gate_yaw_rel"])
    assert spec.gate_pos.dtype == np.float32 and spec.num_gates == len(gy)


def test_ppo_log_summary_finds_the_plateau(tmp_path):
    """tools/ppo_summary.py on a synthetic log: plateau = first 5-iteration mean within 2 % of the final level."""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ppo_summary
    rows = []
    for i in range(100):
        r = 100.0 * min(1.0, i / 40.0)
        rows.append({"iteration": i, "timesteps": (i + 1) * 1000, "wall_s": 1.0 + i, "collect_s": 0.01, "train_s": 0.99,
                     "ep_rew_mean": r, "ep_len_mean": 1200.0, "gates_per_episode": r / 10, "pg_loss": 0.0})
    p = tmp_path / "log.jsonl"
    p.write_text("\n".join(json.dumps(r) for r in rows))
    s = ppo_summary.summarise(str(p))
    assert s["iterations"] == 100 and s["final_ep_rew_mean"] == 100.0 and not s["nan"]
    assert 40.0 <= s["plateau_wall_s"] <= 46.0 and s["train_s_per_iter"] == 0.99
