"""Row f3 cross-check (CPU): the reference's generated flight controller (`c_code/nn_controller.c:68-179`, compiled
where it lies into oracle/_ref) is a second, independent statement of the observation transform
(`update_states_gate`, `3D quad race.ipynb:365-450`) and of the policy network.  Driven with the oracle env's world
state it must command what the oracle's observation fed through the restated network commands."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden


def test_reference_controller_agrees_with_oracle_observation_and_network(tracks):
    from oracle import c_oracle as O
    lib = O.ref_policy_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libnn_policy_ref.so not built (reference not mounted)")
    z = golden("policy_k4")
    w = [z[f"W{l}"] for l in range(4)]
    b = [z[f"b{l}"] for l in range(4)]
    gp, gy, sp = tracks["indi"]  # nn_controller.c:14-38 bakes in the 8-gate rectangle track, E2E model, gates_ahead=1
    C.c_bool.in_dll(lib, "deterministic").value = True
    fp = C.POINTER(C.c_float)
    env = O.OracleEnv("e2e", 1, gp, gy, sp, gates_ahead=1)
    env.disturbance_ranges = np.array([[-0.03, 0.03], [-0.03, 0.03], [-0.01, 0.01], [0, 0], [0, 0], [-0.5, 0.5]])
    np.random.seed(0)
    obs = env.reset()
    lib.nn_reset()
    worst = 0.0
    for t in range(60):
        ws = env.world_states[0].copy()
        d = np.ascontiguousarray(env.disturbances[0, [0, 1, 2, 5]])
        ws[12:16] = (ws[12:16] + 1) / 2 * 8000 + 3000          # the controller takes rpm (`nn_controller.c:117-122`)
        rpm = np.zeros(4, np.float32)
        lib.nn_control(ws.ctypes.data_as(fp), d.ctypes.data_as(fp), rpm.ctypes.data_as(fp))
        a_ref = (rpm - 3000) / 8000 * 2 - 1                     # undo `nn_controller.c:175`
        a_ours = np.clip(O.policy_forward(w, b, obs), -1, 1)[0]
        if env.target_gates[0] != C.c_uint8.in_dll(lib, "target_gate_index").value:
            break  # the controller switches gates on the plane test alone (`:82-92`), the env also needs the box
        worst = max(worst, float(np.abs(a_ref - a_ours).max()))
        obs, _, done, _ = env.step(a_ours[None])
        if done[0]:
            break
    assert t >= 20
    assert worst < 2e-5, worst  # rpm round trip through float32
