"""Rows f2 -> f3 end to end, on the CPU: the policy PPO trained on the GPU simulator (130 s on one B200 with the
tcgen05 update kernels, profiles/r2/ppo/) is exported to the reference's flight-controller C files, compiled with gcc, and flown closed loop
against the CPU ORACLE env (the restatement pinned to the reference) -- the reference's own acceptance test
(`3D quad race.ipynb:4487-4521`: crash rate of the C controller over simulated episodes).  If the GPU simulator's
dynamics, observation transform or reward differed from the reference's, a policy trained there would not fly here.

The network is driven with the ENV's observation (the notebook's `c_network(test_env.states[0])` path, `:4426-4432`,
`:4496-4498`): on the zigzag track gate 6 and gate 0 are the same opening, and `nn_control` -- which advances its
target on the plane test alone (`c_code/nn_controller.c:78-88`) -- skips gate 0 right after gate 6 while the env
waits for a second pass, so the controller's own gate bookkeeping and the env's part ways after one lap (true of the
reference's generated controller too; its observation transform is pinned separately in tests/test_codegen.py)."""
import os

import numpy as np

TRAIN_RANGES = np.array([[-0.03, 0.03], [-0.03, 0.03], [-0.01, 0.01], [0, 0], [0, 0], [-0.5, 0.5]])


def test_gpu_trained_policy_flies_the_oracle_env_through_generated_c(tmp_path, tracks):
    from optimal_quad_control_rl_b200 import codegen as G
    from oracle import c_oracle as O
    data = os.path.join(os.path.dirname(G.__file__), "data", "policy_e2e_zigzag_ppo.npz")
    z = np.load(data)
    n = len(z["dims"]) - 1
    w, b = [z[f"W{l}"] for l in range(n)], [z[f"b{l}"] for l in range(n)]
    gp, gy, sp = tracks["e2e"]
    env = O.OracleEnv("e2e", 1, gp, gy, sp, gates_ahead=1, pause_if_collision=False)
    env.disturbance_ranges = TRAIN_RANGES
    G.export_controller((w, b, z["std"]), env, str(tmp_path), residual_weights=True)
    ctl = G.CController(G.build_controller(str(tmp_path)), "e2e")
    ctl.deterministic = True
    gates, crashes, lengths = [], 0, []
    for ep in range(6):
        np.random.seed(100 + ep)
        obs = env.reset()
        ctl.reset()
        passed = 0
        for t in range(1200):
            a = ctl.forward(obs[0])  # clipped nn_forward of the generated C
            tg0 = env.target_gates[0]
            obs, rew, done, _ = env.step(a[None].astype(np.float32))
            if done[0]:
                crashes += t + 1 < env.max_steps
                break
            passed += env.target_gates[0] != tg0
        gates.append(passed); lengths.append(t + 1)
    print("gates per episode", gates, "lengths", lengths, "crashes", crashes)
    assert crashes <= 1                      # the reference reports the crash rate of its C controller the same way
    assert np.mean(gates) >= 12              # 15.2 per episode in training (with exploration noise)
