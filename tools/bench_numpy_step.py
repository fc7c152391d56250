#!/usr/bin/env python
"""env-steps/s of the NumPy-facing `env.step(actions)` (what SB3 calls), reset_rng="device": pinned host fast path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimal_quad_control_rl_b200 as Q

for n in (100, 4096, 65536, 1 << 20):
    gp, gy, sp = Q.zigzag_track()
    env = Q.Quadcopter3DGates(n, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=0)
    env.disturbance_ranges = Q.training_disturbance_ranges()
    env.reset()
    acts = [np.random.default_rng(i).uniform(-1, 1, (n, 4)).astype(np.float32) for i in range(4)]
    for i in range(5):
        env.step(acts[i & 3])
    k = 200 if n <= 65536 else 30
    t0 = time.perf_counter()
    for i in range(k):
        env.step(acts[i & 3])
    dt = time.perf_counter() - t0
    print(f"env.step NumPy path N={n}: {1e3 * dt / k:.3f} ms/step  {n * k / dt:.3g} env-steps/s")
    env.close()
