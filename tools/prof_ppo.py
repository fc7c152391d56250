#!/usr/bin/env python
"""Where does a PPO update spend its GPU time? (torch.profiler over one train() call; diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optimal_quad_control_rl_b200 as Q

amp = "--amp" in sys.argv
gp, gy, sp = Q.zigzag_track()
env = Q.Quadcopter3DGates(65536, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=0)
env.disturbance_ranges = Q.training_disturbance_ranges()
pk = dict(activation_fn=torch.nn.ReLU, net_arch=[dict(pi=[120, 120, 120], vf=[120, 120, 120])], log_std_init=0)
ppo = Q.PPO("MlpPolicy", env, policy_kwargs=pk, n_steps=128, batch_size=1 << 18, n_epochs=2, gamma=0.999, amp=amp, update="fused" if "--fused" in sys.argv else "torch")
ppo.collect_rollouts(); ppo.train(); torch.cuda.synchronize()
import time
t0 = time.perf_counter(); ppo.collect_rollouts(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("collect_s", t1 - t0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ppo.train()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
