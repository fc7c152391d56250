#!/usr/bin/env python
"""BASELINE config 5: PPO (reference hyper-parameters) on the GPU env, wall-clock to reward.  One JSON line per
iteration on stdout.  usage: tools/train_ppo.py [--variant e2e|indi] [--num-envs N] [--n-steps T] [--seconds S]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="e2e", choices=["e2e", "indi"])
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--batch-size", type=int, default=1 << 18)
    ap.add_argument("--n-epochs", type=int, default=10)
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--iterations", type=int, default=None)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--amp", action="store_true", help="BF16 autocast of the update's GEMMs")
    ap.add_argument("--bootstrap", default="none", choices=["none", "sb3_a8"])
    ap.add_argument("--update", default="auto", choices=["auto", "torch", "fused"])
    ap.add_argument("--evaluate", default="auto", choices=["auto", "torch", "device", "mixed"],
                    help="values / old log-probs over the rollout buffer (see PPO.__init__)")
    ap.add_argument("--save", default=None, help="checkpoint path written at the end (model.save)")
    a = ap.parse_args()
    import optimal_quad_control_rl_b200 as Q
    if a.variant == "e2e":
        gp, gy, sp = Q.zigzag_track()
        env = Q.Quadcopter3DGates(a.num_envs, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=0)
        env.disturbance_ranges = Q.training_disturbance_ranges()
    else:
        gp, gy, sp = Q.rectangle_track()
        env = Q.Quadcopter3DGatesINDI(a.num_envs, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=0)
    import torch
    pk = dict(activation_fn=torch.nn.ReLU, net_arch=[dict(pi=[120, 120, 120], vf=[120, 120, 120])], log_std_init=0)
    ppo = Q.PPO("MlpPolicy", Q.VecMonitor(env), policy_kwargs=pk, n_steps=a.n_steps, batch_size=a.batch_size,
                n_epochs=a.n_epochs, gamma=0.999, learning_rate=a.lr, amp=a.amp, bootstrap=a.bootstrap, update=a.update, evaluate=a.evaluate)
    ppo.learn(wall_clock_s=None if a.iterations else a.seconds, iterations=a.iterations,
              log=lambda r: print(json.dumps(r), flush=True))
    if a.save:
        ppo.save(a.save)


if __name__ == "__main__":
    main()
