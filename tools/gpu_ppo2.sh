#!/bin/bash
# PPO hyper-parameter trials for BASELINE config 5 (wall-clock to reward): smaller rollouts / minibatches = more updates per second
tag=${1:-p2}
trial() { name=$1; shift
  timeout 300 python tools/train_ppo.py --variant e2e --num-envs 65536 "$@" > gpurun_out/${tag}_$name.jsonl 2> gpurun_out/${tag}_$name.err
  python tools/ppo_summary.py gpurun_out/${tag}_$name.jsonl | tee gpurun_out/${tag}_${name}_summary.json | cut -c1-1200; tail -n 2 gpurun_out/${tag}_$name.err
}
trial n32_b64k_amp --n-steps 32 --batch-size 65536 --amp --seconds 100
trial n64_b128k_amp --n-steps 64 --batch-size 131072 --amp --seconds 100
