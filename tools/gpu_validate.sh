#!/bin/bash
# usage: tools/gpu_validate.sh <tag> : GPU tests, smoke, default bench lines, BASELINE configs 2/3 (small N) and 5 (PPO)
tag=${1:-v}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_e2e.json 2> gpurun_out/${tag}_bench_e2e.err; tail -c 1800 gpurun_out/${tag}_bench_e2e.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>&1; tail -c 600 gpurun_out/${tag}_bench_reference.json
bash tools/exp_small_n.sh 2>&1 | tee gpurun_out/${tag}_small_n.log
timeout 300 python tools/train_ppo.py --variant e2e --num-envs 65536 --seconds 150 > gpurun_out/${tag}_ppo_e2e.jsonl 2> gpurun_out/${tag}_ppo_e2e.err
tail -n 3 gpurun_out/${tag}_ppo_e2e.jsonl; tail -n 5 gpurun_out/${tag}_ppo_e2e.err
