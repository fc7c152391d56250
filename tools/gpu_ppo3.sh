#!/bin/bash
# BASELINE config 5, long run to the plateau (BF16-autocast update) + 2-GPU sanity of bench.py and the multi-GPU test
tag=${1:-p3}
timeout 500 python tools/train_ppo.py --variant e2e --num-envs 65536 --seconds 330 --amp --save gpurun_out/${tag}_e2e_model > gpurun_out/${tag}_e2e_amp_long.jsonl 2> gpurun_out/${tag}_e2e_amp_long.err
python tools/ppo_summary.py gpurun_out/${tag}_e2e_amp_long.jsonl | tee gpurun_out/${tag}_e2e_amp_long_summary.json | cut -c1-1500; tail -n 2 gpurun_out/${tag}_e2e_amp_long.err
