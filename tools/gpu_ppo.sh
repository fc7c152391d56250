#!/bin/bash
# usage: tools/gpu_ppo.sh <tag> [seconds] : row f2 tests + BASELINE config 5 (PPO on N=65536 E2E), TF32 and BF16-autocast updates
tag=${1:-ppo}; secs=${2:-170}
python -m pytest tests/test_gpu_ppo.py -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
timeout 400 python tools/train_ppo.py --variant e2e --num-envs 65536 --seconds $secs --save gpurun_out/${tag}_e2e_model > gpurun_out/${tag}_e2e.jsonl 2> gpurun_out/${tag}_e2e.err
python tools/ppo_summary.py gpurun_out/${tag}_e2e.jsonl | tee gpurun_out/${tag}_e2e_summary.json; tail -n 3 gpurun_out/${tag}_e2e.err
timeout 300 python tools/train_ppo.py --variant e2e --num-envs 65536 --seconds 90 --amp > gpurun_out/${tag}_e2e_amp.jsonl 2> gpurun_out/${tag}_e2e_amp.err
python tools/ppo_summary.py gpurun_out/${tag}_e2e_amp.jsonl | tee gpurun_out/${tag}_e2e_amp_summary.json; tail -n 3 gpurun_out/${tag}_e2e_amp.err
timeout 300 python tools/train_ppo.py --variant indi --num-envs 65536 --seconds 60 > gpurun_out/${tag}_indi.jsonl 2> gpurun_out/${tag}_indi.err
python tools/ppo_summary.py gpurun_out/${tag}_indi.jsonl | tee gpurun_out/${tag}_indi_summary.json; tail -n 3 gpurun_out/${tag}_indi.err
