timeout 300 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout_fused.py tests/test_gpu_ppo.py -q 2>&1 | tail -5
python tools/_probe_nvlink.py 2>&1 | tail -12
bash tools/gpu.sh numpy c6
bash tools/gpu.sh ncu c6
bash tools/gpu.sh ncu-tc c6
bash tools/gpu.sh bench c6
