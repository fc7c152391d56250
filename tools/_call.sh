timeout 600 python -m pytest tests/test_gpu_train.py -q -s 2>&1 | tail -80
timeout 600 python -m pytest tests/test_gpu_ppo.py -q 2>&1 | tail -15
