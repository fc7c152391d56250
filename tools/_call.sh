timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/d13_pytest.log
bash tools/gpu.sh multi d13 2
