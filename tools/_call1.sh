timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/e3_pytest.log
bash tools/gpu.sh multi e3 2
