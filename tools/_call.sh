timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/d11_pytest.log
python tools/bench_numpy_step.py 2>&1 | tee gpurun_out/d11_numpy_step.log
QS_HOST_THREADS=1 python tools/bench_numpy_step.py 2>&1 | tail -1 | sed 's/^/1 staging thread: /' | tee -a gpurun_out/d11_numpy_step.log
./build/exp/policy_stages 2>&1 | tee gpurun_out/d11_policy_stages.log | tail -3
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" 
