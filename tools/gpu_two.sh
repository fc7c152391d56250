#!/bin/bash
# 2-GPU sanity: the multi-GPU test, bench.py under torchrun (ours + reference arm), fused rollout on both ranks
tag=${1:-g2}
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 50 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err; tail -c 1500 gpurun_out/${tag}_bench_2gpu.json; tail -2 gpurun_out/${tag}_bench_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 20 --warmup 3 2>/dev/null | tail -c 500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload rollout_fused --steps 128 --warmup 32 2>/dev/null | tail -c 700
