#!/bin/bash
# usage: tools/gpu_round.sh <tag> : the round's evidence in one gpurun call (tests, smoke, bench lines, ncu)
tag=${1:-r}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_e2e.json 2> gpurun_out/${tag}_bench_e2e.err; tail -c 1800 gpurun_out/${tag}_bench_e2e.json
python bench.py --variant indi > gpurun_out/${tag}_bench_indi.json 2> gpurun_out/${tag}_bench_indi.err; tail -c 700 gpurun_out/${tag}_bench_indi.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>&1; tail -c 600 gpurun_out/${tag}_bench_reference.json
bash tools/ncu_step.sh ${tag}
