#!/bin/bash
# usage: tools/gpu_fused.sh <tag> : fused closed-loop rollout kernel -- tests, then fused vs unfused timings
tag=${1:-fz}
timeout 300 python -m pytest tests/test_gpu_rollout_fused.py -x -q 2>&1 | tail -25 | tee gpurun_out/${tag}_pytest_fused.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_rollout_fused.py 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest_all.log
for n in 1048576 65536; do
for w in rollout rollout_unfused rollout_fused; do
  timeout 200 python bench.py --workload $w --num-envs $n --steps 384 --warmup 64 > gpurun_out/${tag}_bench_${w}_$n.json 2> gpurun_out/${tag}_bench_${w}_$n.err
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/${tag}_bench_${w}_$n.json').read().strip().splitlines()[-1]); print('$w N=$n: %.2f us/step  %.3g env-steps/s  launches %d' % (1e3*d['ms_per_step'], d['value'], d['gpu_launches']))
except Exception as e: print('$w N=$n failed', e); print(open('gpurun_out/${tag}_bench_${w}_$n.err').read()[-800:])"
done; done
