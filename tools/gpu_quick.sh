#!/bin/bash
# quick check of the policy / fused rollout kernels: tests + two bench lines
tag=${1:-q}
timeout 400 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout_fused.py tests/test_gpu_ppo.py -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.log
for w in policy rollout_fused; do
  timeout 200 python bench.py --workload $w --steps 384 --warmup 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w: %.2f us/step  %.3g /s' % (1e3*d['ms_per_step'], d['value']))"
done 2>&1 | tee gpurun_out/${tag}_bench.log
