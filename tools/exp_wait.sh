#!/bin/bash
# back-off between polls of the MMA-completion mbarrier in policy_kernel / rollout_kernel
tag=${1:-w}
for ns in 0 20 50 100 200 400; do
  for w in rollout_fused policy; do
  QS_MMA_WAIT_NS=$ns timeout 200 python bench.py --workload $w --steps 256 --warmup 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w wait_ns=$ns: %.2f us/step' % (1e3*d['ms_per_step']))"
  done
done 2>&1 | tee gpurun_out/${tag}_wait.log
QS_MMA_WAIT_NS=50 timeout 300 python -m pytest tests/test_gpu_rollout_fused.py tests/test_gpu_policy.py -x -q 2>&1 | tail -3
