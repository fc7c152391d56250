#!/bin/bash
# usage: tools/gpu_l2.sh <tag> : L2 residency hints experiment (QS_L2_HINTS) + fused-rollout INDI diagnostic
tag=${1:-l2}
timeout 120 python tools/diag_fused.py indi 4096 25 2>&1 | tail -30 | tee gpurun_out/${tag}_diag_indi.log
timeout 120 python tools/diag_fused.py indi 4096 25 det 2>&1 | tail -30 | tee gpurun_out/${tag}_diag_indi_det.log
QS_L2_HINTS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_hints.log
for v in e2e indi; do for n in 262144 1048576 2097152 4194304; do for h in 0 1; do
  QS_L2_HINTS=$h timeout 200 python bench.py --variant $v --num-envs $n --steps 1000 --warmup 100 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v N=$n hints=$h: %.2f us/step  %.3g env-steps/s  frac %.3f' % (1e3*d['ms_per_step'], d['value'], d['roofline']['frac']))"
done; done; done 2>&1 | tee gpurun_out/${tag}_l2_sweep.log
