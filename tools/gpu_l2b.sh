#!/bin/bash
# usage: tools/gpu_l2b.sh <tag> : all GPU tests after the explicit-arithmetic rewrite, then the L2 pinning budget sweep
tag=${1:-l2b}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest_all.log
QS_L2_HINTS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_hints.log
run() { # variant n hints keep
  QS_L2_HINTS=$3 QS_L2_KEEP_MB=$4 timeout 200 python bench.py --variant $1 --num-envs $2 --steps 1000 --warmup 100 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 N=$2 hints=$3 keep=$4MB: %.2f us/step  %.3g env-steps/s  frac %.3f' % (1e3*d['ms_per_step'], d['value'], d['roofline']['frac']))"
}
{
run e2e 1048576 0 0
for k in 32 48 56 64 72 80 100; do run e2e 1048576 1 $k; done
run indi 1048576 0 0
for k in 40 56 64 80; do run indi 1048576 1 $k; done
run e2e 4194304 0 0
for k in 48 64; do run e2e 4194304 1 $k; done
run indi 4194304 1 56
run e2e 65536 1 56
run e2e 65536 0 56
} 2>&1 | tee gpurun_out/${tag}_l2_sweep.log
