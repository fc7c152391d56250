#!/bin/bash
tag=${1:-ns}
for v in e2e indi; do
for lib in "" build/exp/libquadsim_nocompute.so; do
for n in 262144 524288 1048576 2097152 4194304 8388608; do
  QS_LIB=$lib python bench.py --variant $v --num-envs $n --steps 400 --warmup 40 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '${lib:-product}', $n, '%.2f us' % (1e3*d['ms_per_step']), 'frac %.3f' % d['roofline']['frac'])"
done; done; done
