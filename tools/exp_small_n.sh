#!/bin/bash
# BASELINE configs 2 and 3 (small N: launch-latency regime) + rollout at small N
for cfg in "e2e 4096" "e2e 65536" "indi 262144" "indi 4096"; do set -- $cfg
  for g in 0 20; do
  python bench.py --variant $1 --num-envs $2 --steps 2000 --warmup 200 --graph $g --no-cpu-baseline --e2e-steps 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 N=$2 graph=$g: %.2f us/step  %.3g env-steps/s  frac %.3f  e2e(host) %.3g' % (1e3*d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']))"
  done
done
for n in 4096 65536; do
  python bench.py --workload rollout --num-envs $n --steps 2000 --warmup 200 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rollout N=$n: %.2f us/step  %.3g env-steps/s' % (1e3*d['ms_per_step'], d['value']))"
  python bench.py --workload policy --num-envs $n --steps 2000 --warmup 200 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('policy N=$n: %.2f us/step' % (1e3*d['ms_per_step']))"
done
