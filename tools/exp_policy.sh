#!/bin/bash
# policy-forward and rollout measurements (row f1)
tag=${1:-pol}
for w in policy rollout; do
  python bench.py --workload $w --steps 400 --warmup 40 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  tail -c 900 gpurun_out/${tag}_bench_$w.json; tail -3 gpurun_out/${tag}_bench_$w.err
done
QS_POLICY_CTAS_PER_SM=1 python bench.py --workload policy --steps 400 --warmup 40 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 CTA/SM: %.1f us frac %.3f' % (1e3*d['ms_per_step'], d['roofline']['frac']))"
ncu --set full --import-source on --clock-control none -k regex:policy_kernel -s 20 -c 1 -o gpurun_out/${tag}_policy -f \
    python bench.py --workload policy --steps 8 --warmup 20 --graph 0 > gpurun_out/${tag}_ncu_policy.log 2>&1
true
