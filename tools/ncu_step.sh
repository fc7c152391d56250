#!/bin/bash
# usage: tools/ncu_step.sh <tag>
#   gpurun_out/<tag>_launches.csv            every kernel launch of a short default bench (gpu__time_duration.sum)
#   gpurun_out/<tag>_step_{e2e,indi}.ncu-rep one steady-state launch of the step kernel, --set full + source
tag=${1:-ncu}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 40 --warmup 20 --graph 0 --no-cpu-baseline --e2e-steps 3 > gpurun_out/${tag}_launches_bench.log 2>&1
for v in e2e indi; do
  ncu --set full --import-source on --clock-control none -k regex:step_kernel -s 60 -c 1 -o gpurun_out/${tag}_step_$v -f \
      python bench.py --variant $v --steps 8 --warmup 60 --graph 0 --no-cpu-baseline --e2e-steps 3 > gpurun_out/${tag}_ncu_$v.log 2>&1
done
true
