#!/bin/bash
# DRAM traffic of step launches in the steady state (ncu --cache-control none: caches are NOT flushed between passes),
# with and without the L2 residency hints
tag=${1:-l2n}
for v in indi e2e; do for h in 1 0; do
  QS_L2_HINTS=$h ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      -k regex:step_kernel -s 60 -c 3 --csv --log-file gpurun_out/${tag}_${v}_hints$h.csv \
      python bench.py --variant $v --steps 8 --warmup 60 --graph 0 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1
  echo "== $v hints=$h"; grep -E "dram__bytes|duration|hit_rate" gpurun_out/${tag}_${v}_hints$h.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done; done 2>&1 | tee gpurun_out/${tag}_summary.log
