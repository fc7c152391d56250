#!/bin/bash
# usage: tools/exp_gather.sh <ngpus> : NCCL all-gather vs the step kernel's own P2P stores (BASELINE config 4 topology)
n=${1:-2}; port=29520
for m in none nccl p2p; do
  port=$((port+1)); extra="--gather-obs $m"; [ $m = none ] && extra=""
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 200 --warmup 20 $extra --no-cpu-baseline --num-envs ${2:-1048576} > gpurun_out/gather_${n}_$m.log 2>&1
  python - <<PY
import json
ok=False
for l in open("gpurun_out/gather_${n}_$m.log"):
    if l.startswith("{"):
        d=json.loads(l); ok=True; print("$m", "%.1f us/step" % (1e3*d["ms_per_step"]), "%.3g env-steps/s" % d["value"], d["config"]["parallelism"])
if not ok: print("$m FAILED"); print(open("gpurun_out/gather_${n}_$m.log").read()[-1500:])
PY
done
