#!/bin/bash
# usage: tools/gpu_round2.sh <tag> : the round's evidence in one gpurun call -- tests, smoke, bench lines, ncu launch list,
# ncu --set full of the step kernels, the policy kernel and the fused rollout kernel
tag=${1:-r}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_e2e.json 2> gpurun_out/${tag}_bench_e2e.err; tail -c 1900 gpurun_out/${tag}_bench_e2e.json
python bench.py --variant indi > gpurun_out/${tag}_bench_indi.json 2> gpurun_out/${tag}_bench_indi.err; tail -c 900 gpurun_out/${tag}_bench_indi.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>&1; tail -c 400 gpurun_out/${tag}_bench_reference.json
for w in policy rollout_unfused rollout_fused; do
  python bench.py --workload $w --steps 384 --warmup 64 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -c 500 gpurun_out/${tag}_bench_$w.json
done
bash tools/ncu_step.sh ${tag}
ncu --set full --import-source on --clock-control none -k regex:policy_kernel -s 20 -c 1 -o gpurun_out/${tag}_policy -f \
    python bench.py --workload policy --steps 8 --warmup 20 --graph 0 > gpurun_out/${tag}_ncu_policy.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/${tag}_rollout_fused -f \
    python bench.py --workload rollout_fused --rollout-steps 8 --steps 16 --warmup 16 > gpurun_out/${tag}_ncu_rollout.log 2>&1
ls -la gpurun_out/*.ncu-rep
