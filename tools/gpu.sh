#!/bin/bash
# ONE script for everything that runs on the GPU box through gpurun:   gpurun -- 'bash tools/gpu.sh <cmd> <tag> [args]'
# Results land in gpurun_out/<tag>_*.  Commands:
#   check              pytest -m gpu + smoke()
#   bench              the bench lines: default (E2E 2^20 + configs), INDI, the reference arm, policy / rollout workloads
#   ncu                launch list of a short bench + `ncu --set full` of the E2E and INDI step kernels (+ steady-state DRAM traffic)
#   ncu-tc             `ncu --set full` of the tcgen05 kernels (policy, fused rollout)
#   ncu-train          `ncu --set full` of the PPO training kernel + launch list of one update
#   variants [libs..]  bench.py over experimental builds / launch modes of the step kernel (QS_LIB, QS_* env switches)
#   nsweep  [libs..]   per-step time vs N (2^16 .. 2^22) for the product library and each extra library
#   numpy              the NumPy-facing env.step (what SB3 calls) at N = 100 .. 2^20
#   ppo <seconds> [train_ppo.py args]   BASELINE config 5: PPO to the plateau, log + summary
#   multi <ngpus>      multi-GPU test + bench.py under torchrun (ours, reference arm)
#   sanitize [secs]    compute-sanitizer: memcheck over the GPU tests, racecheck + synccheck over smoke()
#   hostpipe           qs_step_host: host-path tests + sweep of (chunks, first-chunk divisor) on the e2e number
# Experimental libraries are built HERE (CPU box) first:   bash tools/gpu.sh build-variants name=-DFLAG ...
cmd=${1:-check}; tag=${2:-r}; shift 2 2>/dev/null
mkdir -p gpurun_out
line() {  # line <label> : one-line summary of the bench JSON on stdin
python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
d=[json.loads(l) for l in t if l.startswith('{')]
if not d: print('$1 FAILED', t[-3:]); sys.exit()
d=d[-1]; r=d.get('roofline') or {}; e=d.get('e2e') or {}
print('$1: %.2f us/step  %.4g %s  frac %s  e2e %s  clk %s' % (1e3*d['ms_per_step'], d['value'], d['unit'], ('%.3f' % r['frac']) if r else '-', ('%.3g' % e['value']) if e else '-', (d.get('clocks') or {}).get('sm_mhz')))
for c in d.get('configs', []): print('   ', c['name'], '%.2f us/step' % (1e3*c['ms_per_step']), '%.4g' % c['value'], 'frac', '%.3f' % (c.get('roofline') or {}).get('frac', c.get('per_gpu_roofline_frac', 0)), c.get('ms_per_step_l2_flushed', ''))
if 'selfcheck' in d: print('    selfcheck', d['selfcheck'])
if 'cpu_baseline' in d: print('    cpu', d['cpu_baseline']['kind'], '%.4g' % d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:60], '| port %.4g' % d.get('cpu_port', {}).get('value', 0))
"
}
Q="--no-cpu-baseline --no-configs --e2e-steps 3"
case $cmd in
build-variants)
  mkdir -p build/exp
  python - "$tag" "$@" <<'PY'
import sys; sys.path.insert(0, '.')
from optimal_quad_control_rl_b200 import build as B
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    print(B.build_library(out=f"build/exp/libquadsim_{name}.so", extra_flags=tuple(flags.split(",")) if flags else ()))
PY
  ;;
check)
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.log
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/${tag}_smoke.log
  ;;
bench)
  timeout 600 python bench.py > gpurun_out/${tag}_bench_e2e.json 2> gpurun_out/${tag}_bench_e2e.err; line e2e < gpurun_out/${tag}_bench_e2e.json; tail -2 gpurun_out/${tag}_bench_e2e.err
  timeout 300 python bench.py --variant indi --no-configs --no-cpu-baseline > gpurun_out/${tag}_bench_indi.json 2> gpurun_out/${tag}_bench_indi.err; line indi < gpurun_out/${tag}_bench_indi.json
  timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2>gpurun_out/${tag}_bench_reference.err; line reference < gpurun_out/${tag}_bench_reference.json
  for w in policy rollout_unfused rollout_fused; do
    timeout 300 python bench.py --workload $w --steps 384 --warmup 64 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; line $w < gpurun_out/${tag}_bench_$w.json
  done
  ;;
ncu)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
      python bench.py --steps 40 --warmup 20 --graph 0 $Q > gpurun_out/${tag}_launches_bench.log 2>&1
  for v in e2e indi; do
    ncu --set full --import-source on --clock-control none -k regex:step_kernel -s 60 -c 1 -o gpurun_out/${tag}_step_$v -f \
        python bench.py --variant $v --steps 8 --warmup 60 --graph 0 $Q > gpurun_out/${tag}_ncu_$v.log 2>&1
    # steady-state DRAM traffic: caches NOT flushed between launches (what consecutive steps really move)
    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        -k regex:step_kernel -s 60 -c 3 --csv --log-file gpurun_out/${tag}_traffic_$v.csv \
        python bench.py --variant $v --steps 8 --warmup 60 --graph 0 $Q > /dev/null 2>&1
  done
  ls -la gpurun_out/${tag}_*.ncu-rep
  ;;
ncu-tc)
  ncu --set full --import-source on --clock-control none -k regex:policy_kernel -s 20 -c 1 -o gpurun_out/${tag}_policy -f \
      python bench.py --workload policy --steps 8 --warmup 20 --graph 0 > gpurun_out/${tag}_ncu_policy.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/${tag}_rollout_fused -f \
      python bench.py --workload rollout_fused --rollout-steps 8 --steps 16 --warmup 16 > gpurun_out/${tag}_ncu_rollout.log 2>&1
  ls -la gpurun_out/${tag}_*.ncu-rep
  ;;
ncu-train)
  ncu --set full --import-source on --clock-control none -k regex:ppo_grad_kernel -s 30 -c 1 -o gpurun_out/${tag}_ppo_grad -f \
      python tools/train_ppo.py --num-envs 65536 --iterations 2 --update fused > gpurun_out/${tag}_ncu_train.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ppo_ -s 200 -c 40 --csv --log-file gpurun_out/${tag}_train_launches.csv \
      python tools/train_ppo.py --num-envs 65536 --iterations 2 --update fused > /dev/null 2>&1
  ls -la gpurun_out/${tag}_*.ncu-rep
  ;;
variants)
  B="python bench.py --steps 1000 --warmup 100 $Q"
  { $B | line e2e_base
    $B --no-stats | line e2e_nostats
    QS_PDL=0 $B | line e2e_nopdl
    $B --graph 0 | line e2e_nograph
    $B --variant indi | line indi_base
    for lib in "$@"; do
      QS_LIB=build/exp/libquadsim_$lib.so $B | line e2e_$lib
      QS_LIB=build/exp/libquadsim_$lib.so $B --variant indi | line indi_$lib
    done; } 2>&1 | tee gpurun_out/${tag}_variants.log
  ;;
nsweep)
  for lib in "" "$@"; do for v in e2e indi; do for n in 65536 262144 524288 1048576 2097152 4194304; do
    QS_LIB=${lib:+build/exp/libquadsim_$lib.so} python bench.py --variant $v --num-envs $n --steps 400 --warmup 40 $Q 2>/dev/null | line "$v ${lib:-product} N=$n"
  done; done; done 2>&1 | tee gpurun_out/${tag}_nsweep.log
  ;;
numpy)
  python tools/bench_numpy_step.py 2>&1 | tee gpurun_out/${tag}_numpy_step.log
  ;;
ppo)
  secs=${1:-120}; shift
  timeout $((secs + 200)) python tools/train_ppo.py --variant e2e --num-envs 65536 --seconds $secs "$@" --save gpurun_out/${tag}_model \
      > gpurun_out/${tag}_ppo.jsonl 2> gpurun_out/${tag}_ppo.err
  python tools/ppo_summary.py gpurun_out/${tag}_ppo.jsonl | tee gpurun_out/${tag}_ppo_summary.json | cut -c1-1500; tail -n 2 gpurun_out/${tag}_ppo.err
  ;;
multi)
  n=${1:-2}
  timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_multi.log
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 400 --warmup 50 \
      > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; line "${n}gpu" < gpurun_out/${tag}_bench_${n}gpu.json; tail -3 gpurun_out/${tag}_bench_${n}gpu.err
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --impl reference --steps 5 --warmup 1 2>/dev/null | tail -c 600
  ;;
hostpipe)
  timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "host_step or numpy_step" 2>&1 | tail -3
  for cd in "4 1" "4 4" "4 8" "5 4" "6 4" "4 2" "3 4" "4 1" "4 4"; do
    set -- $cd
    QS_HOST_CHUNKS=$1 QS_HOST_FIRST_DIV=$2 timeout 60 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 80 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $1 first_div $2 e2e %.4g  us/step(host) %.1f' % (d['e2e']['value'], 1048576/d['e2e']['value']*1e6))"
  done | tee gpurun_out/${tag}_host_first_div_sweep.log
  ;;
sanitize)
  secs=${1:-110}
  for grp in "step:tests/test_gpu_parity.py tests/test_gpu_chain.py tests/test_gpu_packed_obs.py" \
             "tensor:tests/test_gpu_policy.py tests/test_gpu_rollout_fused.py" "train:tests/test_gpu_train.py tests/test_gpu_ppo.py"; do
    name=${grp%%:*}; files=${grp#*:}
    timeout $secs compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 7 \
        python -m pytest $files -m gpu -q -p no:cacheprovider --timeout 120 > gpurun_out/${tag}_memcheck_$name.log 2>&1
    echo "memcheck $name rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned|Timeout" gpurun_out/${tag}_memcheck_$name.log | sort | uniq -c | head -12
  done
  for tool in racecheck synccheck; do
    timeout 110 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 7 --launch-timeout 100 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "SUMMARY|smoke|hazard|Barrier error" gpurun_out/${tag}_$tool.log | sort | uniq -c | head -12
  done
  ;;
*) echo "unknown command $cmd"; exit 2;;
esac
