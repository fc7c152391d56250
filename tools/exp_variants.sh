#!/bin/bash
# Runs bench.py over experimental builds / launch modes of the step kernel; one line per variant in gpurun_out/$1_*.log
tag=${1:-exp}; mkdir -p gpurun_out
run() { name=$1; shift; "$@" > gpurun_out/${tag}_${name}.log 2>&1; python - <<PY
import json
ok=False
for l in open("gpurun_out/${tag}_${name}.log"):
    if l.startswith("{"):
        d=json.loads(l); ok=True; print("${name}", "%.2f us/step" % (1e3*d["ms_per_step"]), "frac %.3f" % d["roofline"]["frac"], "e2e %.3g" % d["e2e"]["value"], d["clocks"]["sm_mhz"])
if not ok: print("${name} FAILED"); print(open("gpurun_out/${tag}_${name}.log").read()[-600:])
PY
}
B="python bench.py --steps 500 --warmup 50 --no-cpu-baseline --e2e-steps 3"
run e2e_base $B
QS_PDL=0 run e2e_nopdl $B
run e2e_nograph $B --graph 0
QS_STAGES=3 run e2e_s3 $B
QS_CTAS_PER_SM=4 run e2e_c4 $B
QS_LIB=build/exp/libquadsim_nomlp.so run e2e_nomlp $B
QS_LIB=build/exp/libquadsim_nocompute.so run e2e_nocompute $B
run indi_base $B --variant indi
QS_PDL=0 run indi_nopdl $B --variant indi
run indi_nograph $B --variant indi --graph 0
QS_STAGES=3 run indi_s3 $B --variant indi
QS_STAGES=4 run indi_s4 $B --variant indi
QS_LIB=build/exp/libquadsim_nocompute.so run indi_nocompute $B --variant indi
QS_STAGES=3 QS_LIB=build/exp/libquadsim_nocompute.so run indi_nocompute_s3 $B --variant indi
