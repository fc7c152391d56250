#!/bin/bash
# builds the experimental libraries tools/exp_variants.sh and tools/exp_nsweep.sh select with QS_LIB (run here, on the CPU box)
mkdir -p build/exp
python - <<'PY'
import sys; sys.path.insert(0, '.')
from optimal_quad_control_rl_b200 import build as B
for name, flag in (("nomlp", "-DQS_EXP_NOMLP"), ("nocompute", "-DQS_EXP_NOCOMPUTE")):
    print(B.build_library(out=f"build/exp/libquadsim_{name}.so", extra_flags=(flag,)))
PY
