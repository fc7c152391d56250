#!/bin/bash
tag=${1:-q2}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.log
timeout 200 python tools/bench_numpy_step.py 2>&1 | tail -5 | tee gpurun_out/${tag}_numpy_step.log
timeout 200 python bench.py --variant indi --steps 400 --warmup 50 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -c 900
