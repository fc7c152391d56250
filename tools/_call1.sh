timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/d13_pytest.log
B="python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-configs --e2e-steps 3"
source <(sed -n '/^line()/,/^}/p' tools/gpu.sh)
{ $B 2>/dev/null | line "e2e product"; $B --variant indi 2>/dev/null | line "indi product"; } | tee gpurun_out/d13_variants.log
