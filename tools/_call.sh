B="python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-configs --e2e-steps 3"
source <(sed -n '/^line()/,/^}/p' tools/gpu.sh)
{ for lib in head "" nouni head "" nouni; do
  QS_LIB=${lib:+build/exp/libquadsim_$lib.so} $B 2>/dev/null | line "e2e ${lib:-product}"
  QS_LIB=${lib:+build/exp/libquadsim_$lib.so} $B --variant indi 2>/dev/null | line "indi ${lib:-product}"
done; } 2>&1 | tee gpurun_out/d4_variants.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/d4_pytest.log
