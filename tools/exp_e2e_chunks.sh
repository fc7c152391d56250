for c in 1 2 4 8 16; do QS_HOST_CHUNKS=$c python bench.py --steps 40 --warmup 10 --no-cpu-baseline --e2e-steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $c e2e %.4g' % d['e2e']['value'])"; done
