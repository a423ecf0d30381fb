for k in 2 3 4; do
 for carve in -1 50 65 80 100; do
  FMGI_POOL_K=$k FMGI_POOL_CARVEOUT=$carve python bench.py --no-cpu --no-app --no-secondary --steps 5 --warmup 2 --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K $k carveout $carve kernel_ms %.3f value %.4g' % (d['kernel_ms_per_step'], d['value']))"
 done
done
