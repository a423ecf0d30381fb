for c in 32 64 128 256; do
 for ph in 1.25e7 1.25e6 1e8; do
  FMGI_CHUNK=$c python bench.py --no-cpu --no-app --no-secondary --steps 10 --warmup 3 --e2e-steps 0 --workload example_default_x8 --photons $ph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c photons $ph kernel_ms %.4f value %.4g' % (d['kernel_ms_per_step'], d['value']))"
 done
done
