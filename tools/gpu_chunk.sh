set -u
mkdir -p gpurun_out
for c in 64 128 256 512 1024; do
  FMGI_CHUNK=$c python bench.py --no-cpu --no-app --no-secondary --steps 8 --warmup 3 --e2e-steps 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c example kernel_ms %.3f' % d['kernel_ms_per_step'])"
done
for c in 128 256 512; do
  FMGI_CHUNK=$c python bench.py --no-cpu --no-app --no-secondary --steps 3 --warmup 2 --e2e-steps 0 --workload synth4000_1e9x4 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c synth kernel_ms %.3f' % d['kernel_ms_per_step'])"
done
