set -u
mkdir -p gpurun_out
for tp in 1e6 1e7 1e8 1e9 1e10; do
python bench.py --no-cpu --no-app --steps 3 --warmup 3 --e2e-steps 1 --workload example_default_x8 --total-photons $tp 2>>gpurun_out/sw1.err | tail -1 >> gpurun_out/scale1_sweep_example_x8.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/scale1_sweep_example_x8.jsonl"):
    d=json.loads(l); print(d["n_gpus"], "%.4g"%d["value"], "%.3f ms"%d["ms_per_step"])
PY
