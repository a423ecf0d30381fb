set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "closest or paths or counters or small_bake or synth800_closest or planes or pooled or chunk or edge_cases or general" > gpurun_out/quick_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_tests.log
tail -4 gpurun_out/quick_tests.log
python bench.py --no-cpu --no-app --no-secondary --steps 10 --warmup 3 --e2e-steps 2 2>gpurun_out/quick.err | tail -1 > gpurun_out/quick_example.json
python bench.py --no-cpu --no-app --no-secondary --steps 4 --warmup 2 --e2e-steps 1 --workload synth4000_1e9x4 2>>gpurun_out/quick.err | tail -1 > gpurun_out/quick_synth.json
tail -2 gpurun_out/quick.err
python - <<'PY'
import json
for f in ("quick_example", "quick_synth"):
    d = json.loads(open(f"gpurun_out/{f}.json").read())
    print(f, "value %.4g kernel_ms %.3f e2e %.4g" % (d["value"], d["kernel_ms_per_step"], d["e2e"]["value"]))
PY
