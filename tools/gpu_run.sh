set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "closest or paths or counters or synth or general or planes or occlusion or random or small_bake or edge" > gpurun_out/t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1.log
tail -5 gpurun_out/t1.log
python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 1 --workload synth4000_1e9x4 2>gpurun_out/b1.err | tail -1 > gpurun_out/b8_synth.json
python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 1 2>>gpurun_out/b1.err | tail -1 > gpurun_out/b8_example.json
tail -3 gpurun_out/b1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b8_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d["ms_per_step"], d["config"].get("tier"), d["roofline"].get("rect_tests_per_ray"), d["roofline"].get("issue"))
    except Exception as e: print(f, "ERR", e)
PY
