set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "closest or paths or counters or synth or general or planes or occlusion or random or small_bake or edge" > gpurun_out/t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1.log
tail -5 gpurun_out/t1.log
for tb in 4 3; do
FMGI_TUNE_BLOCKS=$tb python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 1 --workload synth4000_1e9x4 2>gpurun_out/b1.err | tail -1 > gpurun_out/b2_synth_tb$tb.json
FMGI_TUNE_BLOCKS=$tb FMGI_TIER=2 python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 1 2>>gpurun_out/b1.err | tail -1 > gpurun_out/b2_example_grid_tb$tb.json
done
python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 1 2>>gpurun_out/b1.err | tail -1 > gpurun_out/b2_example_soup.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b2_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d["ms_per_step"], d["config"].get("tier"), d["roofline"].get("rect_tests_per_ray"))
    except Exception as e: print(f, "ERR", e)
PY
