set -u
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench2.err | tail -1 > gpurun_out/bench_r1_default_final.json
python bench.py --no-cpu --no-app --steps 4 --warmup 3 --e2e-steps 1 --workload synth4000_1e9x4 2>>gpurun_out/bench2.err | tail -1 > gpurun_out/bench_r1_synth4000_final.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r1_*_final.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d.get("example_bake_wall_s"))
    except Exception as e: print(f, "ERR", e)
PY
