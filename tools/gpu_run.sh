set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_full.log
tail -5 gpurun_out/t_full.log
python bench.py --steps 10 --warmup 3 2>gpurun_out/b6.err | tail -1 > gpurun_out/b6_default.json
python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 1 --workload synth4000_1e9x4 2>>gpurun_out/b6.err | tail -1 > gpurun_out/b6_synth.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b6_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d["ms_per_step"], d["config"].get("tier"), d["roofline"].get("rect_tests_per_ray"), "e2e %.4g"%d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_r1e_example -f python bench.py --no-cpu --no-app --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/ncu_e1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_r1e_synth4000 -f python bench.py --no-cpu --no-app --steps 1 --warmup 1 --e2e-steps 0 --workload synth4000_1e9x4 > gpurun_out/ncu_e2.log 2>&1
