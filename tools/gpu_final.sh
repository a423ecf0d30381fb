# Round-end capture on one B200 (gpurun -- bash tools/gpu_final.sh): full GPU suite, smoke, the bench lines that go
# to profiles/, the ncu launch list of the bench command and one --set full capture per headline workload.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_full.log
tail -4 gpurun_out/t_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_r1_default.json
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench.err | tail -1 > gpurun_out/bench_r1_reference.json
python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 2 --workload synth4000_1e9x4 2>>gpurun_out/bench.err | tail -1 > gpurun_out/bench_r1_synth4000.json
python bench.py --no-cpu --no-app --steps 3 --warmup 3 --e2e-steps 1 --workload synth4000_hires_1e9x4 2>>gpurun_out/bench.err | tail -1 > gpurun_out/bench_r1_synth4000_hires.json
python bench.py --no-cpu --no-app --steps 3 --warmup 3 --e2e-steps 1 --workload example_default_x8 2>>gpurun_out/bench.err | tail -1 > gpurun_out/bench_r1_example_default_x8.json
tail -3 gpurun_out/bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r1_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d.get("ms_per_step"), d.get("config",{}).get("tier"), "e2e %.4g"%d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench.csv python bench.py --no-cpu --no-app --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_r1_bench_example -f python bench.py --no-cpu --no-app --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_r1_bench_synth4000 -f python bench.py --no-cpu --no-app --steps 1 --warmup 1 --e2e-steps 0 --workload synth4000_1e9x4 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
