set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "closest or paths or counters or small_bake or synth800_closest or planes or chunk or edge_cases or general or axis_parallel or tier_fixtures" > gpurun_out/rs_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rs_tests.log
tail -4 gpurun_out/rs_tests.log
for k in 2 3; do
  FMGI_ROOM_STEPS=$k python bench.py --no-cpu --no-app --no-secondary --steps 6 --warmup 3 --e2e-steps 1 2>gpurun_out/rs.err | tail -1 > gpurun_out/rs_example_$k.json
  FMGI_ROOM_STEPS=$k python bench.py --no-cpu --no-app --no-secondary --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_1e9x4 2>>gpurun_out/rs.err | tail -1 > gpurun_out/rs_synth_$k.json
done
tail -2 gpurun_out/rs.err
python - <<'PY'
import json
for k in (2, 3):
    for f in ("example", "synth"):
        try:
            d = json.loads(open(f"gpurun_out/rs_{f}_{k}.json").read())
            print(k, f, "value %.4g kernel_ms %.3f e2e %.4g tests/ray %.3f" % (d["value"], d["kernel_ms_per_step"], d["e2e"]["value"], d["roofline"]["rect_tests_per_ray"]))
        except Exception as e:
            print(k, f, "failed", e)
PY
