set -u
mkdir -p gpurun_out
for k in ${STEPS:-2 3}; do
  FMGI_ROOM_STEPS=$k python bench.py --no-cpu --no-app --no-secondary --steps 8 --warmup 3 --e2e-steps 1 2>gpurun_out/rp.err | tail -1 > gpurun_out/rp_example_$k.json
  FMGI_ROOM_STEPS=$k python bench.py --no-cpu --no-app --no-secondary --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_1e9x4 2>>gpurun_out/rp.err | tail -1 > gpurun_out/rp_synth_$k.json
done
tail -2 gpurun_out/rp.err
python - <<'PY'
import json, os
for k in os.environ.get("STEPS", "2 3").split():
    for f in ("example", "synth"):
        try:
            d = json.loads(open(f"gpurun_out/rp_{f}_{k}.json").read())
            print(k, f, "value %.4g kernel_ms %.3f e2e %.4g tests/ray %.3f" % (d["value"], d["kernel_ms_per_step"], d["e2e"]["value"], d["roofline"]["rect_tests_per_ray"]))
        except Exception as e:
            print(k, f, "failed", e)
PY
