# Round-2 capture on one B200 (gpurun -- bash tools/gpu_final_r2.sh): full GPU suite, smoke, the bench lines that go
# to profiles/ (ours with the secondary workloads, the reference arm), the ncu launch list of the bench command and one
# --set full capture per headline workload, on the SAME build (kernel hash in every line).
set -u
export TAG=${TAG:-r2g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_reference.json
python bench.py --steps 20 --warmup 3 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_default.json
python bench.py --no-cpu --no-app --no-secondary --steps 3 --warmup 3 --workload example_default_x8 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_example_default_x8.json
tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json, os
TAG = os.environ["TAG"]
d = json.loads(open(f"gpurun_out/{TAG}_bench_default.json").read())
print("ours: value %.4g e2e %.4g kernel_ms %.3f hash %s stale %s" % (d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d["src_hash"], "ncu_stale" in d["roofline"]))
print("example_bake", d.get("example_bake"))
r = json.loads(open(f"gpurun_out/{TAG}_bench_reference.json").read())
print("reference: %.4g on %d cores" % (r["value"], r["cpu_baseline"]["cores"]), "ratio e2e %.0f" % (d["e2e"]["value"] / r["value"]))
for k, v in d.get("secondary", {}).items():
    if isinstance(v, dict):
        print(k, "%.4g" % v["value"], "kernel %.2f ms" % v["kernel_ms_per_step"], "e2e %.4g" % v["e2e"]["value"], "e2e/device %.3f" % (v["e2e"]["value"] / v["value"]))
    else:
        for q in v: print(k, q)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu --no-app --no-secondary --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_${TAG}_example -f python bench.py --no-cpu --no-app --no-secondary --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_${TAG}_synth4000 -f python bench.py --no-cpu --no-app --no-secondary --steps 1 --warmup 1 --e2e-steps 0 --workload synth4000_1e9x4 > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/prof_${TAG}*.ncu-rep
