# Round 2, first GPU pass (gpurun -- bash tools/gpu_r2_a.sh): GPU suite, PCIe probe, default bench line with the
# secondary workloads, ncu --set full of the trace kernel on both headline workloads, launch list.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python - > gpurun_out/r2a_pcie.log 2>&1 <<'PY'
import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, h in (("pinned", torch.empty(n, dtype=torch.uint8).pin_memory()), ("pageable", torch.empty(n, dtype=torch.uint8))):
    h.fill_(1)
    for direction in ("h2d", "d2h"):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        print(f"{name} {direction}: {n / best / 1e9:.1f} GB/s")
PY
cat gpurun_out/r2a_pcie.log
python bench.py --steps 10 --warmup 3 2>gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench_default.json
tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2a_bench_default.json").read())
print("value %.4g e2e %.4g kernel_ms %.3f wall %s" % (d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d.get("example_bake")))
print("e2e breakdown", d["e2e"].get("breakdown_ms"), d["e2e"].get("first_call_ms"))
for k, v in d.get("secondary", {}).items():
    if isinstance(v, dict):
        print(k, "%.4g" % v["value"], "kernel %.2f ms" % v["kernel_ms_per_step"], "e2e %.4g" % v["e2e"]["value"], v["e2e"]["breakdown_ms"], "create %.1f" % v["scene_create_ms"])
    else:
        for q in v: print(k, q)
PY
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_r2a_example -f python bench.py --no-cpu --no-app --no-secondary --steps 1 --warmup 1 --e2e-steps 0 > gpurun_out/r2a_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_r2a_synth4000 -f python bench.py --no-cpu --no-app --no-secondary --steps 1 --warmup 1 --e2e-steps 0 --workload synth4000_1e9x4 > gpurun_out/r2a_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --no-cpu --no-app --no-secondary --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/r2a_ncu_launches.log 2>&1
ls -la gpurun_out/prof_r2a*.ncu-rep
