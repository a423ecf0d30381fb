# Round-2 scaling matrix on one 8-GPU box (gpurun --gpus 8 -- bash tools/gpu_scale_r2.sh): the driver-shaped bench line
# (default workload + secondary: synth4000, hi-res, fixed-total sweep) at N = 1, 2, 4, 8.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "in_library" > gpurun_out/r2s_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_tests.log; tail -2 gpurun_out/r2s_tests.log
python bench.py --no-cpu --no-app --steps 10 --warmup 3 2>gpurun_out/r2s_1.err | tail -1 > gpurun_out/r2s_scale_1.json
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r2s_$N.err | tail -1 > gpurun_out/r2s_scale_$N.json
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r2s_scale_{n}.json").read())
    except Exception as e:
        print(n, "ERR", e); continue
    print("N", n, "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
    for k, v in d.get("secondary", {}).items():
        if isinstance(v, dict):
            print("   ", k, "%.4g" % v["value"], "ms %.2f kernel %.2f" % (v["ms_per_step"], v["kernel_ms_per_step"]), "e2e %.4g" % v["e2e"]["value"], "fold %.2f d2h %.2f" % (v["e2e"]["breakdown_ms"]["fold"], v["e2e"]["breakdown_ms"]["atlas_d2h"]))
        else:
            print("   ", k, [(q["total_photons"], "%.4g" % q["value"], "%.3f" % q["kernel_ms_per_step"]) for q in v])
PY
