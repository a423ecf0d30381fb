# multi-GPU pass (gpurun --gpus N -- bash tools/gpu_r2_multi.sh N)
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "in_library or pooled or chunk" > gpurun_out/r2m_tests_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_tests_$N.log
tail -4 gpurun_out/r2m_tests_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/r2m_bench_$N.err | tail -1 > gpurun_out/r2m_bench_$N.json
tail -3 gpurun_out/r2m_bench_$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2m_bench_{n}.json").read())
print("N", d["n_gpus"], "value %.4g" % d["value"], "ms/step %.3f kernel %.3f" % (d["ms_per_step"], d["kernel_ms_per_step"]), "e2e %.4g" % d["e2e"]["value"], d["e2e"]["breakdown_ms"], "first", d["e2e"]["first_call_ms"], d["e2e"]["first_call_init_ms"])
for k, v in d.get("secondary", {}).items():
    if isinstance(v, dict):
        print(k, "%.4g" % v["value"], "ms/step %.2f kernel %.2f non-kernel %.2f" % (v["ms_per_step"], v["kernel_ms_per_step"], v["non_kernel_ms_per_step"]), "e2e %.4g" % v["e2e"]["value"], v["e2e"]["breakdown_ms"])
    else:
        for q in v: print(k, q)
PY
