set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_gpu or shards" > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log; tail -3 gpurun_out/t2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 2 2>gpurun_out/s2.err | tail -1 > gpurun_out/scale2_example.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-cpu --no-app --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_1e9x4 2>>gpurun_out/s2.err | tail -1 > gpurun_out/scale2_synth.json
tail -3 gpurun_out/s2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale2_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d["ms_per_step"], d["n_gpus"], "e2e %.4g"%d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
