#!/bin/bash
# Reduced multi-GPU matrix for the end of round 1 (GPU minutes are charged x8):
#   gpurun --gpus 8 -- bash profiles/run_scale_8.sh
# 8-GPU points of the weak-scaling workloads, the hi-res atlas, the fixed-total photon sweep, and the in-library
# multi-GPU bake; the 1/2-GPU points come from 1- and 2-GPU calls.  Appends JSON lines to gpurun_out/scale8_*.jsonl.
set -u
OUT=gpurun_out
mkdir -p $OUT
run() {  # n, output file, bench args...
  local n=$1 out=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
       bench.py --gpus $n --no-cpu --no-app "$@" 2>>$OUT/scale8.err | tail -1 >> $out
}
run 8 $OUT/scale8_example_1e8x3.jsonl --steps 5 --warmup 3 --e2e-steps 2
run 4 $OUT/scale8_example_1e8x3.jsonl --steps 5 --warmup 3 --e2e-steps 2
run 8 $OUT/scale8_synth4000_1e9x4.jsonl --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_1e9x4
run 8 $OUT/scale8_synth4000_hires.jsonl --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_hires_1e9x4
for tp in 1e6 1e7 1e8 1e9 1e10; do
  run 8 $OUT/scale8_sweep_example_x8.jsonl --steps 3 --warmup 3 --e2e-steps 1 --workload example_default_x8 --total-photons $tp
done
python -m pytest tests -m gpu -x -q -k "multi_gpu" > $OUT/t8.log 2>&1; tail -2 $OUT/t8.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale8_*.jsonl")):
    for l in open(f):
        try:
            d = json.loads(l); print(f.split("/")[-1], d["n_gpus"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], "kernel %.3f" % d["kernel_ms_per_step"], "e2e %.4g" % d["e2e"]["value"])
        except Exception as e:
            print(f, "ERR", e)
PY
