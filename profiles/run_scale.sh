#!/bin/bash
# Multi-GPU measurement matrix (one box, up to 8 B200): run under `gpurun --gpus 8 -- bash profiles/run_scale.sh`.
# Writes JSON lines to gpurun_out/scale_*.jsonl.
#   1. weak scaling 1/2/4/8 GPUs on example_1e8x3 and synth4000_1e9x4 (BASELINE configs[1], [2])
#   2. hi-res 1.83 GB atlas at 8 GPUs: the NCCL atlas reduce under load (configs[3])
#   3. photon-count sweep 1e6..1e10 total photons x 8 bounces on example.png at 1 and 8 GPUs (configs[4])
set -u
OUT=gpurun_out
mkdir -p $OUT
run() {  # n, output file, bench args...
  local n=$1 out=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --no-cpu --no-app "$@" 2>>$OUT/scale.err | tail -1 >> $out
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
       bench.py --gpus $n --no-cpu --no-app "$@" 2>>$OUT/scale.err | tail -1 >> $out; fi
}
NG=$(python -c "import torch; print(torch.cuda.device_count())")
for n in 1 2 4 8; do [ $n -le $NG ] || continue
  run $n $OUT/scale_example_1e8x3.jsonl --steps 5 --warmup 3 --e2e-steps 2
  run $n $OUT/scale_synth4000_1e9x4.jsonl --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_1e9x4
done
[ 8 -le $NG ] && run 8 $OUT/scale_synth4000_hires.jsonl --steps 3 --warmup 2 --e2e-steps 1 --workload synth4000_hires_1e9x4
for tp in 1e6 1e7 1e8 1e9 1e10; do
  for n in 1 8; do [ $n -le $NG ] || continue
    run $n $OUT/scale_sweep_example_x8.jsonl --steps 3 --warmup 3 --e2e-steps 1 --workload example_default_x8 --total-photons $tp
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*.jsonl")):
    print(f)
    for l in open(f):
        d = json.loads(l)
        print("  gpus %d  photons/gpu %.3g  depth %d  value %.4g  ms/step %.3f  kernel %.3f  e2e %.4g  scaling %s" % (
            d["n_gpus"], d["config"]["photons_per_gpu_per_step"], d["config"]["depth"], d["value"], d["ms_per_step"],
            d["kernel_ms_per_step"], d["e2e"]["value"], d["scaling"]))
PY
