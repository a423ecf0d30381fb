set -u
mkdir -p gpurun_out
L=flatmatch-global-illumination_b200/lib
python -m pytest tests -m gpu -x -q -k "closest or paths or counters or synth4000 or planes or random or small_bake or counter_is_opt" > gpurun_out/t1_B.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1_B.log
tail -2 gpurun_out/t1_B.log
for v in B A; do
if [ $v = A ]; then cp $L/libfmgi_cuda_A.so $L/libfmgi_cuda.so; fi
python bench.py --no-cpu --no-app --steps 4 --warmup 3 --e2e-steps 0 --workload synth4000_1e9x4 2>gpurun_out/b1.err | tail -1 > gpurun_out/b10_synth_$v.json
python bench.py --no-cpu --no-app --steps 5 --warmup 3 --e2e-steps 0 2>>gpurun_out/b1.err | tail -1 > gpurun_out/b10_example_$v.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b10_*.json")):
    try:
        d=json.loads(open(f).read()); print(f, "%.4g"%d["value"], d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
